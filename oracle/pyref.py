"""ctypes binding of oracle/_ref/libref_slam.so: the REFERENCE'S OWN hot-path sources compiled verbatim
(oracle/ref_slam.cpp, oracle/Makefile target `ref`) -- TEST INFRASTRUCTURE ONLY.

Used by tests/ (and tools/gen_golden.py) to pin the oracle restatement against the real reference code.
`lib()` returns None when the library is neither prebuilt nor buildable (no /root/reference): tests then fall
back to the golden fixtures that tools/gen_golden.py wrote from it.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from . import pyoracle as po

_DIR = Path(__file__).resolve().parent
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = _DIR / "_ref" / "libref_slam.so"
        if not so.exists():
            if os.path.isdir("/root/reference/openvslam"):
                subprocess.check_call(["make", "-C", str(_DIR), "all"], stdout=subprocess.DEVNULL)
            if not so.exists():
                return None
        po.lib()   # liborb_oracle.so first (libref_slam.so links the cv2-pinned primitives from it)
        L = C.CDLL(str(so))
        L.ref_match_loop_closures.restype = C.c_uint
        L.ref_match_triangulation.restype = C.c_uint
        L.ref_replace_duplication.restype = C.c_uint
        L.ref_bow_create.restype = C.c_void_p
        L.ref_bench_extract.restype = C.c_double
        L.ref_bench_match.restype = C.c_double
        _lib = L
    return _lib


def available():
    return lib() is not None


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


def _u32(a):
    return np.ascontiguousarray(a, np.uint32)


def settings(levels=8, scale_factor=1.2, max_keypoints=2000):
    """StaticSettings: (scaleFactors, levelSigmaSq, maxNumberOfKeypointsPerLevel)."""
    s = np.zeros(levels, np.float32)
    q = np.zeros(levels, np.float32)
    b = np.zeros(levels, np.int32)
    lib().ref_settings(levels, C.c_float(scale_factor), max_keypoints, _p(s), _p(q), _p(b))
    return s, q, b


def features_around(x, y, qx, qy, r):
    x = _f32(x); y = _f32(y)
    out = np.zeros(max(len(x), 1), np.int32)
    n = lib().ref_features_around(_p(x), _p(y), len(x), C.c_float(qx), C.c_float(qy), C.c_float(r), _p(out))
    return out[:n].copy()


def pyramid(p, img):
    img = np.ascontiguousarray(img, np.uint8)
    _, w, h, _ = po.geometry(p)
    total = int((w.astype(np.int64) * h).sum())
    pyr = np.empty(total, np.uint8)
    blur = np.empty(total, np.uint8)
    n = lib().ref_pyramid(C.byref(p), _p(img), img.strides[0], _p(pyr), _p(blur))
    assert n == p.levels
    lv, bl, off = [], [], 0
    for l in range(p.levels):
        k = int(w[l]) * int(h[l])
        lv.append(pyr[off:off + k].reshape(h[l], w[l]))
        bl.append(blur[off:off + k].reshape(h[l], w[l]))
        off += k
    return lv, bl


def extract(p, img, tracks=None, track_ids=None, track_level=0, valid_rect=None):
    """OrbExtractor::detectAndExtract of the reference, end to end."""
    img = np.ascontiguousarray(img, np.uint8)
    nt = 0 if tracks is None else len(tracks)
    cap = 2 * p.max_keypoints + nt + 1024
    x = np.empty(cap, np.float32); y = np.empty(cap, np.float32); a = np.empty(cap, np.float32)
    o = np.empty(cap, np.int32); d = np.empty((cap, 8), np.uint32); tid = np.empty(cap, np.int32)
    txy = None if nt == 0 else _f32(tracks)
    tids = None if nt == 0 else _i32(track_ids if track_ids is not None else np.arange(nt))
    vr = None if valid_rect is None else np.ascontiguousarray(valid_rect, np.float64)
    n = lib().ref_extract(C.byref(p), _p(img), img.strides[0], _p(txy), _p(tids), nt, int(track_level), _p(vr),
                          _p(x), _p(y), _p(a), _p(o), _p(d), _p(tid), cap)
    assert n <= cap
    return dict(n=n, x=x[:n].copy(), y=y[:n].copy(), angle=a[:n].copy(), octave=o[:n].copy(), desc=d[:n].copy(),
                track_id=tid[:n].copy())


def match_loop_closures(dA, aA, nodeA, dB, aB, nodeB, statusA=None, statusB=None, ratio=0.8, require_triangulation=True):
    """matchForLoopClosures; status: 0 no map point, 1 TRIANGULATED, 2 NOT_TRIANGULATED (None: all 1)."""
    dA = _u32(dA); dB = _u32(dB); aA = _f32(aA); aB = _f32(aB); nodeA = _i32(nodeA); nodeB = _i32(nodeB)
    sA = None if statusA is None else np.ascontiguousarray(statusA, np.uint8)
    sB = None if statusB is None else np.ascontiguousarray(statusB, np.uint8)
    m = np.empty(max(len(dA), 1), np.int32)
    n = lib().ref_match_loop_closures(_p(dA), _p(aA), _p(nodeA), _p(sA), len(dA), _p(dB), _p(aB), _p(nodeB), _p(sB), len(dB),
                                      C.c_float(ratio), int(require_triangulation), _p(m))
    return int(n), m[:len(dA)]


def match_triangulation(dA, aA, octA, bearA, nodeA, dB, aB, bearB, nodeB, poseA, poseB, has_mpA=None, has_mpB=None,
                        levels=8, scale_factor=1.2, residual_deg_thr=0.2):
    """matchForTriangulationDBoW; returns (count, matches, E) with E the essential matrix the reference built."""
    dA = _u32(dA); dB = _u32(dB); aA = _f32(aA); aB = _f32(aB); octA = _i32(octA)
    bearA = np.ascontiguousarray(bearA, np.float64); bearB = np.ascontiguousarray(bearB, np.float64)
    nodeA = _i32(nodeA); nodeB = _i32(nodeB)
    hA = None if has_mpA is None else np.ascontiguousarray(has_mpA, np.uint8)
    hB = None if has_mpB is None else np.ascontiguousarray(has_mpB, np.uint8)
    pA = np.ascontiguousarray(poseA, np.float64).reshape(16); pB = np.ascontiguousarray(poseB, np.float64).reshape(16)
    m = np.empty(max(len(dA), 1), np.int32)
    E = np.zeros(9, np.float64)
    n = lib().ref_match_triangulation(_p(dA), _p(aA), _p(octA), _p(bearA), _p(nodeA), _p(hA), len(dA), _p(dB), _p(aB),
                                      _p(bearB), _p(nodeB), _p(hB), len(dB), _p(pA), _p(pB), int(levels),
                                      C.c_float(scale_factor), C.c_float(residual_deg_thr), _p(m), _p(E))
    return int(n), m[:len(dA)], E.reshape(3, 3)


def _queries(pos, norm, min_dist, max_dist, qdesc):
    return (np.ascontiguousarray(pos, np.float64).reshape(-1, 3), _f32(norm).reshape(-1, 3), _f32(min_dist), _f32(max_dist),
            _u32(qdesc).reshape(-1, 8))


def search_by_projection(kx, ky, koct, kdesc, pos, norm, min_dist, max_dist, qdesc, threshold, taken=None,
                         levels=8, scale_factor=1.2):
    """searchByProjection; returns (count, matched keypoint per query, qx, qy, qr, predicted level)."""
    kx = _f32(kx); ky = _f32(ky); koct = _i32(koct); kdesc = _u32(kdesc).reshape(-1, 8)
    pos, norm, min_dist, max_dist, qdesc = _queries(pos, norm, min_dist, max_dist, qdesc)
    tk = None if taken is None else np.ascontiguousarray(taken, np.uint8)
    nq = len(qdesc)
    idx = np.zeros(max(nq, 1), np.int32); qx = np.zeros(max(nq, 1), np.float32); qy = np.zeros(max(nq, 1), np.float32)
    qr = np.zeros(max(nq, 1), np.float32); ql = np.zeros(max(nq, 1), np.int32)
    n = lib().ref_search_by_projection(_p(kx), _p(ky), _p(koct), _p(kdesc), len(kx), _p(tk), _p(pos), _p(norm), _p(min_dist),
                                       _p(max_dist), _p(qdesc), nq, C.c_float(threshold), int(levels), C.c_float(scale_factor),
                                       _p(idx), _p(qx), _p(qy), _p(qr), _p(ql))
    return int(n), idx[:nq], qx[:nq], qy[:nq], qr[:nq], ql[:nq]


def replace_duplication(kx, ky, koct, kdesc, pos, norm, min_dist, max_dist, qdesc, margin, kp_mp=None, q_obs=None,
                        levels=8, scale_factor=1.2):
    """replaceDuplication; returns (fused count, final owner per keypoint, qx, qy, qr, predicted level)."""
    kx = _f32(kx); ky = _f32(ky); koct = _i32(koct); kdesc = _u32(kdesc).reshape(-1, 8)
    pos, norm, min_dist, max_dist, qdesc = _queries(pos, norm, min_dist, max_dist, qdesc)
    km = None if kp_mp is None else _i32(kp_mp)
    qo = None if q_obs is None else _i32(q_obs)
    nq = len(qdesc)
    fin = np.zeros(max(len(kx), 1), np.int32); qx = np.zeros(max(nq, 1), np.float32); qy = np.zeros(max(nq, 1), np.float32)
    qr = np.zeros(max(nq, 1), np.float32); ql = np.zeros(max(nq, 1), np.int32)
    n = lib().ref_replace_duplication(_p(kx), _p(ky), _p(koct), _p(kdesc), len(kx), _p(km), _p(pos), _p(norm), _p(min_dist),
                                      _p(max_dist), _p(qdesc), _p(qo), nq, C.c_float(margin), int(levels),
                                      C.c_float(scale_factor), _p(fin), _p(qx), _p(qy), _p(qr), _p(ql))
    return int(n), fin[:len(kx)], qx[:nq], qy[:nq], qr[:nq], ql[:nq]


def match_sim3(x1, y1, oct1, d1, mp1, x2, y2, oct2, d2, mp2, mp_pos, mp_min, mp_max, mp_desc, mp_status=None, seed_pairs=None,
               levels=8, scale_factor=1.2):
    """matchMapPointsSim3 (identity transform / poses); returns (pairs [n, 2] of keypoint indices, q12 [n1, 3], lvl12,
    q21 [n2, 3], lvl21)."""
    x1 = _f32(x1); y1 = _f32(y1); oct1 = _i32(oct1); d1 = _u32(d1).reshape(-1, 8); mp1 = _i32(mp1)
    x2 = _f32(x2); y2 = _f32(y2); oct2 = _i32(oct2); d2 = _u32(d2).reshape(-1, 8); mp2 = _i32(mp2)
    mp_pos = np.ascontiguousarray(mp_pos, np.float64).reshape(-1, 3); mp_min = _f32(mp_min); mp_max = _f32(mp_max)
    mp_desc = _u32(mp_desc).reshape(-1, 8)
    st = None if mp_status is None else _i32(mp_status)
    seeds = np.zeros((0, 2), np.int32) if seed_pairs is None else _i32(seed_pairs).reshape(-1, 2)
    n1, n2 = len(x1), len(x2)
    pairs = np.zeros((max(min(n1, n2), 1), 2), np.int32)
    q12 = np.zeros((max(n1, 1), 3), np.float32); l12 = np.zeros(max(n1, 1), np.int32)
    q21 = np.zeros((max(n2, 1), 3), np.float32); l21 = np.zeros(max(n2, 1), np.int32)
    n = lib().ref_match_sim3(_p(x1), _p(y1), _p(oct1), _p(d1), _p(mp1), n1, _p(x2), _p(y2), _p(oct2), _p(d2), _p(mp2), n2,
                             _p(mp_pos), _p(mp_min), _p(mp_max), _p(mp_desc), _p(st), len(mp_desc), _p(seeds), len(seeds),
                             int(levels), C.c_float(scale_factor), _p(pairs), _p(q12), _p(l12), _p(q21), _p(l21))
    return pairs[:n].copy(), q12[:n1], l12[:n1], q21[:n2], l21[:n2]


def medoid(desc, offsets):
    """MapPoint::updateDescriptor per segment: the selected descriptors [n_seg, 8]."""
    desc = _u32(desc).reshape(-1, 8)
    offsets = np.ascontiguousarray(offsets, np.int64)
    out = np.zeros((max(len(offsets) - 1, 1), 8), np.uint32)
    lib().ref_medoid(_p(desc), _p(offsets), len(offsets) - 1, _p(out))
    return out[:len(offsets) - 1]


class BowIndex:
    """slam::BowIndex of the reference over a DBoW2 text vocabulary file."""

    def __init__(self, vocabulary_txt, min_in_common_ratio=0.8, score_ratio=0.75):
        self.h = C.c_void_p(lib().ref_bow_create(str(vocabulary_txt).encode(), C.c_float(min_in_common_ratio), C.c_float(score_ratio)))

    def transform(self, desc):
        desc = _u32(desc).reshape(-1, 8)
        node = np.zeros(max(len(desc), 1), np.int32)
        w = np.zeros(max(len(desc), 1), np.uint32); v = np.zeros(max(len(desc), 1), np.float64)
        n = lib().ref_bow_transform(self.h, _p(desc), len(desc), _p(node), _p(w), _p(v), len(w))
        return node[:len(desc)], w[:n].copy(), v[:n].copy()

    def add(self, kf_id, word, value):
        word = _u32(word); value = np.ascontiguousarray(value, np.float64)
        lib().ref_bow_add(self.h, int(kf_id), _p(word), _p(value), len(word))

    def remove(self, kf_id):
        lib().ref_bow_remove(self.h, int(kf_id))

    def similar(self, word, value, self_kf=-1, cap=4096):
        word = _u32(word); value = np.ascontiguousarray(value, np.float64)
        kf = np.zeros(cap, np.int32); sc = np.zeros(cap, np.float32)
        n = lib().ref_bow_similar(self.h, _p(word), _p(value), len(word), int(self_kf), _p(kf), _p(sc), cap)
        return kf[:n].copy(), sc[:n].copy()

    def close(self):
        if self.h:
            lib().ref_bow_destroy(self.h)
            self.h = None


def bench_extract(p, imgs, threads):
    """The reference's OrbExtractor::detectAndExtract over a stack of frames, sharded over `threads` host threads."""
    imgs = np.ascontiguousarray(imgs, np.uint8)
    total = C.c_long(0)
    s = lib().ref_bench_extract(C.byref(p), _p(imgs), imgs.shape[0], int(threads), C.byref(total))
    return float(s), total.value


def bench_match(desc, ang, pairs, threads, ratio=0.8):
    """The reference's matchForLoopClosures (single node, every feature eligible) over keyframe pairs."""
    desc = _u32(desc); ang = _f32(ang); pairs = _i32(pairs)
    total = C.c_long(0)
    s = lib().ref_bench_match(_p(desc), _p(ang), desc.shape[0], desc.shape[1], _p(pairs), len(pairs), C.c_float(ratio),
                              int(threads), C.byref(total))
    return float(s), total.value
