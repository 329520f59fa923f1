"""ctypes binding of the CPU oracle (oracle/liborb_oracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (slam-module_b200/) never does.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
MAX_LEVELS = 16


class Params(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("levels", C.c_int),
                ("scale_factor", C.c_float), ("max_keypoints", C.c_int),
                ("ini_fast_thr", C.c_int), ("min_fast_thr", C.c_int)]


def make_params(width, height, levels=8, scale_factor=1.2, max_keypoints=2000, ini_fast_thr=20, min_fast_thr=7):
    return Params(width, height, levels, scale_factor, max_keypoints, ini_fast_thr, min_fast_thr)


def build(force=False):
    so = _DIR / "liborb_oracle.so"
    if force or not so.exists():
        subprocess.check_call(["make", "-C", str(_DIR), "all"], stdout=subprocess.DEVNULL)
    return so


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(build()))
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_ic_angle.restype = C.c_float
        L.orc_util_cos.restype = C.c_float
        L.orc_util_cos.argtypes = [C.c_float]
        L.orc_util_sin.restype = C.c_float
        L.orc_util_sin.argtypes = [C.c_float]
        L.orc_hamming.restype = C.c_uint
        L.orc_match_bruteforce.restype = C.c_uint
        L.orc_angle_bin.argtypes = [C.c_float]
        L.orc_bench_extract.restype = C.c_double
        L.orc_bench_match.restype = C.c_double
        _lib = L
    return _lib


def ref_lib():
    """The reference's own header-only leaves compiled verbatim (oracle/_ref); None if not built."""
    global _ref
    if _ref is None:
        so = _DIR / "_ref" / "libref_leaves.so"
        if not so.exists():
            if os.path.isdir("/root/reference/openvslam"):
                subprocess.check_call(["make", "-C", str(_DIR), "ref"], stdout=subprocess.DEVNULL)
            if not so.exists():
                return None
        R = C.CDLL(str(so))
        R.ref_cos.restype = C.c_float
        R.ref_cos.argtypes = [C.c_float]
        R.ref_sin.restype = C.c_float
        R.ref_sin.argtypes = [C.c_float]
        R.ref_hamming.restype = C.c_uint
        _ref = R
    return _ref


def geometry(p):
    s = np.zeros(MAX_LEVELS, np.float32)
    w = np.zeros(MAX_LEVELS, np.int32)
    h = np.zeros(MAX_LEVELS, np.int32)
    b = np.zeros(MAX_LEVELS, np.int32)
    lib().orc_level_geometry(C.byref(p), s.ctypes, w.ctypes, h.ctypes)
    lib().orc_level_budgets(C.byref(p), b.ctypes)
    n = p.levels
    return s[:n].copy(), w[:n].copy(), h[:n].copy(), b[:n].copy()


def resize(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(src.ctypes, src.shape[1], src.shape[0], src.strides[0], dst.ctypes, dw, dh, dw)
    return dst


def gaussian7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().orc_gaussian7_u8(src.ctypes, src.shape[1], src.shape[0], src.strides[0], dst.ctypes, dst.strides[0])
    return dst


def pyramid(p, img):
    """Returns (levels, blurred) as lists of 2-D uint8 arrays."""
    img = np.ascontiguousarray(img, np.uint8)
    _, w, h, _ = geometry(p)
    total = int((w.astype(np.int64) * h).sum())
    pyr = np.empty(total, np.uint8)
    blur = np.empty(total, np.uint8)
    lib().orc_pyramid(C.byref(p), img.ctypes, img.strides[0], pyr.ctypes, blur.ctypes)
    lv, bl, off = [], [], 0
    for l in range(p.levels):
        n = int(w[l]) * int(h[l])
        lv.append(pyr[off:off + n].reshape(h[l], w[l]))
        bl.append(blur[off:off + n].reshape(h[l], w[l]))
        off += n
    return lv, bl


def cv_fast(img, thr):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = w * h
    xs = np.empty(cap, np.int32)
    ys = np.empty(cap, np.int32)
    rs = np.empty(cap, np.int32)
    n = lib().orc_cv_fast(img.ctypes, w, h, img.strides[0], int(thr), xs.ctypes, ys.ctypes, rs.ctypes, cap)
    return xs[:n].copy(), ys[:n].copy(), rs[:n].copy()


def fast_response_map(img):
    """FAST response (cornerScore) of every pixel at least 3 px from the border; -1 elsewhere."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.full((h, w), -1, np.int32)
    L = lib()
    base = img.ctypes.data
    for y in range(3, h - 3):
        for x in range(3, w - 3):
            out[y, x] = L.orc_fast_response(C.c_void_p(base + y * img.strides[0] + x), img.strides[0])
    return out


def detect_level(img, budget, ini_thr=20, min_thr=7, with_candidates=False):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = max(budget + 64, 64)
    ccap = w * h // 4 + 16
    xs = np.empty(cap, np.int32)
    ys = np.empty(cap, np.int32)
    rs = np.empty(cap, np.int32)
    cx = np.empty(ccap, np.int32)
    cy = np.empty(ccap, np.int32)
    cr = np.empty(ccap, np.int32)
    nc = C.c_int(0)
    n = lib().orc_detect_level(img.ctypes, w, h, img.strides[0], int(budget), int(ini_thr), int(min_thr),
                               xs.ctypes, ys.ctypes, rs.ctypes, cap, cx.ctypes, cy.ctypes, cr.ctypes, ccap,
                               C.byref(nc))
    assert n <= cap and nc.value <= ccap
    out = (xs[:n].copy(), ys[:n].copy(), rs[:n].copy())
    if with_candidates:
        return out, (cx[:nc.value].copy(), cy[:nc.value].copy(), cr[:nc.value].copy())
    return out


def distribute(cx, cy, cr, area_w, area_h, budget):
    cx = np.ascontiguousarray(cx, np.int32)
    cy = np.ascontiguousarray(cy, np.int32)
    cr = np.ascontiguousarray(cr, np.int32)
    cap = len(cx) + 1
    out = np.empty(cap, np.int32)
    n = lib().orc_distribute(cx.ctypes, cy.ctypes, cr.ctypes, len(cx), int(area_w), int(area_h), int(budget),
                             out.ctypes, cap)
    return out[:n].copy()


def ic_angle(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    return float(lib().orc_ic_angle(img.ctypes, img.strides[0], int(x), int(y)))


def ic_moments(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    a = C.c_int()
    b = C.c_int()
    lib().orc_ic_moments(img.ctypes, img.strides[0], int(x), int(y), C.byref(a), C.byref(b))
    return a.value, b.value


def descriptor(blurred, x, y, angle_deg):
    blurred = np.ascontiguousarray(blurred, np.uint8)
    d = np.empty(8, np.uint32)
    lib().orc_descriptor(blurred.ctypes, blurred.strides[0], int(x), int(y), C.c_float(angle_deg), d.ctypes)
    return d


def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint32)
    b = np.ascontiguousarray(b, np.uint32)
    return int(lib().orc_hamming(a.ctypes, b.ctypes))


def extract(p, img, tracks=None, track_ids=None, track_level=0):
    """OrbExtractor::detectAndExtract for one frame; returns a dict of SoA arrays."""
    img = np.ascontiguousarray(img, np.uint8)
    nt = 0 if tracks is None else len(tracks)
    cap = 2 * p.max_keypoints + nt + 1024     # extreme aspect ratios: more initial quadtree nodes than budget
    x = np.empty(cap, np.float32)
    y = np.empty(cap, np.float32)
    a = np.empty(cap, np.float32)
    o = np.empty(cap, np.int32)
    d = np.empty((cap, 8), np.uint32)
    tid = np.empty(cap, np.int32)
    lx = np.empty(cap, np.int32)
    ly = np.empty(cap, np.int32)
    lc = np.zeros(MAX_LEVELS, np.int32)
    txy = None if nt == 0 else np.ascontiguousarray(tracks, np.float32)
    tids = None if nt == 0 else np.ascontiguousarray(track_ids if track_ids is not None else np.arange(nt), np.int32)
    n = lib().orc_extract(C.byref(p), img.ctypes, img.strides[0],
                          None if txy is None else txy.ctypes, None if tids is None else tids.ctypes,
                          nt, int(track_level), x.ctypes, y.ctypes, a.ctypes, o.ctypes, d.ctypes, tid.ctypes,
                          lx.ctypes, ly.ctypes, cap, lc.ctypes)
    assert n <= cap
    return dict(n=n, x=x[:n].copy(), y=y[:n].copy(), angle=a[:n].copy(), octave=o[:n].copy(),
                desc=d[:n].copy(), track_id=tid[:n].copy(), lvl_x=lx[:n].copy(), lvl_y=ly[:n].copy(),
                level_counts=lc[:p.levels].copy())


def match_bruteforce(dA, aA, dB, aB, ratio=0.8, thr=50, check_orientation=True, ratio_is_double=False):
    dA = np.ascontiguousarray(dA, np.uint32)
    dB = np.ascontiguousarray(dB, np.uint32)
    aA = np.ascontiguousarray(aA, np.float32)
    aB = np.ascontiguousarray(aB, np.float32)
    m = np.empty(len(dA), np.int32)
    n = lib().orc_match_bruteforce(dA.ctypes, aA.ctypes, len(dA), dB.ctypes, aB.ctypes, len(dB),
                                   C.c_float(ratio), C.c_uint(thr), int(check_orientation), int(ratio_is_double),
                                   m.ctypes)
    return int(n), m


def match_bow(dA, aA, nodeA, dB, aB, nodeB, eligA=None, eligB=None, ratio=0.8, thr=50, check_orientation=True, ratio_is_double=False):
    dA = np.ascontiguousarray(dA, np.uint32); dB = np.ascontiguousarray(dB, np.uint32)
    aA = np.ascontiguousarray(aA, np.float32); aB = np.ascontiguousarray(aB, np.float32)
    nodeA = np.ascontiguousarray(nodeA, np.int32); nodeB = np.ascontiguousarray(nodeB, np.int32)
    eA = None if eligA is None else np.ascontiguousarray(eligA, np.uint8)
    eB = None if eligB is None else np.ascontiguousarray(eligB, np.uint8)
    m = np.empty(max(len(dA), 1), np.int32)
    L = lib()
    L.orc_match_bow.restype = C.c_uint
    vp = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
    n = L.orc_match_bow(vp(dA), vp(aA), vp(nodeA), vp(eA), len(dA), vp(dB), vp(aB), vp(nodeB), vp(eB), len(dB),
                        C.c_float(ratio), C.c_uint(thr), int(check_orientation), int(ratio_is_double), vp(m))
    return int(n), m[:len(dA)]


def match_triangulation(dA, aA, octA, bearA, nodeA, dB, aB, bearB, nodeB, E, scale_factors, eligA=None, eligB=None,
                        residual_deg_thr=0.2, thr=50, check_orientation=True):
    dA = np.ascontiguousarray(dA, np.uint32); dB = np.ascontiguousarray(dB, np.uint32)
    aA = np.ascontiguousarray(aA, np.float32); aB = np.ascontiguousarray(aB, np.float32)
    octA = np.ascontiguousarray(octA, np.int32)
    bearA = np.ascontiguousarray(bearA, np.float64); bearB = np.ascontiguousarray(bearB, np.float64)
    nodeA = np.ascontiguousarray(nodeA, np.int32); nodeB = np.ascontiguousarray(nodeB, np.int32)
    E = np.ascontiguousarray(E, np.float64).reshape(9); sf = np.ascontiguousarray(scale_factors, np.float32)
    eA = None if eligA is None else np.ascontiguousarray(eligA, np.uint8)
    eB = None if eligB is None else np.ascontiguousarray(eligB, np.uint8)
    m = np.empty(max(len(dA), 1), np.int32)
    L = lib()
    L.orc_match_triangulation.restype = C.c_uint
    vp = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
    n = L.orc_match_triangulation(vp(dA), vp(aA), vp(octA), vp(bearA), vp(nodeA), vp(eA), len(dA), vp(dB), vp(aB), vp(bearB),
                                  vp(nodeB), vp(eB), len(dB), vp(E), vp(sf), C.c_float(residual_deg_thr), C.c_uint(thr),
                                  int(check_orientation), vp(m))
    return int(n), m[:len(dA)]


def angle_invalid(deltas, ids):
    deltas = np.ascontiguousarray(deltas, np.float32)
    ids = np.ascontiguousarray(ids, np.int32)
    out = np.empty(len(ids) + 1, np.int32)
    n = lib().orc_angle_invalid(deltas.ctypes, ids.ctypes, len(ids), out.ctypes)
    return out[:n].copy()


def bin_order(sizes):
    sizes = np.ascontiguousarray(sizes, np.uint32)
    out = np.empty(30, np.uint32)
    lib().orc_bin_order(sizes.ctypes, out.ctypes)
    return out


def bin_order_heap(sizes):
    sizes = np.ascontiguousarray(sizes, np.uint32)
    out = np.empty(30, np.uint32)
    lib().orc_bin_order_heap(sizes.ctypes, out.ctypes)
    return out


def angle_bin(delta):
    return int(lib().orc_angle_bin(C.c_float(delta)))


def medoid(desc, offsets):
    desc = np.ascontiguousarray(desc, np.uint32).reshape(-1, 8)
    offsets = np.ascontiguousarray(offsets, np.int64)
    best = np.zeros(max(len(offsets) - 1, 1), np.int32)
    lib().orc_medoid(C.c_void_p(desc.ctypes.data), C.c_void_p(offsets.ctypes.data), len(offsets) - 1, C.c_void_p(best.ctypes.data))
    return best[:len(offsets) - 1]


def feature_index(x, y):
    x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
    order = np.zeros(max(len(x), 1), np.int32)
    lib().orc_feature_index(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), len(x), C.c_void_p(order.ctypes.data))
    return order[:len(x)]


def features_around(x, y, qx, qy, r):
    x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
    out = np.zeros(max(len(x), 1), np.int32)
    n = lib().orc_features_around(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), len(x), C.c_float(qx), C.c_float(qy),
                                  C.c_float(r), C.c_void_p(out.ctypes.data))
    return out[:n].copy()


def search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=0, thr=50, taken=None, qlevel=None):
    kx = np.ascontiguousarray(kx, np.float32); ky = np.ascontiguousarray(ky, np.float32)
    koct = np.ascontiguousarray(koct, np.int32); kdesc = np.ascontiguousarray(kdesc, np.uint32).reshape(-1, 8)
    qx = np.ascontiguousarray(qx, np.float32); qy = np.ascontiguousarray(qy, np.float32)
    qr = np.ascontiguousarray(qr, np.float32); qdesc = np.ascontiguousarray(qdesc, np.uint32).reshape(-1, 8)
    ql = None if qlevel is None else np.ascontiguousarray(qlevel, np.int32)
    tk = np.zeros(max(len(kx), 1), np.uint8) if taken is None else taken
    idx = np.zeros(max(len(qx), 1), np.int32)
    dist = np.zeros(max(len(qx), 1), np.uint32)
    L = lib()
    L.orc_search_candidates.restype = C.c_int
    n = L.orc_search_candidates(C.c_void_p(kx.ctypes.data), C.c_void_p(ky.ctypes.data), C.c_void_p(koct.ctypes.data),
                                C.c_void_p(kdesc.ctypes.data), len(kx), C.c_void_p(tk.ctypes.data),
                                C.c_void_p(qx.ctypes.data), C.c_void_p(qy.ctypes.data), C.c_void_p(qr.ctypes.data),
                                C.c_void_p(qdesc.ctypes.data), None if ql is None else C.c_void_p(ql.ctypes.data),
                                len(qx), int(mode), C.c_uint(thr), C.c_void_p(idx.ctypes.data), C.c_void_p(dist.ctypes.data))
    return int(n), idx[:len(qx)], dist[:len(qx)]


def match_sim3(x1, y1, oct1, d1, x2, y2, oct2, d2, q12, q12desc, q12lvl, q21, q21desc, q21lvl):
    """matchMapPointsSim3 (keyframe_matcher.cpp:633-686) on projected queries: q12 [n1, 3] = (x, y, r) of keyframe 1's map
    points in keyframe 2 (r < 0: no query), q21 [n2, 3] the reverse.  Returns the agreed (i, j) pairs [n, 2]."""
    f = lambda a: np.ascontiguousarray(a, np.float32)
    i = lambda a: np.ascontiguousarray(a, np.int32)
    u = lambda a: np.ascontiguousarray(a, np.uint32).reshape(-1, 8)
    x1, y1, x2, y2 = f(x1), f(y1), f(x2), f(y2)
    oct1, oct2, q12lvl, q21lvl = i(oct1), i(oct2), i(q12lvl), i(q21lvl)
    d1, d2, q12desc, q21desc = u(d1), u(d2), u(q12desc), u(q21desc)
    q12 = f(q12).reshape(-1, 3); q21 = f(q21).reshape(-1, 3)
    a = [np.ascontiguousarray(q12[:, k]) for k in range(3)] + [np.ascontiguousarray(q21[:, k]) for k in range(3)]
    out = np.zeros((max(min(len(x1), len(x2)), 1), 2), np.int32)
    vp = lambda v: C.c_void_p(v.ctypes.data)
    L = lib()
    L.orc_match_sim3.restype = C.c_int
    n = L.orc_match_sim3(vp(x1), vp(y1), vp(oct1), vp(d1), len(x1), vp(x2), vp(y2), vp(oct2), vp(d2), len(x2),
                         vp(a[0]), vp(a[1]), vp(a[2]), vp(q12desc), vp(q12lvl), vp(a[3]), vp(a[4]), vp(a[5]), vp(q21desc),
                         vp(q21lvl), vp(out))
    return out[:n].copy()


def bow_transform(vocab, desc, levels_up=4):
    """vocab: dict(child_off, child_ids, node_desc, node_weight, node_word, levels) (see synth.random_vocabulary)."""
    desc = np.ascontiguousarray(desc, np.uint32).reshape(-1, 8)
    n = len(desc)
    word = np.zeros(max(n, 1), np.int32); weight = np.zeros(max(n, 1), np.float64); node = np.zeros(max(n, 1), np.int32)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    nw = np.ascontiguousarray(vocab["node_weight"], np.float64)
    lib().orc_bow_transform(vp(vocab["child_off"]), vp(vocab["child_ids"]), vp(vocab["node_desc"]), vp(nw),
                            vp(vocab["node_word"]), len(vocab["node_word"]), int(vocab["levels"]), vp(desc), n, int(levels_up),
                            vp(word), vp(weight), vp(node))
    return word[:n], weight[:n], node[:n]

def bow_vector(word, weight):
    """DBoW2 BowVector of one keyframe (addWeight in feature order, L1 normalise) -> (words ascending, values)."""
    word = np.ascontiguousarray(word, np.int32); weight = np.ascontiguousarray(weight, np.float64)
    n = len(word)
    vw = np.zeros(max(n, 1), np.uint32); vv = np.zeros(max(n, 1), np.float64)
    L = lib()
    L.orc_bow_vector.restype = C.c_int
    k = L.orc_bow_vector(C.c_void_p(word.ctypes.data), C.c_void_p(weight.ctypes.data), n, C.c_void_p(vw.ctypes.data),
                         C.c_void_p(vv.ctypes.data))
    return vw[:k].copy(), vv[:k].copy()


class BowIndex:
    """Restatement of BowIndex (bow_index.cpp:31-57, 95-176): inverted lists + getBowSimilar."""

    def __init__(self, vocabulary_size):
        L = lib()
        L.orc_bowindex_create.restype = C.c_void_p
        L.orc_bowindex_similar.restype = C.c_int
        self._h = C.c_void_p(L.orc_bowindex_create(int(vocabulary_size)))
        self.n = 0

    def add(self, map_id, kf_id, vec_word, vec_value):
        vw = np.ascontiguousarray(vec_word, np.uint32); vv = np.ascontiguousarray(vec_value, np.float64)
        lib().orc_bowindex_add(self._h, int(map_id), int(kf_id), C.c_void_p(vw.ctypes.data), C.c_void_p(vv.ctypes.data), len(vw))
        self.n += 1

    def remove(self, map_id, kf_id):
        lib().orc_bowindex_remove(self._h, int(map_id), int(kf_id))

    def similar(self, vec_word, vec_value, self_key=(-1, -1), min_in_common_ratio=0.8, score_ratio=0.75):
        vw = np.ascontiguousarray(vec_word, np.uint32); vv = np.ascontiguousarray(vec_value, np.float64)
        cap = max(self.n, 1)
        om = np.zeros(cap, np.int32); ok = np.zeros(cap, np.int32); osc = np.zeros(cap, np.float32)
        k = lib().orc_bowindex_similar(self._h, C.c_void_p(vw.ctypes.data), C.c_void_p(vv.ctypes.data), len(vw), int(self_key[0]),
                                       int(self_key[1]), C.c_float(min_in_common_ratio), C.c_float(score_ratio),
                                       C.c_void_p(om.ctypes.data), C.c_void_p(ok.ctypes.data), C.c_void_p(osc.ctypes.data), cap)
        k = min(k, cap)
        return om[:k].copy(), ok[:k].copy(), osc[:k].copy()

    def close(self):
        if self._h:
            lib().orc_bowindex_destroy(self._h)
            self._h = None


def bench_extract(p, imgs, threads):
    imgs = np.ascontiguousarray(imgs, np.uint8)
    total = C.c_long(0)
    s = lib().orc_bench_extract(C.byref(p), imgs.ctypes, imgs.shape[0], int(threads), C.byref(total))
    return float(s), total.value


def bench_match(desc, ang, pairs, threads, ratio=0.8, thr=50):
    desc = np.ascontiguousarray(desc, np.uint32)
    ang = np.ascontiguousarray(ang, np.float32)
    pairs = np.ascontiguousarray(pairs, np.int32)
    total = C.c_long(0)
    s = lib().orc_bench_match(desc.ctypes, ang.ctypes, desc.shape[0], desc.shape[1], pairs.ctypes, len(pairs),
                              C.c_float(ratio), C.c_uint(thr), int(threads), C.byref(total))
    return float(s), total.value
