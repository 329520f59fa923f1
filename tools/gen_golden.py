#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ (run in the BUILD container only).

Sources of truth:
  * cv2 (opencv-python-headless; the library the reference calls for resize / GaussianBlur /
    fastAtan2, and whose FAST the upstream detector uses)  -> golden_cv2.npz
  * the reference's own header-only leaves compiled verbatim from /root/reference
    (oracle/_ref/libref_leaves.so: Hamming distance, util::cos/sin, rBRIEF pattern, angle checker)
    -> golden_ref_leaves.npz
Neither cv2's C++ headers nor /root/reference exist on the GPU box, so the outputs are committed.

    python tools/gen_golden.py
"""
import ctypes as C
import sys
import zlib
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import slam_module_b200 as sm  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)


def level_sizes(w, h, levels, factor):
    s = np.float32(1.0)
    out = [(w, h)]
    for _ in range(1, levels):
        s = np.float32(factor) * s
        out.append((int(np.round(w / np.float64(s))), int(np.round(h / np.float64(s)))))
    return out


def cv2_pyramid(img, levels, factor):
    lv, bl = [], []
    cur = img
    for l, (w, h) in enumerate(level_sizes(img.shape[1], img.shape[0], levels, factor)):
        if l > 0:
            cur = cv2.resize(cur, (w, h), interpolation=cv2.INTER_LINEAR)
        lv.append(cur)
        bl.append(cv2.GaussianBlur(cur, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
    return lv, bl


def main():
    g = {"cv2_version": np.array(cv2.__version__)}
    # --- pyramid: full planes of a small image, CRCs of the 640x480 and 1280x720 synthetic frames
    small = sm.synth.frame(160, 120, 4242)
    g["pyr_small_img"] = small
    lv, bl = cv2_pyramid(small, 4, 1.2)
    for l in range(4):
        g["pyr_small_l%d" % l] = lv[l]
        g["pyr_small_b%d" % l] = bl[l]
    for name, (w, h, seed) in {"vga": (640, 480, 1000), "hd": (1280, 720, 4000)}.items():
        img = sm.synth.frame(w, h, seed)
        lv, bl = cv2_pyramid(img, 8, 1.2)
        g["pyr_%s_crc" % name] = np.array([[zlib.crc32(a.tobytes()) for a in lv],
                                          [zlib.crc32(a.tobytes()) for a in bl]], np.uint32)
        g["pyr_%s_sizes" % name] = np.array([a.shape[::-1] for a in lv], np.int32)
    # odd sizes, upscale-free shapes incl. the exact-2x INTER_AREA switch
    rng = np.random.default_rng(7)
    for i, (sw, sh, dw, dh) in enumerate([(101, 77, 84, 64), (64, 64, 32, 32), (97, 53, 96, 52), (200, 31, 167, 26)]):
        src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        g["resize%d_src" % i] = src
        g["resize%d_dst" % i] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    for i, (w, h) in enumerate([(9, 8), (33, 17), (70, 70)]):
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        g["blur%d_src" % i] = src
        g["blur%d_dst" % i] = cv2.GaussianBlur(src, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    # --- FAST-9/16 with NMS
    fimg = sm.synth.frame(128, 96, 99)
    g["fast_img"] = fimg
    for thr in (20, 7):
        k = cv2.FastFeatureDetector_create(thr, True).detect(fimg)
        g["fast_kp_thr%d" % thr] = np.array([(int(q.pt[0]), int(q.pt[1]), int(q.response)) for q in k], np.int32).reshape(-1, 3)
    # --- fastAtan2 (scalar path)
    yx = rng.integers(-1300000, 1300000, (4000, 2)).astype(np.float32)
    yx[:8] = [[0, 0], [0, 1], [1, 0], [0, -1], [-1, 0], [1, 1], [-1, -1], [5, -5]]
    g["atan2_yx"] = yx
    g["atan2_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(OUT / "golden_cv2.npz", **g)

    # --- reference leaves
    R = po.ref_lib()
    assert R is not None, "oracle/_ref not built (needs /root/reference)"
    r = {}
    a = rng.integers(0, 2 ** 32, (512, 8), dtype=np.uint32)
    b = rng.integers(0, 2 ** 32, (512, 8), dtype=np.uint32)
    b[:64] = a[:64] ^ (np.uint32(1) << rng.integers(0, 32, (64, 8)).astype(np.uint32))
    r["hamm_a"], r["hamm_b"] = a, b
    r["hamm_d"] = np.array([R.ref_hamming(a[i].ctypes, b[i].ctypes) for i in range(len(a))], np.uint32)
    v = np.concatenate([np.linspace(-20, 20, 4001), rng.uniform(-7, 7, 2000)]).astype(np.float32)
    r["trig_v"] = v
    r["trig_cos"] = np.array([R.ref_cos(float(x)) for x in v], np.float32)
    r["trig_sin"] = np.array([R.ref_sin(float(x)) for x in v], np.float32)
    pat = np.empty(1024, np.float32)
    R.ref_pattern(pat.ctypes)
    r["pattern"] = pat
    r["thr"] = np.array([R.ref_thr_low(), R.ref_thr_high(), R.ref_max_dist()], np.uint32)
    # angle checker: several delta sets with many ties between bins
    sets = []
    for n in (0, 1, 3, 7, 20, 60, 300):
        for rep in range(4):
            centers = rng.uniform(-360, 360, 4)
            d = (centers[rng.integers(0, 4, n)] + rng.normal(0, 8, n)).astype(np.float32)
            d = np.clip(d, -359.9, 359.9).astype(np.float32)
            ids = np.arange(n, dtype=np.int32)
            inv = np.empty(n + 1, np.int32)
            k = R.ref_angle_invalid(d.ctypes, ids.ctypes, n, inv.ctypes)
            sets.append((d, inv[:k].copy()))
    r["angle_n_sets"] = np.array(len(sets))
    for i, (d, inv) in enumerate(sets):
        r["angle_d%d" % i] = d
        r["angle_inv%d" % i] = inv
    np.savez_compressed(OUT / "golden_ref_leaves.npz", **r)
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
