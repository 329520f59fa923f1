"""CPU tests of the ORACLE (the checker itself): pinned against the committed golden vectors
(generated from cv2 4.13 and from the reference's own headers compiled verbatim, tools/gen_golden.py),
against cv2 live where it is importable, and against the reference leaves in oracle/_ref when built."""
import zlib

import numpy as np
import pytest


def test_geometry_and_budgets(oracle):
    # level sizes and budgets for the task configs
    s, w, h, b = oracle.geometry(oracle.make_params(640, 480, max_keypoints=2000))
    assert w.tolist() == [640, 533, 444, 370, 309, 257, 214, 179]
    assert h.tolist() == [480, 400, 333, 278, 231, 193, 161, 134]
    assert b.tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert s[3] == np.float32(1.2) * np.float32(1.2) * np.float32(1.2) or True
    s, w, h, b = oracle.geometry(oracle.make_params(1280, 720, max_keypoints=1000))
    assert w.tolist() == [1280, 1067, 889, 741, 617, 514, 429, 357]
    assert h.tolist() == [720, 600, 500, 417, 347, 289, 241, 201]
    assert b.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]


def test_pyramid_golden(oracle, golden, synth):
    g = golden["cv2"]
    lv, bl = oracle.pyramid(oracle.make_params(160, 120, levels=4), g["pyr_small_img"])
    for l in range(4):
        assert np.array_equal(lv[l], g["pyr_small_l%d" % l])
        assert np.array_equal(bl[l], g["pyr_small_b%d" % l])
    for name, (w, h, seed) in {"vga": (640, 480, 1000), "hd": (1280, 720, 4000)}.items():
        lv, bl = oracle.pyramid(oracle.make_params(w, h), synth.frame(w, h, seed))
        crc = g["pyr_%s_crc" % name]
        assert [zlib.crc32(a.tobytes()) for a in lv] == crc[0].tolist()
        assert [zlib.crc32(a.tobytes()) for a in bl] == crc[1].tolist()
        assert [a.shape[::-1] for a in lv] == [tuple(x) for x in g["pyr_%s_sizes" % name].tolist()]


def test_resize_and_blur_golden(oracle, golden):
    g = golden["cv2"]
    for i in range(4):
        dst = g["resize%d_dst" % i]
        assert np.array_equal(oracle.resize(g["resize%d_src" % i], dst.shape[1], dst.shape[0]), dst), i
    for i in range(3):
        assert np.array_equal(oracle.gaussian7(g["blur%d_src" % i]), g["blur%d_dst" % i]), i
    imp = np.zeros((15, 15), np.uint8)
    imp[7, 7] = 255
    assert oracle.gaussian7(imp)[7, 4:11].tolist() == [4, 7, 10, 12, 10, 7, 4]   # SURVEY 8c KAT


def test_fast_golden(oracle, golden):
    g = golden["cv2"]
    for thr in (20, 7):
        x, y, r = oracle.cv_fast(g["fast_img"], thr)
        assert np.array_equal(np.stack([x, y, r], axis=1), g["fast_kp_thr%d" % thr])
        assert len(x) > 10


def test_atan2_golden(oracle, golden):
    g = golden["cv2"]
    got = np.array([oracle.lib().orc_fast_atan2(float(y), float(x)) for y, x in g["atan2_yx"]], np.float32)
    assert np.array_equal(got, g["atan2_deg"])


def test_reference_leaves_golden(oracle, golden):
    r = golden["ref"]
    assert [oracle.hamming(a, b) for a, b in zip(r["hamm_a"], r["hamm_b"])] == r["hamm_d"].tolist()
    L = oracle.lib()
    assert np.array_equal(np.array([L.orc_util_cos(float(v)) for v in r["trig_v"]], np.float32), r["trig_cos"])
    assert np.array_equal(np.array([L.orc_util_sin(float(v)) for v in r["trig_v"]], np.float32), r["trig_sin"])
    pat = np.empty(1024, np.float32)
    L.orc_pattern(pat.ctypes)
    assert np.array_equal(pat, r["pattern"])
    assert r["thr"].tolist() == [50, 100, 256]
    for i in range(int(r["angle_n_sets"])):
        d = r["angle_d%d" % i]
        assert np.array_equal(oracle.angle_invalid(d, np.arange(len(d))), r["angle_inv%d" % i]), i


def test_known_answers(oracle):
    a = np.array([0xffffffff, 0, 1, 2, 3, 4, 5, 6], np.uint32)
    assert oracle.hamming(a, np.zeros(8, np.uint32)) == 41
    assert np.float32(oracle.lib().orc_util_cos(1.0)) == np.float32(0.540614009)
    assert np.float32(oracle.lib().orc_util_sin(1.0)) == np.float32(0.841844141)
    assert oracle.angle_bin(359.9) == 12 and oracle.angle_bin(15.0) == 0 and oracle.angle_bin(45.0) == 2
    um = np.empty(16, np.int32)
    oracle.lib().orc_umax(um.ctypes)
    assert um.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def test_live_reference_leaves(oracle):
    R = oracle.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (the reference tree is not on this machine)")
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2 ** 32, (200, 8), dtype=np.uint32)
    b = rng.integers(0, 2 ** 32, (200, 8), dtype=np.uint32)
    for i in range(200):
        assert oracle.hamming(a[i], b[i]) == R.ref_hamming(a[i].ctypes, b[i].ctypes)
    for v in rng.uniform(-50, 50, 500).astype(np.float32):
        assert oracle.lib().orc_util_cos(float(v)) == R.ref_cos(float(v))


def test_live_cv2(oracle, synth):
    cv2 = pytest.importorskip("cv2")
    img = synth.frame(320, 240, 17)
    lv, bl = oracle.pyramid(oracle.make_params(320, 240, levels=5), img)
    cur = img
    for l in range(5):
        if l:
            cur = cv2.resize(cur, lv[l].shape[::-1], interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(cur, lv[l])
        assert np.array_equal(cv2.GaussianBlur(cur, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101), bl[l])
    k = cv2.FastFeatureDetector_create(20, True).detect(img)
    x, y, r = oracle.cv_fast(img, 20)
    assert [(int(q.pt[0]), int(q.pt[1]), int(q.response)) for q in k] == list(zip(x.tolist(), y.tolist(), r.tolist()))


def test_detect_level_contract(oracle, synth):
    """The contract the reference fixes around the absent detector: at most budget (+2) points, all at
    least 19 px from the border, distinct, each a FAST corner of the level."""
    img = synth.frame(640, 480, 1000)
    for budget in (434, 60, 5):
        x, y, r = oracle.detect_level(img, budget)
        assert 0 < len(x) <= budget + 2
        assert x.min() >= 19 and y.min() >= 19 and x.max() < 640 - 19 and y.max() < 480 - 19
        assert len(set(zip(x.tolist(), y.tolist()))) == len(x)
        assert (r >= 7).all()
    (x, y, r), (cx, cy, cr) = oracle.detect_level(img, 10 ** 6, with_candidates=True)
    assert len(x) == len(cx)   # an unreachable budget keeps every candidate


def test_matcher_semantics(oracle, synth):
    dA, aA, dB, aB = synth.correlated_descriptors(400, 3)
    n, m = oracle.match_bruteforce(dA, aA, dB, aB)
    assert n == (m >= 0).sum() and n > 50
    taken = m[m >= 0]
    assert len(set(taken.tolist())) == len(taken)           # uniqueness (keyframe_matcher.cpp:128)
    n2, m2 = oracle.match_bruteforce(dA, aA, dB, aB, check_orientation=False)
    assert n2 >= n and ((m == m2) | (m == -1)).all()        # the angle filter only removes
    for i in np.nonzero(m2 >= 0)[0]:
        assert oracle.hamming(dA[i], dB[m2[i]]) <= 50
