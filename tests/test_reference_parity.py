"""The oracle restatement against THE REFERENCE'S OWN SOURCES compiled verbatim (oracle/_ref/libref_slam.so: static_settings,
feature_search, orb_extractor, image_pyramid, feature_detector, keyframe_matcher, map_point, bow_index .cpp; see
oracle/ref_slam.cpp for what is shimmed).  Two legs per case (tests/refcases.py):
  * golden: oracle == tests/golden/golden_ref_slam.npz (written from the reference by tools/gen_golden.py) -- always runs;
  * live:   oracle == the reference library, when oracle/_ref is prebuilt or /root/reference is there to build it.
"""
from functools import partial
from pathlib import Path

import numpy as np
import pytest

import refcases as rc

GOLDEN = Path(__file__).resolve().parent / "golden" / "golden_ref_slam.npz"

# name -> (case function taking backend [, golden]), needs the reference's recorded queries?
CASES = {
    "settings": (rc.case_settings, False),
    "feature_search": (rc.case_feature_search, False),
    "pyramid_crc": (rc.case_pyramid_crc, False),
    "extract_vga2000": (partial(rc.case_extract, name="vga2000"), False),
    "extract_vga1000": (partial(rc.case_extract, name="vga1000"), False),
    "extract_odd": (partial(rc.case_extract, name="odd"), False),
    "extract_tracks": (partial(rc.case_extract, name="tracks"), False),
    "loop_closures_a": (partial(rc.case_loop_closures, seed=3, require=True), False),
    "loop_closures_b": (partial(rc.case_loop_closures, seed=4, require=False), False),
    "loop_closures_bf": (partial(rc.case_loop_closures_bruteforce, seed=5), False),
    "triangulation_a": (partial(rc.case_triangulation, seed=6, thr_deg=0.2), False),
    "triangulation_b": (partial(rc.case_triangulation, seed=7, thr_deg=1.0), False),
    "search_by_projection": (partial(rc.case_search_by_projection, seed=8), True),
    "replace_duplication": (partial(rc.case_replace_duplication, seed=9), True),
    "sim3": (partial(rc.case_sim3, seed=10), True),
    "medoid": (rc.case_medoid, False),
}


def _assert_same(got, want, name):
    assert set(got) == set(want), name
    for k in want:
        a, b = np.asarray(got[k]), np.asarray(want[k])
        assert a.shape == b.shape, (name, k, a.shape, b.shape)
        assert np.array_equal(a, b), (name, k, int((a != b).sum()))


@pytest.fixture(scope="module")
def golden_ref():
    g = np.load(GOLDEN)
    out = {}
    for key in g.files:
        case, arr = key.split("/", 1)
        out.setdefault(case, {})[arr] = g[key]
    return out


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name, golden_ref):
    fn, needs_queries = CASES[name]
    want = golden_ref[name]
    got = fn("oracle", golden=want) if needs_queries else fn("oracle")
    _assert_same(got, want, name)
    # the cases must exercise something
    if "matches" in want:
        assert (want["matches"] >= 0).sum() > 20, name
    if name == "sim3":
        assert len(want["pairs"]) > 50
    if name == "search_by_projection":
        assert (want["idx"] >= 0).sum() > 50 and (want["qr"] < 0).sum() > 5
    if name == "replace_duplication":
        assert int(want["n"][0]) > 50 and (want["final"] >= 0).sum() > 30


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_live(name):
    if not rc.pr.available():
        pytest.skip("oracle/_ref/libref_slam.so not built and no /root/reference to build it from")
    fn, needs_queries = CASES[name]
    want = fn("ref")
    got = fn("oracle", golden=want) if needs_queries else fn("oracle")
    _assert_same(got, want, name)


def test_bow_index_matches_reference(tmp_path, golden_ref):
    got = rc.case_bow("oracle", tmp_path)
    _assert_same(got, golden_ref["bow"], "bow")
    assert len(got["sim0_kf"]) >= 1 and (got["node"] >= 0).sum() > 1000
    if rc.pr.available():
        _assert_same(got, rc.case_bow("ref", tmp_path), "bow live")


def test_reference_live_baseline_sized():
    """BASELINE configs 1-3 at full size, live against the reference library: 1000- and 2000-keypoint VGA frames, a 720p
    frame, and 2000 x 2000 brute-force matching."""
    if not rc.pr.available():
        pytest.skip("reference library not available")
    import slam_module_b200 as sm
    po, pr = rc.po, rc.pr
    for (w, h, mk, seed) in [(640, 480, 1000, 2000), (640, 480, 2000, 2001), (1280, 720, 2000, 4000)]:
        p = po.make_params(w, h, max_keypoints=mk)
        img = sm.synth.frame(w, h, seed)
        a, b = po.extract(p, img), pr.extract(p, img)
        assert a["n"] == b["n"] and a["n"] >= mk * 0.9
        for k in ("x", "y", "angle", "octave", "desc", "track_id"):
            assert np.array_equal(a[k], b[k]), (w, h, mk, k)
    for name, img in sm.synth.degenerate_frames(640, 480).items():
        p = po.make_params(640, 480, max_keypoints=1000)
        a, b = po.extract(p, img), pr.extract(p, img)
        assert a["n"] == b["n"], name
        assert np.array_equal(a["desc"], b["desc"]) and np.array_equal(a["angle"], b["angle"]), name
    for seed in (11, 12):
        _assert_same(rc.case_loop_closures_bruteforce("oracle", seed), rc.case_loop_closures_bruteforce("ref", seed), "bf%d" % seed)


def test_reference_cropped_camera_live():
    """tracker::Camera::isValidPixel with a cropped valid region (orb_extractor.cpp:101, 221-237): the reference drops
    keypoints whose full-resolution position is invalid; the oracle models the all-valid camera only, so the reference run
    must equal the oracle's output filtered by the same rectangle."""
    if not rc.pr.available():
        pytest.skip("reference library not available")
    import slam_module_b200 as sm
    po, pr = rc.po, rc.pr
    p = po.make_params(640, 480, max_keypoints=1000)
    img = sm.synth.frame(640, 480, 77)
    rect = (60.0, 40.0, 600.5, 431.25)
    b = pr.extract(p, img, valid_rect=rect)
    a = po.extract(p, img)
    keep = (a["x"] >= np.float32(rect[0])) & (a["x"] < rect[2]) & (a["y"] >= np.float32(rect[1])) & (a["y"] < rect[3])
    assert 0 < keep.sum() < a["n"] and b["n"] == int(keep.sum())
    for k in ("x", "y", "angle", "octave", "desc"):
        assert np.array_equal(a[k][keep], b[k]), k
