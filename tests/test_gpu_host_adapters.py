"""The C++ adapters (slam-module_b200/host/: ImagePyramid, FeatureDetector, OrbExtractor,
matchForLoopClosures with the reference's signatures) driven by tests/cpp/host_adapter_main.cpp the way
the reference's own callers drive the original classes; every dumped artefact is compared bit for bit
with the CPU oracle."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "slam-module_b200" / "host"


def _build():
    subprocess.check_call(["make", "-C", str(HOST), "all"], stdout=subprocess.DEVNULL)
    return HOST / "host_adapter_main"


def test_host_adapter_library_builds_and_exports():
    """CPU check: the adapter library builds with plain g++ and exports the reference's entry points."""
    _build()
    syms = subprocess.check_output(["nm", "-DC", str(HOST / "libslam_frontend.so")], text=True)
    for s in ("slam::ImagePyramid::build(", "slam::FeatureDetector::build(", "slam::OrbExtractor::build(",
              "slam::matchForLoopClosures(", "slam::matchForTriangulationDBoW(", "slam::searchByProjection(",
              "slam::replaceDuplication<std::vector<slam::MpId", "slam::replaceDuplication<std::set<slam::MpId",
              "slam::matchMapPointsSim3(", "slam::updateDescriptors(", "slam::FeatureSearch::create(",
              "slam::loadVocabularyText(", "slam::BowIndex::transform(", "slam::BowIndex::getBowSimilar(", "slam::StaticSettings::maxNumberOfKeypointsPerLevel()",
              "slam::match::compute_descriptor_distance_32("):
        assert s in syms, s


def _kps(prefix):
    xya = np.fromfile(str(prefix) + "_xya.f32", np.float32).reshape(-1, 3)
    return dict(x=xya[:, 0], y=xya[:, 1], angle=xya[:, 2], octave=np.fromfile(str(prefix) + "_octave.i32", np.int32),
                desc=np.fromfile(str(prefix) + "_desc.u32", np.uint32).reshape(-1, 8))


def _same(got, ref, keep=None):
    keep = np.ones(ref["n"], bool) if keep is None else keep
    assert len(got["x"]) == int(keep.sum())
    for k in ("x", "y", "angle", "octave", "desc"):
        assert np.array_equal(got[k], ref[k][keep]), k


@pytest.mark.gpu
def test_host_adapters_match_oracle(tmp_path, oracle, synth):
    exe = _build()
    w, h, maxkp = 640, 480, 1000
    a, b = synth.frame(w, h, 7100), synth.frame(w, h, 7101)
    a.tofile(tmp_path / "a.raw")
    b.tofile(tmp_path / "b.raw")
    voc = synth.random_vocabulary(4, 5, 21)                     # levelsUp = 4 of 5 levels: feature-vector nodes at level 1
    for name, key, t in (("child_off.i32", "child_off", np.int32), ("child_ids.i32", "child_ids", np.int32),
                         ("node_word.i32", "node_word", np.int32), ("node_desc.u32", "node_desc", np.uint32),
                         ("node_weight.f64", "node_weight", np.float64)):
        np.ascontiguousarray(voc[key], t).tofile(tmp_path / ("voc_" + name))
    np.array([voc["levels"]], np.int32).tofile(tmp_path / "voc_levels.i32")
    import refcases
    refcases.write_vocabulary_txt(voc, str(tmp_path / "voc.txt"), 4)     # the DBoW2 text format bow_index.cpp:11-19 loads
    r = subprocess.run([str(exe), str(w), str(h), str(tmp_path / "a.raw"), str(tmp_path / "b.raw"), str(tmp_path), str(maxkp)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "matchers ok" in r.stdout and "vocabulary text loader ok" in r.stdout, r.stdout
    p = oracle.make_params(w, h, max_keypoints=maxkp)
    _, _, _, budgets = oracle.geometry(p)

    # ImagePyramid::update / getLevel / getBlurredLevel
    lv, bl = oracle.pyramid(p, a)
    dims = np.fromfile(tmp_path / "dims.i32", np.int32).reshape(-1, 2)
    for l in range(8):
        assert tuple(dims[l]) == (lv[l].shape[1], lv[l].shape[0])
        assert np.array_equal(np.fromfile(tmp_path / ("pyr_%d.u8" % l), np.uint8).reshape(lv[l].shape), lv[l]), l
        assert np.array_equal(np.fromfile(tmp_path / ("blur_%d.u8" % l), np.uint8).reshape(bl[l].shape), bl[l]), l

    # FeatureDetector::detect: per-level keypoints in level coordinates, in the oracle's order
    det = np.fromfile(tmp_path / "detect.i32", np.int32).reshape(-1, 3)
    for l in range(8):
        ox, oy, _ = oracle.detect_level(lv[l], int(budgets[l]))
        d = det[det[:, 0] == l]
        assert d[:, 1].tolist() == ox.tolist() and d[:, 2].tolist() == oy.tolist(), l

    # OrbExtractor::detectAndExtract with tracker features (level 1) and a camera rejecting x >= w - 40
    tracks = np.array([[(7 + (i * 37) % (w - 3)) + 0.25 * (i % 4), (5 + (i * 53) % (h - 3)) + 0.5 * (i % 2)] for i in range(40)],
                      np.float32)
    ids = (100 + np.arange(40)).astype(np.int32)
    ok = tracks[:, 0].astype(np.float64) < w - 40.0           # camera.isValidPixel on the tracker points
    ref = oracle.extract(p, a, tracks=tracks[ok], track_ids=ids[ok], track_level=1)
    keep = (ref["track_id"] >= 0) | (ref["x"].astype(np.float64) < w - 40.0)   # dropInvalidKeypoints on detected ones
    got = _kps(tmp_path / "kpsA")
    _same(got, ref, keep)
    assert np.array_equal(np.fromfile(tmp_path / "idsA.i32", np.int32), ref["track_id"][keep])
    assert (ref["track_id"] >= 0).sum() > 10 and (~keep).sum() > 0

    refB = oracle.extract(p, b)
    gotB = _kps(tmp_path / "kpsB")
    _same(gotB, refB)
    refA = oracle.extract(p, a)
    _same(_kps(tmp_path / "batch0"), refB)
    _same(_kps(tmp_path / "batch1"), refA)

    # matchForLoopClosures with map-point filters == brute force on the eligible subsets
    n1, n2 = refA["n"], refB["n"]
    i1 = np.array([i for i in range(n1) if i % 5 != 4 and i % 7 != 6])
    i2 = np.array([i for i in range(n2) if i % 3 != 2])
    n, m = oracle.match_bruteforce(refA["desc"][i1], refA["angle"][i1], refB["desc"][i2], refB["angle"][i2])
    want = np.full(n1, -1, np.int32)
    want[i1[m >= 0]] = i2[m[m >= 0]]
    lm = np.fromfile(tmp_path / "loop_matches.i32", np.int32)
    assert lm[-1] == n and np.array_equal(lm[:-1], want)
    assert n > 0

    # ... and with bowFeatureVec: the reference's node-bucketed comparison
    nodeA = ((refA["desc"][:, 0] ^ refA["desc"][:, 3]) % np.uint32(7)).astype(np.int32)
    nodeB = ((refB["desc"][:, 0] ^ refB["desc"][:, 3]) % np.uint32(7)).astype(np.int32)
    eA = np.zeros(n1, np.uint8); eA[i1] = 1
    eB = np.zeros(n2, np.uint8); eB[i2] = 1
    nw, mw = oracle.match_bow(refA["desc"], refA["angle"], nodeA, refB["desc"], refB["angle"], nodeB, eA, eB)
    lw = np.fromfile(tmp_path / "loop_matches_bow.i32", np.int32)
    assert lw[-1] == nw and np.array_equal(lw[:-1], mw)
    assert nw > 0

    nb, mb = oracle.match_bruteforce(refA["desc"], refA["angle"], refB["desc"], refB["angle"])
    bf = np.fromfile(tmp_path / "bf_matches.i32", np.int32)
    assert bf[-1] == nb and np.array_equal(bf[:-1], mb)

    k = min(len(got["x"]), n2)
    hd = np.fromfile(tmp_path / "hamming.u32", np.uint32)
    assert hd.tolist() == [oracle.hamming(got["desc"][i], refB["desc"][i]) for i in range(k)]

    # BowIndex adapter: transform -> BowVector / FeatureVector, add / remove, getBowSimilar (bow_index.cpp:44-176)
    dA, dB = refA["desc"], refB["desc"]
    sets = [dA, dB, dB, dA[:n1 // 2], dB[n2 // 3:], dA[n1 // 4:], dB[:50], dA[:0]]
    idx = oracle.BowIndex(int(voc["node_word"].max()) + 1)
    vecs = []
    for i, d in enumerate(sets):
        word, weight, node = oracle.bow_transform(voc, d, 4)
        vecs.append(oracle.bow_vector(word, weight))
        if i != 5:                                              # keyframe 6 is removed again by the driver
            idx.add(i % 2, i + 1, *vecs[-1])
        if i == 0:
            assert np.array_equal(np.fromfile(tmp_path / "bow_vec_word.u32", np.uint32), vecs[0][0])
            assert np.array_equal(np.fromfile(tmp_path / "bow_vec_value.f64", np.float64), vecs[0][1])
            keep = weight > 0
            order = np.lexsort((np.arange(len(d))[keep], node[keep]))     # std::map node order, features ascending
            assert np.array_equal(np.fromfile(tmp_path / "bow_fv_node.u32", np.uint32), node[keep][order].astype(np.uint32))
            assert np.array_equal(np.fromfile(tmp_path / "bow_fv_feature.u32", np.uint32), np.arange(len(d))[keep][order].astype(np.uint32))
            assert len(vecs[0][0]) > 50
    ids = np.fromfile(tmp_path / "bow_similar_ids.i32", np.int32).tolist()
    scores = np.fromfile(tmp_path / "bow_similar_scores.f32", np.float32).tolist()
    for q in (0, 1, 4):
        m, k, sc = idx.similar(*vecs[q], self_key=(0, q + 1), min_in_common_ratio=0.8, score_ratio=0.75)
        cnt = ids.pop(0)
        assert cnt == len(m)
        for j in range(cnt):
            assert (ids.pop(0), ids.pop(0)) == (int(m[j]), int(k[j]))
            assert scores.pop(0) == float(sc[j])
    assert not ids and not scores
    idx.close()



@pytest.mark.gpu
def test_matcher_adapters_on_reference_classes():
    """slam_matchers.hpp instantiated with the REFERENCE'S OWN Keyframe / MapPoint / MapDB classes, run next to the
    reference's own searchByProjection / replaceDuplication / matchMapPointsSim3 / matchForTriangulationDBoW /
    MapPoint::updateDescriptor on identical scenes (tests/cpp/ref_adapter_main.cpp, built by `make -C oracle ref_adapter`
    where /root/reference exists; the binary travels to the GPU box under oracle/_ref/)."""
    exe = ROOT / "oracle" / "_ref" / "ref_adapter_main"
    if not exe.exists():
        pytest.skip("oracle/_ref/ref_adapter_main was not built (no reference tree at build time)")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout, r.stdout
