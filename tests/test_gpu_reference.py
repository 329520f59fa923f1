"""libslamgpu.so (through its C ABI) against THE REFERENCE ITSELF: the outputs of the reference's own sources compiled
verbatim (oracle/_ref/libref_slam.so), taken from tests/golden/golden_ref_slam.npz and -- when the prebuilt library
travelled to the GPU box -- recomputed live.  Same seeded cases as tests/test_reference_parity.py (tests/refcases.py)."""
from pathlib import Path

import numpy as np
import pytest

import refcases as rc
from test_reference_parity import CASES, GOLDEN, _assert_same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(slamgpu):
    g = rc.GpuImpl(slamgpu)
    yield g
    g.close()


@pytest.fixture(scope="module")
def golden_ref():
    g = np.load(GOLDEN)
    out = {}
    for key in g.files:
        case, arr = key.split("/", 1)
        out.setdefault(case, {})[arr] = g[key]
    return out


def _trim(want, got):
    """Keys the C ABI does not produce (FeatureSearch queries: the index order is what the ABI exposes)."""
    return {k: want[k] for k in got}


@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_reference(name, gpu, golden_ref):
    fn, needs_queries = CASES[name]
    want = golden_ref[name]
    got = fn(gpu, golden=want) if needs_queries else fn(gpu)
    _assert_same(got, _trim(want, got), name)
    if rc.pr.available():          # prebuilt oracle/_ref on the box: the same comparison against a live run
        live = fn("ref")
        _assert_same(got, _trim(live, got), name + " (live)")


def test_gpu_bow_matches_reference(gpu, golden_ref, tmp_path):
    got = rc.case_bow(gpu, tmp_path)
    _assert_same(got, golden_ref["bow"], "bow")


def test_gpu_sim3_edge_cases(gpu):
    """No queries at all, every query dead (r < 0), and one-sided agreement."""
    rng = np.random.default_rng(3)
    n = 50
    x = rng.uniform(0, 100, n).astype(np.float32); y = rng.uniform(0, 100, n).astype(np.float32)
    octv = np.zeros(n, np.int32); d = rng.integers(0, 2 ** 32, (n, 8), dtype=np.uint32)
    dead = np.stack([x, y, np.full(n, -1, np.float32)], axis=1)
    live = np.stack([x, y, np.full(n, 3, np.float32)], axis=1)
    lvl = np.zeros(n, np.int32)
    assert len(gpu.match_sim3(x, y, octv, d, x, y, octv, d, dead, d, lvl, dead, d, lvl)) == 0
    assert len(gpu.match_sim3(x, y, octv, d, x, y, octv, d, live, d, lvl, dead, d, lvl)) == 0      # 2 -> 1 never answers
    both = gpu.match_sim3(x, y, octv, d, x, y, octv, d, live, d, lvl, live, d, lvl)
    want = rc.po.match_sim3(x, y, octv, d, x, y, octv, d, live, d, lvl, live, d, lvl)
    assert np.array_equal(both, want) and len(both) == n
