"""Host-side multi-GPU logic on CPU: block partitioning of frames / pair lists and the final
host gather over torch.distributed with the gloo backend, world_size 2.  The per-rank compute is the CPU
oracle here (test infrastructure); on the GPU box the same partition feeds one sg_ctx per rank."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_block_range_partitions_exactly():
    from slam_module_b200 import sharding as sh
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            blocks = [sh.block_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
            rr = np.sort(np.concatenate([sh.round_robin(n, r, world) for r in range(world)]))
            assert np.array_equal(rr, np.arange(n))
    with pytest.raises(ValueError):
        sh.block_range(10, 2, 2)


def test_unordered_pair_enumeration():
    from slam_module_b200 import sharding as sh
    for n in (2, 3, 17, 200):
        want = np.array([(i, j) for i in range(n) for j in range(i + 1, n)], np.int64)
        i, j = sh.unordered_pair(np.arange(sh.n_unordered_pairs(n)), n)
        assert np.array_equal(np.stack([i, j], 1), want)
        got = np.concatenate([sh.pair_block(n, r, 3) for r in range(3)])
        assert np.array_equal(got, want)
    # config 5 scale: 10 000 keyframes -> 49 995 000 pairs; spot-check both ends and block seams
    n = 10000
    total = sh.n_unordered_pairs(n)
    assert total == 49995000
    k = np.array([0, 1, 9998, 9999, total // 2, total - 2, total - 1])
    i, j = sh.unordered_pair(k, n)
    assert (i < j).all() and (j < n).all()
    assert np.array_equal(i * n - i * (i + 1) // 2 + (j - i - 1), k)
    assert sh.pair_block(n, 7, 8, limit=5).shape == (5, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as td
    sys.path.insert(0, str(ROOT))
    import slam_module_b200 as sm
    from slam_module_b200 import sharding as sh
    from oracle import pyoracle as po
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # matching: 5 descriptor sets, all 10 unordered pairs block-sharded over the ranks
        d, a = sm.synth.random_descriptors(5, 120, 3)
        d[1, :60] = d[0, :60]          # make some pairs actually match
        a[1, :60] = a[0, :60]
        pairs = sh.pair_block(5, rank, world)
        counts = np.array([po.match_bruteforce(d[i], a[i], d[j], a[j])[0] for i, j in pairs], np.int64)
        all_pairs = sh.host_gather(pairs, td)
        all_counts = sh.host_gather(counts, td)
        # extraction: 3 frames, block-sharded
        lo, hi = sh.block_range(3, rank, world)
        p = po.make_params(160, 120, levels=3, max_keypoints=100)
        n_kp = np.array([po.extract(p, sm.synth.frame(160, 120, 40 + f))["n"] for f in range(lo, hi)], np.int64)
        all_kp = sh.host_gather(n_kp, td)
        t = sh.max_over_ranks(1.0 + rank, td)
        td.barrier()
        if rank == 0:
            q.put((all_pairs, all_counts, all_kp, t))
    finally:
        td.destroy_process_group()


def test_world_size_2_gloo_gather_matches_single_process(oracle, synth):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    pairs, counts, n_kp, t = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from slam_module_b200 import sharding as sh
    d, a = synth.random_descriptors(5, 120, 3)
    d[1, :60] = d[0, :60]
    a[1, :60] = a[0, :60]
    want_pairs = sh.pair_block(5, 0, 1)
    want = np.array([oracle.match_bruteforce(d[i], a[i], d[j], a[j])[0] for i, j in want_pairs], np.int64)
    assert np.array_equal(pairs, want_pairs) and np.array_equal(counts, want)
    assert counts[0] >= 40                       # sets 0 and 1 share 60 descriptors
    p = oracle.make_params(160, 120, levels=3, max_keypoints=100)
    want_kp = [oracle.extract(p, synth.frame(160, 120, 40 + f))["n"] for f in range(3)]
    assert n_kp.tolist() == want_kp
    assert t == 2.0                              # max over ranks
