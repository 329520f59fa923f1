"""Profiling target: 256-frame extraction steps as SINGLE full-batch launches on one stream (no slicing), then -- with
`match` as the first argument -- 2048-pair matching steps.  3 warm-up steps, then 2 steps for ncu to capture (`-s` skips
the warm-up launches of the kernel in question: 3 per single-launch kernel, 3 / 21 for the two pyramid kernels)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
import slam_module_b200 as sm
from slam_module_b200 import slamgpu

W, H, FRAMES = bench.W, bench.H, bench.FRAMES
ctx = slamgpu.Context(W, H, max_keypoints=bench.MAXKP, max_frames=FRAMES)
ctx.set_overlap(1)
batches = [bench.make_frames(FRAMES, 10000 + 100 * b) for b in range(2)]
bufs = [ctx.device_buffer(FRAMES * W * H).upload(b) for b in batches]
for i in range(5):
    ctx.extract_device(bufs[i % 2].ptr, W, W * H, FRAMES)
ctx.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "match":
    d, a = sm.synth.random_descriptors(bench.MATCH_SETS, bench.MATCH_N, 5)
    db = slamgpu.DescriptorDB(ctx, d, a)
    pairs = np.random.default_rng(1).integers(0, bench.MATCH_SETS, (bench.MATCH_PAIRS, 2)).astype(np.int32)
    dp = ctx.device_buffer(pairs.nbytes).upload(pairs)
    dc = ctx.device_buffer(4 * bench.MATCH_PAIRS)
    for _ in range(5):
        db.match_pairs_device(dp.ptr, bench.MATCH_PAIRS, dc.ptr)
    ctx.synchronize()
print("done")
