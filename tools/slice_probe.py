"""Device-resident extraction step time against the number of batch slices and compute streams (sg_set_overlap)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from slam_module_b200 import slamgpu

W, H, FRAMES = 640, 480, 256
ctx = slamgpu.Context(W, H, max_keypoints=2000, max_frames=FRAMES)
batches = [bench.make_frames(FRAMES, 10000 + 100 * b) for b in range(4)]
bufs = [ctx.device_buffer(FRAMES * W * H).upload(b) for b in batches]
for streams, parts in [(4, 4), (2, 2), (3, 3), (2, 4), (4, 8), (6, 6), (8, 8), (1, 1), (1, 4), (3, 6), (5, 5)]:
    ctx.set_overlap(streams * 1000 + parts)
    for i in range(5):
        ctx.extract_device(bufs[i % 4].ptr, W, W * H, FRAMES)
    ctx.synchronize()
    ctx.timer_start()
    for i in range(40):
        ctx.extract_device(bufs[i % 4].ptr, W, W * H, FRAMES)
    ms = ctx.timer_stop() / 40
    print("streams %d slices %2d : %.3f ms per 256 frames = %.0f frames/s" % (streams, parts, ms, FRAMES / ms * 1e3), flush=True)
