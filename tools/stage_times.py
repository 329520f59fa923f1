#!/usr/bin/env python3
"""Kernel-iteration aid (run under gpurun): per-stage CUDA-event times of one 256-frame extraction step (BASELINE
configs[1]), the overlapped whole-step time, and a bit-exactness check of a few frames against the oracle.
    python tools/stage_times.py [frames] [check_frames]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import slam_module_b200 as sm  # noqa: E402,F401
from slam_module_b200 import slamgpu  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_check = int(sys.argv[2]) if len(sys.argv) > 2 else 4
W, H = bench.W, bench.H
ctx = slamgpu.Context(W, H, max_keypoints=bench.MAXKP, max_frames=frames)
batches = [bench.make_frames(frames, 10000 + 100 * b) for b in range(2)]
dev = [ctx.device_buffer(frames * W * H).upload(b) for b in batches]
for i in range(5):
    ctx.extract_device(dev[i % 2].ptr, W, W * H, frames)
ctx.synchronize()
ctx.timer_start()
for i in range(20):
    ctx.extract_device(dev[i % 2].ptr, W, W * H, frames)
step_ms = ctx.timer_stop() / 20
ctx.set_profiling(True)
for i in range(10):
    ctx.extract_device(dev[i % 2].ptr, W, W * H, frames)
stage = ctx.stage_ms()
ctx.set_profiling(False)
bad = 0
if n_check:
    got = ctx.detect_and_extract(batches[0][:n_check])
    p = po.make_params(W, H, max_keypoints=bench.MAXKP)
    for f in range(n_check):
        ref = po.extract(p, batches[0][f])
        ok = got[f]["n"] == ref["n"] and all(np.array_equal(got[f][k], ref[k]) for k in ("x", "y", "octave", "angle", "desc"))
        bad += 0 if ok else 1
print(json.dumps({"frames": frames, "step_ms_overlapped": step_ms, "frames_per_s": frames / step_ms * 1e3,
                  "stage_ms": {k: round(v, 4) for k, v in stage.items()}, "frames_checked": n_check, "frames_wrong": bad}))
