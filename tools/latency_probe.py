"""Single-frame extraction (BASELINE configs[0]: 640x480, 1000 keypoints): host-call latency of sg_extract and the
kernel chain alone; run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel durations."""
import os, sys, time, statistics, ctypes as C
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import slam_module_b200 as sm
from slam_module_b200 import slamgpu

W, H = 640, 480
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
lib = slamgpu.lib()
with slamgpu.Context(W, H, max_keypoints=1000, max_frames=1) as c:
    pin = slamgpu.PinnedArray((1, H, W), np.uint8)
    pin.array[0] = sm.synth.frame(W, H, 1000)
    arrs, ks = c._alloc_out(1, pinned=True)
    lat = []
    for i in range(n):
        t0 = time.perf_counter()
        c._check(lib.sg_extract(c._h, pin.array.ctypes.data, W, W * H, 1, None, None, None, C.byref(ks)))
        lat.append(time.perf_counter() - t0)
    d = c.device_buffer(W * H).upload(pin.array)
    for _ in range(5):
        c.extract_device(d.ptr, W, W * H, 1)
    c.synchronize()
    c.timer_start()
    for _ in range(n):
        c.extract_device(d.ptr, W, W * H, 1)
    dev_us = c.timer_stop() * 1e3 / n
    print("sg_extract host call: median %.1f us, min %.1f us; kernel chain alone: %.1f us; %d keypoints"
          % (1e6 * statistics.median(lat[n // 4:]), 1e6 * min(lat), dev_us, int(arrs["count"][0])))
