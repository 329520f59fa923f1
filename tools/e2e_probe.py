#!/usr/bin/env python3
"""Where does the end-to-end (host buffer) time of sg_extract go?  Raw pinned H2D / D2H bandwidth, the call
with and without result download, and the chunk size sweep."""
import ctypes as C, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import slam_module_b200 as sm
from slam_module_b200 import slamgpu
W, H, F = 640, 480, 256
lib = slamgpu.lib()
ctx = slamgpu.Context(W, H, max_frames=F)
pin = slamgpu.PinnedArray((F, H, W), np.uint8)
base = [sm.synth.frame(W, H, 100 + i) for i in range(8)]
for i in range(F):
    pin.array[i] = base[i % 8]
dbuf = ctx.device_buffer(F * W * H)
def t(fn, n=10):
    fn(); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    ctx.synchronize()
    return (time.perf_counter() - t0) / n
dt = t(lambda: lib.sg_memcpy_h2d(ctx._h, dbuf.ptr, pin.array.ctypes.data, F * W * H))
print("raw H2D %.1f MB: %.3f ms  %.1f GB/s" % (F * W * H / 1e6, dt * 1e3, F * W * H / dt / 1e9))
arrs, ks = ctx._alloc_out(F, pinned=True)
nb = sum(a.nbytes for a in arrs.values())
dt = t(lambda: lib.sg_memcpy_d2h(ctx._h, arrs["desc"].ctypes.data, dbuf.ptr, arrs["desc"].nbytes))
print("raw D2H %.1f MB: %.3f ms  %.1f GB/s" % (arrs["desc"].nbytes / 1e6, dt * 1e3, arrs["desc"].nbytes / dt / 1e9))
dt = t(lambda: ctx.extract_device(dbuf.ptr, W, W * H, F))
print("device-resident extract: %.3f ms" % (dt * 1e3))
none = slamgpu.Keypoints(None, None, None, None, None, None, None, None, arrs["count"].ctypes.data, None)
for chunk in (16, 24, 32, 40, 48, 64, 128):
    ctx.set_pipeline_chunk(chunk)
    a = t(lambda: ctx._check(lib.sg_extract(ctx._h, pin.array.ctypes.data, W, W * H, F, None, None, None, C.byref(ks))))
    b = t(lambda: ctx._check(lib.sg_extract(ctx._h, pin.array.ctypes.data, W, W * H, F, None, None, None, C.byref(none))))
    print("chunk %3d: full %.3f ms (%.0f fps)   counts only %.3f ms" % (chunk, a * 1e3, F / a, b * 1e3))
for n in (8, 16, 32, 64, 128, 256):
    dt = t(lambda: ctx.extract_device(dbuf.ptr, W, W * H, n), 20)
    print("device-resident extract of %3d frames: %.3f ms  (%.1f us/frame)" % (n, dt * 1e3, dt * 1e6 / n))
for parts in (1, 2, 4, 6, 8):
    ctx.set_overlap(parts)
    dt = t(lambda: ctx.extract_device(dbuf.ptr, W, W * H, F), 20)
    print("overlap parts %d: device-resident extract %.3f ms (%.0f fps)" % (parts, dt * 1e3, F / dt))
dA, aA, dB, aB = sm.synth.correlated_descriptors(2000, 7)
ctx.match_bruteforce(dA, aA, dB, aB)
t0 = time.perf_counter()
for _ in range(20):
    n, m = ctx.match_bruteforce(dA, aA, dB, aB)
print("sg_match_bruteforce 2000 x 2000, host buffers: %.3f ms per call (%d matches)" % ((time.perf_counter() - t0) / 20 * 1e3, n))
