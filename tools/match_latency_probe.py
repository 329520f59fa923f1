"""Latency of the single-call matchers (one keyframe pair per call) and of small batches."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "slam-module_b200"))
import slamgpu, synth

ctx = slamgpu.Context(640, 480, max_frames=1)
dA, aA, dB, aB = synth.correlated_descriptors(2000, 500)
def timed(fn, n=20):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    return (time.perf_counter() - t0) / n * 1e3, r
ms, (n, _) = timed(lambda: ctx.match_bruteforce(dA, aA, dB, aB))
print("sg_match_bruteforce 2000 x 2000 (host buffers): %.3f ms per call, %d matches" % (ms, n))
db = slamgpu.DescriptorDB(ctx, np.stack([dA, dB]), np.stack([aA, aB]))
for npairs in (1, 4, 16, 64, 256):
    pairs = np.tile(np.array([[0, 1]], np.int32), (npairs, 1))
    d_pairs = ctx.device_buffer(pairs.nbytes).upload(pairs)
    d_counts = ctx.device_buffer(4 * npairs)
    def run():
        db.match_pairs_device(d_pairs.ptr, npairs, d_counts.ptr); ctx.synchronize()
    ms, _ = timed(run)
    print("match_pairs_device %4d pairs: %.3f ms per call = %.1f us per pair" % (npairs, ms, 1e3 * ms / npairs))
