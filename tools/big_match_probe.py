import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/slam-module_b200')
import numpy as np
import slamgpu, synth
from oracle import pyoracle as po
ctx = slamgpu.Context(640, 480, max_frames=1)
for nA, nB in ((8000, 6000), (20000, 15000), (65535, 300), (300, 65535)):
    rng = np.random.default_rng(nA)
    dB = rng.integers(0, 2 ** 32, (nB, 8), dtype=np.uint32)
    aB = rng.uniform(0, 360, nB).astype(np.float32)
    src = rng.integers(0, nB, nA)
    dA = dB[src].copy()
    flips = rng.integers(0, 40, nA)
    for i in range(nA):
        for b in rng.integers(0, 256, flips[i]):
            dA[i, b >> 5] ^= np.uint32(1 << (int(b) & 31))
    aA = ((aB[src] + rng.normal(0, 4, nA)) % 360).astype(np.float32)
    t0 = time.time(); rn, rm = po.match_bruteforce(dA, aA, dB, aB); t1 = time.time()
    n, m = ctx.match_bruteforce(dA, aA, dB, aB); t2 = time.time()
    n, m = ctx.match_bruteforce(dA, aA, dB, aB); t3 = time.time()
    print(nA, nB, "matches", n, rn, "equal:", n == rn and np.array_equal(m, rm), "rescans", ctx.rescans(), "oracle %.1f s, gpu first call %.1f ms, second %.1f ms" % (t1 - t0, 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
