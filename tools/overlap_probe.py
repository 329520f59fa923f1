"""Device-resident extraction throughput with 1 or 2 contexts alternating batches (cross-step overlap)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "slam-module_b200"))
import slamgpu, synth

W, H, FRAMES = 640, 480, 256
imgs = synth.frames(W, H, 8, 100)[np.arange(FRAMES) % 8]
for n_ctx in (1, 2, 3):
    ctxs = [slamgpu.Context(W, H, max_keypoints=2000, max_frames=FRAMES) for _ in range(n_ctx)]
    bufs = [c.device_buffer(FRAMES * W * H).upload(np.ascontiguousarray(np.roll(imgs, i, axis=0))) for i, c in enumerate(ctxs)]
    for parts in (4, 2, 1):
        for c in ctxs:
            c.set_overlap(parts)
        def run(n):
            for i in range(n):
                c = ctxs[i % n_ctx]
                c.extract_device(bufs[i % n_ctx].ptr, W, W * H, FRAMES)
            for c in ctxs:
                c.synchronize()
        run(6)
        t0 = time.perf_counter()
        run(60)
        dt = (time.perf_counter() - t0) / 60
        print("contexts %d  slices %d : %.3f ms per batch = %.0f frames/s" % (n_ctx, parts, dt * 1e3, FRAMES / dt))
    for c in ctxs:
        c.close()
