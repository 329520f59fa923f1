"""Streaming sg_extract_submit / sg_extract_wait throughput against the chunk size and the batches in flight."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "slam-module_b200"))
import slamgpu, synth

W, H, FRAMES = 640, 480, 256
DEPTH = 4
ctx = slamgpu.Context(W, H, max_keypoints=2000, max_frames=DEPTH * FRAMES)
pins = []
for b in range(3):
    p = slamgpu.PinnedArray((FRAMES, H, W), np.uint8)
    p.array[...] = synth.frames(W, H, 8, 100 + b)[np.arange(FRAMES) % 8]
    pins.append(p)
outs = [ctx.alloc_outputs(FRAMES, pinned=True) for _ in range(DEPTH)]


def run(n, depth):
    tickets = [None] * depth
    for i in range(n):
        s = i % depth
        if tickets[s] is not None:
            ctx.extract_wait(tickets[s])
        tickets[s] = ctx.extract_submit(pins[i % 3].array, s * FRAMES, outs[s][1])
    for t in tickets:
        if t is not None:
            ctx.extract_wait(t)


for chunk in (43, 64, 86, 128, 256):
    ctx.set_pipeline_chunk(chunk)
    for depth in (1, 2, 3, 4):
        run(3, depth)
        t0 = time.perf_counter()
        run(12, depth)
        dt = (time.perf_counter() - t0) / 12
        # host-side submit cost alone
        t0 = time.perf_counter()
        t = ctx.extract_submit(pins[0].array, 0, outs[0][1])
        ts = time.perf_counter() - t0
        ctx.extract_wait(t)
        print("chunk %3d  in flight %d : %.3f ms per batch = %.0f frames/s   (submit call %.3f ms)" % (chunk, depth, dt * 1e3, FRAMES / dt, ts * 1e3))
