"""Where the time of one stereo step (bench.py configs[3]) goes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "slam-module_b200"))
import slamgpu, synth

SW, SH, PAIRS = 1280, 720, 16
c4 = slamgpu.Context(SW, SH, max_keypoints=2000, max_frames=2 * PAIRS)
left = [synth.frame(SW, SH, 4000 + i) for i in range(4)]
frames = np.empty((2 * PAIRS, SH, SW), np.uint8)
for i in range(PAIRS):
    frames[2 * i] = left[i % 4]
    frames[2 * i + 1] = np.roll(left[i % 4], -6 - (i % 3), axis=1)
dfr = c4.device_buffer(frames.nbytes).upload(frames)
cap = c4.cap
pairs = np.array([(4 * i, 4 * i + 2) for i in range(PAIRS)], np.int32)
d_pairs = c4.device_buffer(pairs.nbytes).upload(pairs)
d_counts = c4.device_buffer(4 * PAIRS)
for it in range(8):
    t = [time.perf_counter()]
    c4.extract_device(dfr.ptr, SW, SW * SH, 2 * PAIRS); c4.synchronize(); t.append(time.perf_counter())
    counts, _ = c4.extract_download(2 * PAIRS, only_counts=True); t.append(time.perf_counter())
    offs = np.zeros(4 * PAIRS + 1, np.int64)
    offs[1::2] = np.arange(2 * PAIRS) * cap + counts
    offs[2::2] = (np.arange(2 * PAIRS) + 1) * cap
    v = c4.device_views()
    db = slamgpu.DescriptorDB(c4, None, None, offsets=offs, device_ptrs=(v.desc, v.angle)); t.append(time.perf_counter())
    db.match_pairs_device(d_pairs.ptr, PAIRS, d_counts.ptr); c4.synchronize(); t.append(time.perf_counter())
    db.close(); t.append(time.perf_counter())
    print("iter %d: extract %.2f  counts %.2f  db_create %.2f  match %.2f  db_close %.2f ms" % ((it,) + tuple(1e3 * (b - a) for a, b in zip(t, t[1:]))))
c4.set_profiling(True)
for it in range(3):
    db = slamgpu.DescriptorDB(c4, None, None, offsets=offs, device_ptrs=(v.desc, v.angle), view=True)
    db.match_pairs_device(d_pairs.ptr, PAIRS, d_counts.ptr); c4.synchronize()
    db.close()
print("stage ms", c4.stage_ms(), "rescans", c4.rescans(), "matches", d_counts.download(np.uint32, PAIRS)[:4])
db = slamgpu.DescriptorDB(c4, None, None, offsets=offs, device_ptrs=(v.desc, v.angle), view=True)
t0 = time.perf_counter(); n, m = db.match_pairs(pairs); t1 = time.perf_counter()
print("host-variant call %.2f ms, rescans over %d pairs: %d" % (1e3 * (t1 - t0), PAIRS, c4.rescans()))
