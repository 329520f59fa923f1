"""Soak: the same batch through the device-resident, the single-call and the streaming entry points, many times; every
output array must be bit-identical every time (atomics only decide the order of intermediate lists)."""
import os, sys, zlib
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "slam-module_b200"))
import slamgpu, synth

W, H, F, IT = 640, 480, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 60
imgs = synth.frames(W, H, 16, 777)[np.arange(F) % 16]
ctx = slamgpu.Context(W, H, max_keypoints=2000, max_frames=2 * F)
pin = slamgpu.PinnedArray((F, H, W), np.uint8); pin.array[...] = imgs
buf = ctx.device_buffer(imgs.nbytes).upload(imgs)
outs = [ctx.alloc_outputs(F, pinned=True) for _ in range(2)]


def digest(arrs):
    h = 0
    cnt = arrs["count"]
    for k in ("x", "y", "angle", "octave", "desc", "track_id", "lvl_x", "lvl_y"):
        for f in range(F):
            h = zlib.crc32(np.ascontiguousarray(arrs[k][f, :cnt[f]]).tobytes(), h)
    return zlib.crc32(cnt.tobytes(), h)


ref = None
bad = 0
for it in range(IT):
    mode = it % 3
    if mode == 0:
        ctx.extract_device(buf.ptr, W, W * H, F)
        got = ctx.extract_download(F)
        arrs = {k: np.stack([np.pad(g[k], [(0, ctx.cap - g["n"])] + [(0, 0)] * (g[k].ndim - 1)) for g in got]) for k in ("x", "y", "angle", "octave", "desc", "track_id", "lvl_x", "lvl_y")}
        arrs["count"] = np.array([g["n"] for g in got], np.int32)
    elif mode == 1:
        import ctypes as C
        ctx._check(slamgpu.lib().sg_extract(ctx._h, pin.array.ctypes.data, W, W * H, F, None, None, None, C.byref(outs[0][1])))
        arrs = outs[0][0]
    else:
        t0 = ctx.extract_submit(pin.array, 0, outs[0][1]); t1 = ctx.extract_submit(pin.array, F, outs[1][1])
        ctx.extract_wait(t0); ctx.extract_wait(t1)
        assert digest(outs[0][0]) == digest(outs[1][0])
        arrs = outs[1][0]
    d = digest(arrs)
    if ref is None:
        ref = d
    bad += d != ref
print("iterations %d, differing digests %d, keypoints per frame %.1f" % (IT, bad, float(np.mean(arrs["count"]))))
sys.exit(1 if bad else 0)
