import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/slam-module_b200')
import numpy as np
import slamgpu, synth
from oracle import pyoracle as po
for (w, h, kp) in ((3840, 2160, 8000), (1920, 480, 1500), (480, 1920, 1500), (1024, 96, 300), (4096, 64, 500), (64, 4096, 500)):
    img = synth.frame(w, h, 99)
    t0 = time.time()
    ref = po.extract(po.make_params(w, h, max_keypoints=kp), img)
    t1 = time.time()
    with slamgpu.Context(w, h, max_keypoints=kp, max_frames=1) as ctx:
        got = ctx.detect_and_extract(img)[0]
    same = got["n"] == ref["n"] and all(np.array_equal(got[k], ref[k]) for k in ("x", "y", "octave", "angle", "desc"))
    print(w, h, kp, "n =", got["n"], ref["n"], "equal:", same, "oracle %.1f s" % (t1 - t0))
