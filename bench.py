#!/usr/bin/env python3
"""Benchmark of the ORB front-end + Hamming matching hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libslamgpu.so)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

One "step" = one pass of the hot path over one batch: ORB extraction (pyramid -> FAST + quadtree ->
orientation -> rBRIEF) of 256 synthetic 640x480 frames, 8 levels, 2000 keypoints (BASELINE.json
configs[1]).  `value` = frames/s with the input batch already resident in HBM; `e2e` = the same
through the C-ABI call that takes HOST buffers (H2D of the frames and D2H of the keypoints inside
the timed region).  The second half of the metric, Hamming matching (configs[2]: 2000 x 2000
descriptors per keyframe pair, ratio 0.8, angle histogram), is timed in the same run and reported
under "matching" with its own roofline (integer POPC issue rate, measured by a micro-benchmark).

Multi-GPU (torchrun, one rank per GPU): frames / keyframe pairs are sharded by rank, no collective
on the data path; torch.distributed (NCCL) is used only for the barrier and the max-over-ranks.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W, H, LEVELS, FACTOR, MAXKP = 640, 480, 8, 1.2, 2000
FRAMES = 256                      # frames per step per GPU
WORKLOAD = "ORB extraction, %d synthetic 640x480 frames per GPU per step, 8 levels x1.2, 2000 keypoints (BASELINE configs[1])" % FRAMES
IN_FLIGHT = 3                     # batches in flight in the streaming end-to-end pass (frame slots of the context)
N_BATCHES = 4                     # rotating input batches: 4 x 78.6 MB > 126 MB of L2
ALGO_BYTES_PYRAMID = 2208264      # A0 + 2*sum(A_l) per 640x480 frame
ALGO_BYTES_FAST = 950532          # sum(A_l): every level read once
MATCH_SETS, MATCH_N, MATCH_PAIRS = 32, 2000, 2048   # descriptor sets, descriptors per set, pairs per step
METRIC = "ORB frames/sec (640x480, 8 lvls, 2k kp) + Hamming matches/sec at 1/2/4/8 B200"


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        p = json.loads(f.read_text())
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def traffic_of(stage):
    """ncu-measured DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the stage's kernel(s) for one full
    256-frame step, from the newest profiles/traffic_r*.json (written by tools/summarize_profiles.py from full-batch
    `ncu --set full` captures of the shipped kernels); returns (bytes, tag) or (None, None)."""
    files = sorted((ROOT / "profiles").glob("traffic_r*.json"))
    if not files:
        return None, None
    t = json.loads(files[-1].read_text())
    e = t.get({"fast": "fast_cells_kernel", "pyramid": "pyr_fast_kernel (all 8 launches)"}.get(stage, ""), None)
    if e is None and stage == "pyramid":
        e = t.get("pyr_fast_kernel<1> level 1")
    return (e.get("bytes_per_launch"), t.get("tag", files[-1].stem)) if e else (None, None)


def make_frames(n, seed0):
    """n distinct synthetic frames: 32 generated from scratch, the rest rolled / flipped variants."""
    import slam_module_b200 as sm
    base = [sm.synth.frame(W, H, seed0 + i) for i in range(min(n, 32))]
    rng = np.random.default_rng(seed0)
    out = np.empty((n, H, W), np.uint8)
    for i in range(n):
        b = base[i % len(base)]
        if i >= len(base):
            b = np.roll(b, (int(rng.integers(1, H)), int(rng.integers(1, W))), axis=(0, 1))
            if rng.random() < 0.5:
                b = b[:, ::-1]
        out[i] = b
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 6:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
            except ValueError:
                continue
            for nm, v in zip(names, c[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        # median over the samples taken under load (top half of the observed clocks)
        under = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(under) if under else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    td = None
    if world > 1:
        import torch
        import torch.distributed as td_
        torch.cuda.set_device(local)
        # NCCL prints its version banner on stdout when the communicator comes up: keep stdout for the ONE json line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            td_.init_process_group("nccl", device_id=torch.device("cuda", local))
            t = torch.zeros(1, device=torch.device("cuda", local))
            td_.all_reduce(t)                      # forces communicator creation inside the redirected region
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        td = td_
    return rank, world, local, td


def barrier_max(td, local, value):
    """max over ranks of a python float (device tensor all-reduce); identity without a process group."""
    if td is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=torch.device("cuda", local))
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def cpu_extract_baseline(be, frames, threads):
    from oracle import pyoracle
    p = pyoracle.make_params(W, H, levels=LEVELS, scale_factor=FACTOR, max_keypoints=MAXKP)
    secs, total = be.bench_extract(p, frames, threads)
    return len(frames) / secs, secs, total


def reference_backend():
    """(module, kind, description): oracle/_ref/libref_slam.so -- the reference's own orb_extractor.cpp / image_pyramid.cpp /
    feature_detector.cpp / keyframe_matcher.cpp compiled verbatim (OpenCV primitives and the absent corner detector from
    the oracle port, see oracle/ref_slam.cpp) -- when it was built, else the oracle port."""
    from oracle import pyoracle as po
    po.build()
    try:
        from oracle import pyref as pr
        if pr.available():
            return pr, po, "reference", ("the reference's own sources compiled verbatim (oracle/_ref/libref_slam.so: orb_extractor.cpp, "
                                         "image_pyramid.cpp, feature_detector.cpp, keyframe_matcher.cpp; OpenCV primitives and the absent corner "
                                         "detector from the oracle port)")
    except Exception:
        pass
    return po, po, "port", "oracle port of the reference CPU path (oracle/_ref was not built: no reference tree at build time)"


def cv2_baseline(frames, desc_sets, pairs, threads):
    """'Best available CPU library' column: OpenCV's own optimised kernels on all host cores.
    Not the reference and not bit-comparable with it -- cv2.ORB is OpenCV's ORB (Harris-ranked, no cell grid / quadtree);
    reported so that the GPU numbers can be read against SIMD CPU code rather than against a scalar port only."""
    try:
        import cv2
    except Exception as e:  # pragma: no cover
        return {"unavailable": str(e)}
    from concurrent.futures import ThreadPoolExecutor
    cv2.setNumThreads(1)
    local = threading.local()

    def orb_frame(i):
        if not hasattr(local, "orb"):
            local.orb = cv2.ORB_create(nfeatures=MAXKP, scaleFactor=FACTOR, nlevels=LEVELS, edgeThreshold=19, patchSize=31, fastThreshold=20)
        kp, _ = local.orb.detectAndCompute(frames[i], None)
        return len(kp)

    def stage_frame(i):
        # the OpenCV calls the reference itself makes (image_pyramid.cpp:75-85) + cv::FAST(20, nms) on every level
        if not hasattr(local, "fast"):
            local.fast = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True)
        img, n = frames[i], 0
        scale = np.float32(1.0)
        for l in range(LEVELS):
            if l:
                scale = np.float32(FACTOR) * scale
                size = (int(round(W / float(scale))), int(round(H / float(scale))))
                img = cv2.resize(img, size, interpolation=cv2.INTER_LINEAR)
            cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            n += len(local.fast.detect(img, None))
        return n

    def match_pair(k):
        if not hasattr(local, "bf"):
            local.bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        a, b = pairs[k]
        m = local.bf.knnMatch(desc_sets[a], desc_sets[b], k=2)
        return sum(1 for x in m if len(x) == 2 and x[0].distance <= 50 and x[0].distance <= 0.8 * x[1].distance)

    out = {"cores": threads, "opencv": cv2.__version__}
    with ThreadPoolExecutor(threads) as ex:
        for name, fn, n, unit in (("orb", orb_frame, len(frames), "frames/s"), ("pyramid_fast", stage_frame, len(frames), "frames/s"),
                                  ("bfmatcher", match_pair, len(pairs), "keyframe pairs/s")):
            list(ex.map(fn, range(min(n, threads))))                    # warm-up: per-thread objects
            t0 = time.perf_counter()
            r = list(ex.map(fn, range(n)))
            dt = time.perf_counter() - t0
            out[name] = {"value": n / dt, "unit": unit, "items": n, "seconds": dt, "mean_result": float(np.mean(r))}
    out["orb"]["what"] = "cv2.ORB_create(2000, 1.2, 8).detectAndCompute per frame (OpenCV's complete ORB)"
    out["pyramid_fast"]["what"] = "cv2.resize chain + cv2.GaussianBlur(7x7, sigma 2) + cv2.FAST(20, nms) on all 8 levels per frame"
    out["bfmatcher"]["what"] = "cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) + ratio 0.8 on 2000 x 2000 descriptors per pair"
    out["bfmatcher"]["descriptor_pair_distances_per_s"] = out["bfmatcher"]["value"] * MATCH_N * MATCH_N
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (oracle/_ref when the
    reference compiled at build time, else the oracle port), frames sharded over all threads; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    be, po, kind, what = reference_backend()
    cores = os.cpu_count() or 1
    sample = FRAMES            # the whole step: 256 frames
    frames = make_frames(sample, 9000)
    for _ in range(max(1, args.warmup)):
        cpu_extract_baseline(be, frames[:max(2 * cores, 8)], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_extract_baseline(be, frames, cores)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    # matching leg on a few pairs
    import slam_module_b200 as sm
    d, a = sm.synth.random_descriptors(8, MATCH_N, 5)
    pairs = np.array([(i % 8, (i + 1) % 8) for i in range(4 * max(8, cores))], np.int32)
    be.bench_match(d, a, pairs[:cores], cores)
    msec, _ = be.bench_match(d, a, pairs, cores)
    cv2_cols = None
    if not args.skip_cv2:
        cv2_cols = cv2_baseline(frames[:max(64, 4 * cores)], [np.ascontiguousarray(x).view(np.uint8).reshape(MATCH_N, 32) for x in d],
                                pairs[:2 * cores], cores)
    desc = "%s, %d frames per step sharded over %d host threads" % (what, sample, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": sample, "levels": LEVELS, "scale_factor": FACTOR,
                   "max_keypoints": MAXKP},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": desc, "cv2": cv2_cols},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "matching": {"value": len(pairs) * MATCH_N * MATCH_N / msec, "unit": "descriptor-pair distances/s",
                     "keyframe_pairs_per_s": len(pairs) / msec, "cores": cores},
        "matching_value": len(pairs) * MATCH_N * MATCH_N / msec,
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def extra_configs(slamgpu, sm, sh, rank, world, local, barrier_max_, td, config5_full=False):
    """The remaining BASELINE.json configs, each on a bounded sample, reported beside the headline
    (configs[1]) and the matching leg (configs[2]):
      configs[0]  one 640x480 frame, 1000 keypoints: latency of sg_extract (host buffers) and of the kernels;
      configs[3]  1280x720 stereo pairs: extract L and R, match L -> R (M1 semantics), pairs sharded by rank;
      configs[4]  loop-closure scale: 10 000 keyframes x 2000 descriptors resident per GPU (640 MB), the
                  all-pairs candidate list block-sharded over the ranks, a bounded block of it timed."""
    import ctypes as C
    lib = slamgpu.lib()
    out = {}
    # ---- configs[0]: single-frame latency -------------------------------------------------------------
    with slamgpu.Context(W, H, levels=LEVELS, scale_factor=FACTOR, max_keypoints=1000, max_frames=1, device=local) as c1:
        pin = slamgpu.PinnedArray((1, H, W), np.uint8)
        pin.array[0] = sm.synth.frame(W, H, 1000)
        arrs, ks = c1._alloc_out(1, pinned=True)
        lat = []
        for i in range(25):
            t0 = time.perf_counter()
            c1._check(lib.sg_extract(c1._h, pin.array.ctypes.data, W, W * H, 1, None, None, None, C.byref(ks)))
            lat.append(time.perf_counter() - t0)
        dbuf = c1.device_buffer(W * H).upload(pin.array)
        for _ in range(3):
            c1.extract_device(dbuf.ptr, W, W * H, 1)
        c1.synchronize()
        c1.timer_start()
        for _ in range(20):
            c1.extract_device(dbuf.ptr, W, W * H, 1)
        dev_us = c1.timer_stop() * 1e3 / 20
        out["config0_single_frame"] = {"workload": "1 frame 640x480, 8 levels, 1000 keypoints", "keypoints": int(arrs["count"][0]),
                                       "latency_us_host_call_median": 1e6 * statistics.median(lat[5:]),
                                       "latency_us_kernels_only": dev_us}
        if rank == 0 and world == 1:
            # the same frame through the reference's CPU path on ONE host thread
            be, po_, kind, _ = reference_backend()
            p1 = po_.make_params(W, H, levels=LEVELS, scale_factor=FACTOR, max_keypoints=1000)
            one = np.ascontiguousarray(pin.array[:1])
            be.bench_extract(p1, np.repeat(one, 2, axis=0), 1)
            secs, _ = be.bench_extract(p1, np.repeat(one, 8, axis=0), 1)
            out["config0_single_frame"]["cpu_single_thread_ms_per_frame"] = 1e3 * secs / 8
            out["config0_single_frame"]["cpu_kind"] = kind
    # ---- configs[3]: 1280x720 stereo stream --------------------------------------------------------------
    SW, SH, PAIRS = 1280, 720, 16
    with slamgpu.Context(SW, SH, levels=LEVELS, scale_factor=FACTOR, max_keypoints=MAXKP, max_frames=2 * PAIRS, device=local) as c4:
        left = [sm.synth.frame(SW, SH, 4000 + 10 * rank + i) for i in range(4)]
        frames = np.empty((2 * PAIRS, SH, SW), np.uint8)
        for i in range(PAIRS):
            frames[2 * i] = left[i % 4]
            frames[2 * i + 1] = np.roll(left[i % 4], -6 - (i % 3), axis=1)      # right view: horizontal disparity
        dfr = c4.device_buffer(frames.nbytes).upload(frames)
        cap = c4.cap
        pairs = np.array([(4 * i, 4 * i + 2) for i in range(PAIRS)], np.int32)   # sets 2f (keypoints) / 2f+1 (padding)
        d_pairs = c4.device_buffer(pairs.nbytes).upload(pairs)
        d_counts = c4.device_buffer(4 * PAIRS)

        def stereo_step():
            c4.extract_device(dfr.ptr, SW, SW * SH, 2 * PAIRS)
            counts, _ = c4.extract_download(2 * PAIRS, only_counts=True)          # host gather of the counts
            offs = np.zeros(4 * PAIRS + 1, np.int64)
            offs[1::2] = np.arange(2 * PAIRS) * cap + counts
            offs[2::2] = (np.arange(2 * PAIRS) + 1) * cap
            v = c4.device_views()
            db = slamgpu.DescriptorDB(c4, None, None, offsets=offs, device_ptrs=(v.desc, v.angle), view=True)   # no copy
            db.match_pairs_device(d_pairs.ptr, PAIRS, d_counts.ptr)
            c4.synchronize()
            db.close()
            return counts

        for _ in range(2):
            counts = stereo_step()
        barrier_max_(td, local, 0.0)
        t0 = time.perf_counter()
        for _ in range(5):
            stereo_step()
        dt = barrier_max_(td, local, time.perf_counter() - t0)
        nm = d_counts.download(np.uint32, PAIRS)
        out["config3_stereo_stream"] = {"workload": "1280x720 stereo pairs: extract L+R (2000 kp) and match L->R, %d pairs per step per GPU" % PAIRS,
                                        "stereo_pairs_per_s": world * PAIRS * 5 / dt, "frames_per_s": world * 2 * PAIRS * 5 / dt,
                                        "keypoints_per_frame": float(counts.mean()), "matches_per_pair": float(nm.mean()),
                                        "sharding": "stereo pairs by rank; L and R of a pair stay on one GPU"}
    # ---- configs[4]: loop-closure scale --------------------------------------------------------------------
    KF, N, SAMPLE = 10000, MATCH_N, 4096
    rng = np.random.default_rng(99)
    base = rng.integers(0, 2 ** 32, (64, N, 8), dtype=np.uint32)                 # 64 distinct places
    noise = np.ones((157, N, 8), np.uint32) * np.uint32(0xffffffff)
    for _ in range(4):                                                           # each bit flipped with p = 1/16
        noise &= rng.integers(0, 2 ** 32, (157, N, 8), dtype=np.uint32)
    ang_base = rng.uniform(0, 360, (64, N)).astype(np.float32)
    with slamgpu.Context(W, H, max_frames=1, device=local) as c5:
        desc = np.empty((KF, N, 8), np.uint32)
        ang = np.empty((KF, N), np.float32)
        for k0 in range(0, KF, 64):
            n = min(64, KF - k0)
            desc[k0:k0 + n] = base[:n] ^ noise[(k0 // 64) % 157][None]
            ang[k0:k0 + n] = ang_base[:n]
        db = slamgpu.DescriptorDB(c5, desc, ang)                                 # replicated on every GPU
        total = sh.n_unordered_pairs(KF)
        lo, hi = sh.block_range(total, rank, world)
        # a bounded block of this rank's share of the all-pairs list, taken where every 64th pair revisits a place
        blk = sh.pair_block(KF, rank, world, limit=SAMPLE)
        d_blk = c5.device_buffer(blk.nbytes).upload(blk)
        d_cnt = c5.device_buffer(4 * SAMPLE)
        db.match_pairs_device(d_blk.ptr, len(blk), d_cnt.ptr)
        c5.synchronize()
        barrier_max_(td, local, 0.0)
        c5.timer_start()
        db.match_pairs_device(d_blk.ptr, len(blk), d_cnt.ptr)
        ms = barrier_max_(td, local, c5.timer_stop())
        cnt = d_cnt.download(np.uint32, len(blk))
        rate = world * len(blk) / (ms * 1e-3)
        out["config4_loop_closure_scale"] = {
            "workload": "10000 keyframes x 2000 descriptors resident per GPU (%.0f MB), all %d unordered pairs block-sharded over %d rank(s); "
                        "timed sample: the first %d pairs of each rank's block" % (desc.nbytes / 1e6, total, world, len(blk)),
            "keyframe_pairs_per_s": rate, "descriptor_pair_distances_per_s": rate * N * N,
            "projected_full_job_s": total / rate, "pairs_in_rank_block": int(hi - lo),
            "pairs_with_matches_in_sample": int((cnt > 0).sum())}
        if config5_full:
            # the whole job: every rank matches its complete block of the all-pairs list (counts only), wall time from the
            # barrier before the first launch to the slowest rank's last count; 256 random pairs of rank 0 re-done by the oracle
            full = sh.pair_block(KF, rank, world)
            d_full = c5.device_buffer(full.nbytes).upload(full)
            d_fcnt = c5.device_buffer(4 * len(full))
            c5.synchronize()
            barrier_max_(td, local, 0.0)
            t0 = time.perf_counter()
            db.match_pairs_device(d_full.ptr, len(full), d_fcnt.ptr)
            c5.synchronize()
            wall = barrier_max_(td, local, time.perf_counter() - t0)
            fcnt = d_fcnt.download(np.uint32, len(full))
            tot_matches = barrier_max_(td, local, 0.0)   # (barrier) ; matches are summed below through the max trick per rank
            sums = [barrier_max_(td, local, float(fcnt.sum()) if r == rank else 0.0) for r in range(world)]
            checked, bad = 0, 0
            if rank == 0:
                from oracle import pyoracle as po
                po.build()
                pick = np.random.default_rng(5).choice(len(full), size=min(256, len(full)), replace=False)
                # every 64th pair revisits a place: make sure matching pairs are among the checked ones
                hot = np.flatnonzero(fcnt > 0)
                if len(hot):
                    pick[:min(64, len(hot))] = hot[np.random.default_rng(6).choice(len(hot), size=min(64, len(hot)), replace=False)]
                for k in pick:
                    i, j = int(full[k, 0]), int(full[k, 1])
                    n_ref, _ = po.match_bruteforce(desc[i], ang[i], desc[j], ang[j])
                    checked += 1
                    bad += int(n_ref != int(fcnt[k]))
            out["config4_loop_closure_scale"]["full_job"] = {
                "pairs": int(total), "pairs_this_rank": int(len(full)), "wall_s": wall, "keyframe_pairs_per_s": total / wall,
                "descriptor_pair_distances_per_s": total / wall * N * N, "matches_total": float(sum(sums)),
                "pairs_with_matches_rank0": int((fcnt > 0).sum()), "oracle_spot_check": {"pairs": checked, "count_mismatches": bad}}
            del tot_matches
        db.close()
    # ---- candidate-list matchers and descriptor medoid (host-buffer calls) -------------
    with slamgpu.Context(W, H, max_frames=1, device=local) as c6:
        rng = np.random.default_rng(5)
        nk, nq = 2000, 1500
        kx = rng.uniform(0, W, nk).astype(np.float32); ky = rng.uniform(0, H, nk).astype(np.float32)
        koct = rng.integers(0, 8, nk).astype(np.int32)
        kdesc = rng.integers(0, 2 ** 32, (nk, 8), dtype=np.uint32)
        src = rng.integers(0, nk, nq)
        qx = (kx[src] + rng.normal(0, 3, nq)).astype(np.float32); qy = (ky[src] + rng.normal(0, 3, nq)).astype(np.float32)
        qr = np.full(nq, 20, np.float32)
        qdesc = kdesc[src] ^ (np.uint32(1) << rng.integers(0, 32, (nq, 8)).astype(np.uint32))
        sizes = rng.integers(2, 30, 4000)
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        mdesc = rng.integers(0, 2 ** 32, (int(offs[-1]), 8), dtype=np.uint32)
        def timed(fn, n=5):
            fn()
            t0 = time.perf_counter()
            for _ in range(n):
                r = fn()
            return (time.perf_counter() - t0) / n, r
        t_proj, (n_proj, _, _) = timed(lambda: c6.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=1, thr=100,
                                                                    taken=np.zeros(nk, np.uint8)))
        t_dup, (n_dup, _, _) = timed(lambda: c6.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=0, thr=50))
        t_med, _ = timed(lambda: c6.medoid(mdesc, offs))
        vocab = sm.synth.random_vocabulary(10, 4, 3)                     # 11 111 nodes, 10 000 words
        voc = slamgpu.Vocabulary(c6, vocab)
        t_bow, (bw, _, bn) = timed(lambda: voc.transform(kdesc, 2))
        nodes_k = bn
        nodes_q = voc.transform(qdesc, 2)[2]
        ang_k = rng.uniform(0, 360, nk).astype(np.float32); ang_q = rng.uniform(0, 360, nq).astype(np.float32)
        t_mbow, (n_mbow, _) = timed(lambda: c6.match_bow(qdesc, ang_q, nodes_q, kdesc, ang_k, nodes_k, check_orientation=False))
        bword, bweight, _ = voc.transform(kdesc, 2)
        t_bvec, (vw, vv) = timed(lambda: voc.bow_vector(bword, bweight))
        NB = 256                                                        # one extraction batch worth of keyframes per call
        bdesc = np.tile(kdesc, (NB, 1)) ^ np.repeat(rng.integers(0, 2 ** 32, (NB, 1), dtype=np.uint32), nk, axis=0)
        t_btr, (bw_all, bwt_all, _) = timed(lambda: voc.transform(bdesc, 2), n=3)
        boffs = (np.arange(NB + 1, dtype=np.int64) * nk)
        t_bvb, vecs_b = timed(lambda: voc.bow_vector_batch(bw_all, bwt_all, boffs), n=3)
        voc.close()
        NKF = 10000
        vecs = sm.synth.random_bow_vectors(NKF, 100000, 1500, 9, n_topics=40)
        bdb = slamgpu.BowDatabase(c6, NKF, 2048)
        for i, (w_, v_) in enumerate(vecs):
            bdb.add(0, i, w_, v_)
        t_sim, (_, sim_kf, _) = timed(lambda: bdb.similar(*vecs[17], self_key=(0, 17)))
        bdb.close()
        out["next_rows"] = {
            "search_by_projection": {"workload": "%d projected map points against %d keypoints, radius 20 px (keyframe_matcher.cpp:295-414 inner loop)" % (nq, nk),
                                     "ms_per_call_host_buffers": t_proj * 1e3, "matches": n_proj},
            "replace_duplication": {"workload": "same queries, best only, thr 50 (keyframe_matcher.cpp:482-499)",
                                    "ms_per_call_host_buffers": t_dup * 1e3, "matches": n_dup},
            "map_point_medoid": {"workload": "%d map points, 2..29 observations each (map_point.cpp:75-116)" % len(sizes),
                                 "ms_per_call_host_buffers": t_med * 1e3, "map_points_per_s": len(sizes) / t_med},
            "bow_transform": {"workload": "%d descriptors through a synthetic 10-ary, 4-level vocabulary tree (bow_index.cpp:59-93)" % nk,
                              "ms_per_call_host_buffers": t_bow * 1e3, "distinct_words": int(len(np.unique(bw)))},
            "bow_vector": {"workload": "BowVector of %d features: per-word sums in feature order + L1 norm, doubles (DBoW2 transform / normalize)" % nk,
                           "ms_per_call_host_buffers": t_bvec * 1e3, "words": int(len(vw))},
            "bow_batch": {"workload": "%d keyframes x %d features per call: tree descent, then BowVectors (one CTA per keyframe)" % (NB, nk),
                          "transform_ms_per_call_host_buffers": t_btr * 1e3, "vector_ms_per_call_host_buffers": t_bvb * 1e3,
                          "keyframes_per_s": NB / (t_btr + t_bvb)},
            "bow_similar": {"workload": "getBowSimilar (bow_index.cpp:95-176): one query against %d stored keyframes of 750..1700 words" % NKF,
                            "ms_per_call_host_buffers": t_sim * 1e3, "keyframes_scored_per_s": NKF / t_sim, "candidates": int(len(sim_kf))},
            "match_for_loop_closures_bow": {"workload": "%d x %d features in DBoW2 node buckets (keyframe_matcher.cpp:65-146)" % (nq, nk),
                                            "ms_per_call_host_buffers": t_mbow * 1e3, "matches": n_mbow}}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-cv2", action="store_true", help="leave out the cv2 (OpenCV SIMD) columns of the CPU baseline")
    ap.add_argument("--config5-full", action="store_true",
                    help="run BASELINE configs[4] in full: all 49 995 000 keyframe pairs of 10 000 keyframes, block-sharded over the ranks")
    ap.add_argument("--skip-extra-configs", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling aid: leave out the host-buffer (chunked) pass")
    ap.add_argument("--pipe-chunk", type=int, default=0, help="frames per pipeline chunk of sg_extract (0: library default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local, td = dist_setup(args.gpus)
    import slam_module_b200 as sm
    from slam_module_b200 import slamgpu

    if slamgpu.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libslamgpu has no CPU fallback")
    hbm_peak, sm_max_mhz, peak_src = peaks()

    ctx = slamgpu.Context(W, H, levels=LEVELS, scale_factor=FACTOR, max_keypoints=MAXKP, max_frames=IN_FLIGHT * FRAMES, device=local)
    frame_bytes = W * H
    # ---- inputs: rank-specific frames, resident in HBM (N_BATCHES rotating batches) ---------------------
    host_batches, dev_batches = [], []
    for b in range(N_BATCHES):
        pin = slamgpu.PinnedArray((FRAMES, H, W), np.uint8)
        pin.array[...] = make_frames(FRAMES, 10000 + 1000 * rank + 100 * b)
        host_batches.append(pin)
        dev_batches.append(ctx.device_buffer(FRAMES * frame_bytes).upload(pin.array))

    def step_device(i):
        ctx.extract_device(dev_batches[i % N_BATCHES].ptr, W, frame_bytes, FRAMES)

    # ---- device-resident throughput ----------------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    ctx.synchronize()
    # clocks ramp up over the first tens of milliseconds of load: keep warming (untimed) until 150 ms have passed
    t_w, extra_warm = time.perf_counter(), 0
    while time.perf_counter() - t_w < 0.15:
        step_device(extra_warm)
        ctx.synchronize()
        extra_warm += 1
    counts, _ = ctx.extract_download(FRAMES, only_counts=True)
    kp_per_frame = float(counts.mean())
    sampler = ClockSampler(local)
    launches0 = ctx.launch_count()
    ctx.synchronize()
    barrier_max(td, local, 0.0)                      # barrier
    ctx.timer_start()
    for i in range(args.steps):
        step_device(args.warmup + i)
    ms = ctx.timer_stop()                            # records the end event and synchronises
    ms = barrier_max(td, local, ms)
    launches = ctx.launch_count() - launches0
    # per-stage event times: a second pass with profiling on (the stages then run on ONE stream; the timed pass above
    # overlaps independent slices of the batch on several streams, which would blur the per-kernel durations)
    ctx.set_profiling(True)
    for i in range(max(3, min(args.steps, 10))):
        step_device(i)
    stage = ctx.stage_ms()
    ctx.set_profiling(False)
    value = world * FRAMES * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer entry point (H2D + kernels + D2H) -----------------------------
    out_arrs, out_struct = ctx._alloc_out(FRAMES, pinned=True)   # pinned: the D2H copies overlap the kernels
    # what detectAndExtract returns (orb_extractor.hpp:16-20): position, angle, octave, descriptor, track id per keypoint.
    # The integer level coordinates are a test convenience of the library: not requested, not copied.
    E2E_SKIP = ("lvl_x", "lvl_y")
    for k in E2E_SKIP:
        setattr(out_struct, k, None)
    if args.pipe_chunk:
        ctx.set_pipeline_chunk(args.pipe_chunk)
    lib = slamgpu.lib()
    import ctypes as C

    def step_host(i):
        hb = host_batches[i % N_BATCHES].array
        ctx._check(lib.sg_extract(ctx._h, hb.ctypes.data, W, frame_bytes, FRAMES, None, None, None, C.byref(out_struct)))

    # raw pinned host->device copy of one batch (one cudaMemcpyAsync, all ranks at once): the ceiling of any host-buffer path
    h2d_gbs = None
    if not args.skip_e2e:
        scratch = ctx.device_buffer(FRAMES * frame_bytes)
        scratch.upload(host_batches[0].array)
        barrier_max(td, local, 0.0)
        t0 = time.perf_counter()
        for i in range(20):
            scratch.upload(host_batches[i % N_BATCHES].array)
        h2d_s = barrier_max(td, local, time.perf_counter() - t0)
        h2d_gbs = 20 * FRAMES * frame_bytes / h2d_s / 1e9        # per rank, while every rank copies
        scratch.free()
    e2e_steps = min(max(args.steps, 20), 200)      # its own step count (reported in e2e.steps): long enough to amortise the fill / drain of the batches in flight
    e2e_value = None
    if not args.skip_e2e:
        step_host(0)
        ctx.synchronize()
        barrier_max(td, local, 0.0)
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            step_host(1 + i)
        ctx.synchronize()
        e2e_s = barrier_max(td, local, time.perf_counter() - t0)
        e2e_value = world * FRAMES * e2e_steps / e2e_s
    d2h_bytes = sum(int(a.nbytes) for k, a in out_arrs.items() if k not in E2E_SKIP)
    # ... and the streaming form of the same call: IN_FLIGHT batches in flight on disjoint ranges of the context's frame
    # slots, so the copies of batch i+1 run under the kernels and the copy-out of batch i.  Every step still moves
    # its inputs H2D and its results D2H inside the timed region; the region ends when the last batch is on the host.
    e2e_stream = None
    if not args.skip_e2e:
        out2 = [(out_arrs, out_struct)] + [ctx.alloc_outputs(FRAMES, pinned=True) for _ in range(IN_FLIGHT - 1)]
        for _, ks_ in out2[1:]:
            for k in E2E_SKIP:
                setattr(ks_, k, None)
        tickets = [None] * IN_FLIGHT

        def run_stream(n, first):
            for i in range(n):
                s_ = i % IN_FLIGHT
                if tickets[s_] is not None:
                    ctx.extract_wait(tickets[s_])
                tickets[s_] = ctx.extract_submit(host_batches[(first + i) % N_BATCHES].array, s_ * FRAMES, out2[s_][1])
            for t in range(IN_FLIGHT):
                if tickets[t] is not None:
                    ctx.extract_wait(tickets[t])
                    tickets[t] = None
        run_stream(IN_FLIGHT, 0)
        ctx.synchronize()
        barrier_max(td, local, 0.0)
        t0 = time.perf_counter()
        run_stream(e2e_steps, IN_FLIGHT)
        e2e_stream_s = barrier_max(td, local, time.perf_counter() - t0)
        e2e_stream = world * FRAMES * e2e_steps / e2e_stream_s
        assert int(out2[1][0]["count"].min()) > 0

    # ---- matching (configs[2]) ---------------------------------------------------------------------------
    rngm = np.random.default_rng(77 + rank)
    dA, aA, dB, aB = sm.synth.correlated_descriptors(MATCH_N, 500 + rank)
    sets_d, sets_a = [dA, dB], [aA, aB]
    for s in range(MATCH_SETS - 2):
        d = sets_d[s % 2].copy()
        flip = rngm.integers(0, 256, (MATCH_N, 12))
        for k in range(12):
            d[np.arange(MATCH_N), flip[:, k] >> 5] ^= (np.uint32(1) << (flip[:, k] & 31).astype(np.uint32))
        perm = rngm.permutation(MATCH_N)
        sets_d.append(d[perm])
        sets_a.append(((sets_a[s % 2] + rngm.normal(0, 3, MATCH_N)) % 360).astype(np.float32)[perm])
    db = slamgpu.DescriptorDB(ctx, np.stack(sets_d), np.stack(sets_a))
    pairs = rngm.integers(0, MATCH_SETS, (MATCH_PAIRS, 2)).astype(np.int32)
    d_pairs = ctx.device_buffer(pairs.nbytes).upload(pairs)
    d_counts = ctx.device_buffer(4 * MATCH_PAIRS)
    for _ in range(3):
        db.match_pairs_device(d_pairs.ptr, MATCH_PAIRS, d_counts.ptr)
    ctx.synchronize()
    m_launch0 = ctx.launch_count()
    barrier_max(td, local, 0.0)
    m_steps = max(3, min(args.steps, 10))
    ctx.timer_start()
    for _ in range(m_steps):
        db.match_pairs_device(d_pairs.ptr, MATCH_PAIRS, d_counts.ptr)
    m_ms = barrier_max(td, local, ctx.timer_stop())
    m_launches = ctx.launch_count() - m_launch0
    # per-kernel event times: separate pass with profiling on (one stream; the timed pass overlaps sub-chunks)
    ctx.set_profiling(True)
    for _ in range(3):
        db.match_pairs_device(d_pairs.ptr, MATCH_PAIRS, d_counts.ptr)
    m_stage = ctx.stage_ms()
    ctx.set_profiling(False)
    mcounts = d_counts.download(np.uint32, MATCH_PAIRS)
    pair_dists = float(MATCH_PAIRS) * MATCH_N * MATCH_N
    match_value = world * pair_dists * m_steps / (m_ms * 1e-3)
    # host-buffer matching call (pairs up, counts + match rows down)
    E2E_PAIRS = min(1024, MATCH_PAIRS)
    pin_n, pin_m = slamgpu.PinnedArray((E2E_PAIRS,), np.uint32), slamgpu.PinnedArray((E2E_PAIRS, MATCH_N), np.int32)
    n_host, m_host = db.match_pairs(pairs[:E2E_PAIRS], out=(pin_n.array, pin_m.array))            # warm-up (scratch growth)
    barrier_max(td, local, 0.0)
    t0 = time.perf_counter()
    for _ in range(3):
        n_host, m_host = db.match_pairs(pairs[:E2E_PAIRS], out=(pin_n.array, pin_m.array))
    match_e2e_s = barrier_max(td, local, time.perf_counter() - t0)
    match_e2e = world * 3 * E2E_PAIRS * MATCH_N * MATCH_N / match_e2e_s
    assert np.array_equal(n_host, mcounts[:E2E_PAIRS])
    popc_peak, _ = ctx.microbench_popc()
    clocks = sampler.stop()
    extras = None
    if not args.skip_extra_configs:
        from slam_module_b200 import sharding as sh
        extras = extra_configs(slamgpu, sm, sh, rank, world, local, barrier_max, td, config5_full=args.config5_full)

    # ---- CPU baseline beside it (rank 0, N == 1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        be, po, cpu_kind, cpu_what = reference_backend()
        cores = os.cpu_count() or 1
        sample = FRAMES
        cpu_extract_baseline(be, host_batches[0].array[:2 * cores], cores)                # warm-up: heap growth, thread start
        v, secs, _ = cpu_extract_baseline(be, host_batches[0].array[:sample], cores)
        v2, secs2, _ = cpu_extract_baseline(be, host_batches[1].array[:sample], cores)   # second pass: take the better
        if v2 > v:
            v, secs = v2, secs2
        port_v = None
        if be is not po:                                                                  # the scalar port beside it
            cpu_extract_baseline(po, host_batches[0].array[:2 * cores], cores)
            port_v, _, _ = cpu_extract_baseline(po, host_batches[0].array[:sample], cores)
        n_cpu_pairs = 4 * max(8, cores)
        be.bench_match(np.stack(sets_d[:8]), np.stack(sets_a[:8]), pairs[:cores] % 8, cores)
        msec, _ = be.bench_match(np.stack(sets_d[:8]), np.stack(sets_a[:8]), pairs[:n_cpu_pairs] % 8, cores)
        cv2_cols = None
        if not args.skip_cv2:
            cv2_cols = cv2_baseline(host_batches[0].array[:max(64, 4 * cores)],
                                    [np.ascontiguousarray(x).view(np.uint8).reshape(MATCH_N, 32) for x in sets_d[:8]],
                                    pairs[:2 * cores] % 8, cores)
        # parity of this very run against the oracle (checker role): keypoint sets, angles, descriptors, match indices
        N_CHECK = 8
        pp = po.make_params(W, H, levels=LEVELS, scale_factor=FACTOR, max_keypoints=MAXKP)
        got = ctx.detect_and_extract(host_batches[0].array[:N_CHECK])
        kp_equal, max_dang, desc_bad, n_kp = True, 0.0, 0, 0
        feats = []
        for f in range(N_CHECK):
            ref = po.extract(pp, host_batches[0].array[f])
            gf = got[f]
            same = gf["n"] == ref["n"] and all(np.array_equal(gf[k], ref[k]) for k in ("x", "y", "octave"))
            kp_equal = kp_equal and same
            if same:
                n_kp += ref["n"]
                if ref["n"]:
                    da = np.abs(gf["angle"].astype(np.float64) - ref["angle"].astype(np.float64))
                    max_dang = max(max_dang, float(np.deg2rad(np.minimum(da, 360.0 - da)).max()))
                desc_bad += int((gf["desc"] != ref["desc"]).any(axis=1).sum())
            feats.append(ref)
        match_bad = 0
        for f in range(0, N_CHECK, 2):
            a_, b_ = feats[f], feats[f + 1]
            ng, mg = ctx.match_bruteforce(a_["desc"], a_["angle"], b_["desc"], b_["angle"])
            nr, mr = po.match_bruteforce(a_["desc"], a_["angle"], b_["desc"], b_["angle"])
            match_bad += int((mg != mr).sum()) + int(ng != nr)
        parity = {"frames_checked": N_CHECK, "keypoints_checked": int(n_kp), "keypoint_sets_equal": bool(kp_equal),
                  "max_angle_diff_rad": max_dang, "descriptor_mismatches": int(desc_bad),
                  "keyframe_pairs_checked": N_CHECK // 2, "match_index_mismatches": int(match_bad),
                  "checker": "oracle port (oracle/), same frames as the timed batch"}
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": cpu_kind, "parity": parity,
               "sample": "%s on the %d frames of one step, sharded over %d host threads (%.1f s per pass, best of 2)"
                         % (cpu_what, sample, cores, secs),
               "port_value": port_v, "cv2": cv2_cols,
               "matching_value": n_cpu_pairs * MATCH_N * MATCH_N / msec, "matching_unit": "descriptor-pair distances/s",
               "matching_sample": "%d keyframe pairs over %d threads (%.1f s)" % (n_cpu_pairs, cores, msec)}

    if rank == 0:
        # roofline of the dominant kernel of the step (per-stage CUDA events, averaged over the timed steps)
        stage_bytes = {"pyramid": ALGO_BYTES_PYRAMID * FRAMES, "fast": ALGO_BYTES_FAST * FRAMES}
        dom = max(("pyramid", "fast", "distribute", "describe"), key=lambda k: stage[k])
        stages = {}
        for k in ("pyramid", "fast", "distribute", "describe"):
            e = {"ms": stage[k], "share": stage[k] / max(sum(stage[s] for s in ("pyramid", "fast", "distribute", "describe")), 1e-9)}
            if k in stage_bytes:
                e["achieved_gbs"] = stage_bytes[k] / (stage[k] * 1e-3) / 1e9
                e["frac_of_hbm"] = e["achieved_gbs"] / hbm_peak
            stages[k] = e
        roof_stage = dom if dom in stage_bytes else "pyramid"
        roofline = {"kernel": {"pyramid": "pyr_fast_kernel (8 launches: blur of level 0 + 7 fused resize+blur levels; traffic: the level-1 launch)",
                               "fast": "fast_cells_kernel (1 launch per step)"}[roof_stage],
                    "note": "shared-memory and integer-issue bound, not HBM bound: l1tex ~77%, ALU pipe ~69%, issue slots ~68% busy at 5 CTAs per SM (profiles/, newest ncu_full_*.md)",
                    "bound": "hbm", "achieved": stages[roof_stage]["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": stages[roof_stage]["frac_of_hbm"], "traffic": traffic_of(roof_stage)[0], "traffic_source": traffic_of(roof_stage)[1],
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch_group": stage_bytes[roof_stage], "dominant_stage_by_time": dom}
        topk_ms = m_stage["match_topk"]
        popc_achieved = 8.0 * min(MATCH_PAIRS, MATCH_PAIRS) * MATCH_N * MATCH_N / (topk_ms * 1e-3) if topk_ms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "frames_per_step_per_gpu": FRAMES, "levels": LEVELS, "scale_factor": FACTOR, "max_keypoints": MAXKP,
                       "keypoints_per_frame": kp_per_frame, "sharding": "frames by rank, no collective", "extra_untimed_warmup_steps": extra_warm,
                       "l2": "inputs larger than L2: %d rotating batches x %.1f MB" % (N_BATCHES, FRAMES * frame_bytes / 1e6)},
            "e2e": {"value": e2e_stream, "unit": "frames/s", "h2d_bytes_per_step": FRAMES * frame_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                    "api": "sg_extract_submit / sg_extract_wait (pinned host buffers; every batch is copied H2D, processed and copied "
                           "D2H; %d batches in flight, chunked H2D / kernels / D2H pipeline inside each)" % IN_FLIGHT,
                    "h2d_ceiling": None if h2d_gbs is None else {
                        "gbs_per_gpu_while_all_ranks_copy": h2d_gbs, "aggregate_gbs": h2d_gbs * world,
                        "frames_per_s": world * h2d_gbs * 1e9 / frame_bytes,
                        "frac": (e2e_stream / (world * h2d_gbs * 1e9 / frame_bytes)) if e2e_stream else None,
                        "what": "20 raw pinned host->device copies of one 78.6 MB batch per rank, all ranks concurrently: the "
                                "ceiling of any path that takes host frames (307 KB per frame over PCIe)"},
                    "single_call_value": e2e_value,
                    "single_call_api": "sg_extract (one synchronous call per batch, same pipeline, nothing in flight across calls)"},
            "gpu_launches": int(launches),
            "matching_value": match_value, "matching_unit": "descriptor-pair distances/s", "matching_e2e_value": match_e2e,
            "stereo_pairs_per_s": (extras or {}).get("config3_stereo_stream", {}).get("stereo_pairs_per_s"),
            "single_frame_latency_us": (extras or {}).get("config0_single_frame", {}).get("latency_us_host_call_median"),
            "stages": stages,
            "roofline": roofline,
            "matching": {"metric": "Hamming matches/sec (2000x2000 per keyframe pair, ratio 0.8, angle histogram; configs[2])",
                         "value": match_value, "unit": "descriptor-pair distances/s",
                         "keyframe_pairs_per_s": world * MATCH_PAIRS * m_steps / (m_ms * 1e-3),
                         "pairs_per_step_per_gpu": MATCH_PAIRS, "steps": m_steps, "ms_per_step": m_ms / m_steps,
                         "mean_matches_per_pair": float(mcounts.mean()), "gpu_launches": int(m_launches),
                         "stages_ms": {"topk": m_stage["match_topk"], "resolve": m_stage["match_resolve"]},
                         "e2e": {"value": match_e2e, "unit": "descriptor-pair distances/s",
                                 "api": "sg_match_pairs (pinned host buffers: pair list up, counts and match rows down), %d keyframe pairs per call" % E2E_PAIRS},
                         "roofline": {"kernel": "hamming_topk_kernel", "bound": "int-popc", "achieved": popc_achieved,
                                      "peak": popc_peak, "unit": "POPC.b32/s",
                                      "frac": (popc_achieved / popc_peak) if popc_achieved else None,
                                      "peak_source": "measured live: dependent-free POPC micro-benchmark on this GPU",
                                      "algorithmic_ops_per_pair": 8, "issued_popc_per_pair": 4,
                                      "frac_of_issued": (0.5 * popc_achieved / popc_peak) if popc_achieved else None,
                                      "note": "achieved counts the 8 POPC.b32 per pair of the reference algorithm; the kernel "
                                              "compresses the 8 XOR words with carry-save adders (LOP3) and issues 4 POPC per "
                                              "pair, so the algorithmic rate can exceed the POPC issue peak"}},
            "clocks": clocks,
        }
        if extras is not None:
            line["other_configs"] = extras
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    db.close()
    ctx.close()
    if td is not None:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
