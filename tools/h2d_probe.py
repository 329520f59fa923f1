#!/usr/bin/env python3
"""Raw pinned host<->device copy ceilings while EVERY rank copies at once (the feed limit of any host-buffer path).
Launch under torchrun like bench.py; rank 0 prints one JSON line.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py

Per rank and step: 78.6 MB up (256 frames of 640x480) and, in the bidirectional leg, 31.2 MB down (the keypoint arrays of
one extraction step) on a second stream -- the traffic of one bench.py e2e step.  Chunked like sg_extract_submit (128 frames
per cudaMemcpyAsync)."""
import json
import os
import time

import torch
import torch.distributed as td

UP, DOWN, CHUNKS, STEPS = 256 * 640 * 480, 31_193_088, 2, 40


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    h_up = [torch.empty(UP, dtype=torch.uint8).pin_memory() for _ in range(2)]
    d_up = torch.empty(UP, dtype=torch.uint8, device=dev)
    h_dn = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
    d_dn = torch.empty(DOWN, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier_max(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    def run(bidir, steps):
        c = UP // CHUNKS
        for i in range(steps):
            with torch.cuda.stream(s_up):
                for k in range(CHUNKS):
                    d_up[k * c:(k + 1) * c].copy_(h_up[i % 2][k * c:(k + 1) * c], non_blocking=True)
            if bidir:
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.synchronize()

    out = {"n_gpus": world, "up_bytes_per_step": UP, "down_bytes_per_step": DOWN, "steps": STEPS}
    for name, bidir in (("h2d_only", False), ("h2d_with_d2h", True)):
        run(bidir, 5)
        barrier_max(0.0)
        t0 = time.perf_counter()
        run(bidir, STEPS)
        dt = barrier_max(time.perf_counter() - t0)
        out[name] = {"h2d_gbs_per_gpu": STEPS * UP / dt / 1e9, "h2d_gbs_aggregate": world * STEPS * UP / dt / 1e9,
                     "d2h_gbs_aggregate": (world * STEPS * DOWN / dt / 1e9) if bidir else 0.0,
                     "frames_per_s_ceiling": world * STEPS * 256 / dt}
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
