#!/usr/bin/env python3
"""Executed warp instructions of one captured launch, aggregated over SOURCE-LINE RANGES (phases) of a .cu file:
    python tools/phase_mix.py <rep> <file.cu> name:lo-hi [name:lo-hi ...] [--px N] [--launch K]
Lines outside every range are reported as 'other'.  Instructions of inlined helpers count for the phase that calls them."""
import csv, subprocess, sys
args = [a for a in sys.argv[1:] if not a.startswith("--")]
rep, cu = args[0], args[1]
px = float(sys.argv[sys.argv.index("--px") + 1]) if "--px" in sys.argv else None
launch = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else 0
n_launches = int(sys.argv[sys.argv.index("--launches") + 1]) if "--launches" in sys.argv else 1   # launches captured in the report
if px is not None:
    args = [a for a in args if a != str(int(px)) and a != sys.argv[sys.argv.index("--px") + 1]]
if "--launch" in sys.argv:
    args = [a for a in args if a != sys.argv[sys.argv.index("--launch") + 1]]
if "--launches" in sys.argv:
    args = [a for a in args if a != sys.argv[sys.argv.index("--launches") + 1]]
ranges = []
for a in args[2:]:
    name, r = a.rsplit(":", 1)
    lo, hi = r.split("-")
    ranges.append((name, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# The page lists, per captured launch, one block per source file: the kernel's .cu first (many lines), then short blocks
# for inlined helpers (tma.cuh, the CUDA intrinsics headers).  Every source-line row is followed by its SASS rows (address,
# opcode, counters).  An inlined instruction appears under the helper's line AND, usually, under its call site in the .cu;
# source rows holding string literals (asm volatile(...)) break the CSV quoting, so only the SASS rows are counted, each
# ADDRESS once: by its .cu line where it has one, else by the phase of the nearest attributed address below it.
hdr, blocks, cur, line = None, [], None, None
for r in rows:
    if r and r[0] == "Line No":
        hdr, cur, line = r, {"lines": 0, "sass": []}, None
        blocks.append(cur)
        continue
    if hdr is None or not r:
        continue
    if r[0].isdigit():
        line = int(r[0])
        cur["lines"] += 1
    elif r[0] == "" and len(r) > 3 and r[2].startswith("0x"):
        back = len(hdr) - hdr.index("Instructions Executed")
        cur["sass"].append((int(r[2], 16), int(r[-back] or 0), line))
starts = [i for i, b in enumerate(blocks) if b["lines"] >= 20]       # a launch starts at every block of 20 or more lines
lo = starts[launch]
hi = starts[launch + 1] if launch + 1 < len(starts) else len(blocks)
count, where = {}, {}
for addr, n, ln in blocks[lo]["sass"]:
    count[addr] = n
    for name, a, b in ranges:
        if a <= ln <= b:
            where[addr] = name
            break
    # (a line of the .cu outside every range -- a __device__ helper defined above the kernel -- is treated like an
    #  inlined header helper: the phase of its neighbours)
for b in blocks[lo + 1:hi]:
    for addr, n, ln in b["sass"]:
        count.setdefault(addr, n)
order = sorted(count)
prev = None
for addr in order:                                                   # helper-only instructions: nearest attributed address below
    if addr in where:
        prev = where[addr]
    elif prev is not None:
        where[addr] = prev
nxt = None
for addr in reversed(order):
    if addr in where:
        nxt = where[addr]
    else:
        where[addr] = nxt or "other"
tot = sum(count.values())
agg = {}
for addr, n in count.items():
    agg[where[addr]] = agg.get(where[addr], 0) + n
print("launch %d: %d warp instructions%s" % (launch, tot, (" = %.1f thread-instructions per px" % (tot * 32 / px)) if px else ""))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print("  %-34s %12d  %5.1f%%%s" % (k, v, 100.0 * v / tot, ("  %5.1f /px" % (v * 32 / px)) if px else ""))
