#!/usr/bin/env python3
"""Executed warp instructions of one captured launch, aggregated over SOURCE-LINE RANGES (phases) of a .cu file:
    python tools/phase_mix.py <rep> <file.cu> name:lo-hi [name:lo-hi ...] [--px N] [--launch K]
Lines outside every range are reported as 'other'.  Inlined helper lines (tma.cuh, ...) are attributed to 'helpers'."""
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
blocks, hdr, cur = [], None, None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        cur = []
        blocks.append(cur)
        continue
    if hdr and cur is not None and len(r) == len(hdr) and r[0].isdigit():
        cur.append(r)
# the source page lists, per captured launch, one block per source file: the kernel's .cu first (many lines), then a few
# short blocks for the inlined helpers (tma.cuh, crt headers) -- a launch starts at every block of 20 or more lines
starts = [i for i, b in enumerate(blocks) if len(b) >= 20]
lo = starts[launch]
hi = starts[launch + 1] if launch + 1 < len(starts) else len(blocks)
main_rows = set(id(r) for r in blocks[lo])
blk = [r for b in blocks[lo:hi] for r in b]
ci = hdr.index("Instructions Executed")
fi = hdr.index("File Path") if "File Path" in hdr else None
tot = sum(int(r[ci] or 0) for r in blk)
agg = {}
for r in blk:
    n = int(r[ci] or 0)
    line = int(r[0])
    where = "other"
    if id(r) not in main_rows:
        where = "inlined helpers (tma.cuh, ...)"
    else:
        for name, lo, hi in ranges:
            if lo <= line <= hi:
                where = name
                break
    agg[where] = agg.get(where, 0) + n
print("launch %d: %d warp instructions%s" % (launch, tot, (" = %.1f thread-instructions per px" % (tot * 32 / px)) if px else ""))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print("  %-34s %12d  %5.1f%%%s" % (k, v, 100.0 * v / tot, ("  %5.1f /px" % (v * 32 / px)) if px else ""))
