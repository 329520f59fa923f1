#!/usr/bin/env python3
"""Per-opcode executed-instruction mix of every kernel in an ncu report (source page, SASS view).
    python tools/sass_mix.py gpurun_out/prof_X.ncu-rep [px_per_launch]
Runs in the build container (ncu -i needs no GPU)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
px = [float(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else None   # per captured launch (or one value for all)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
kern, hdr, mix, stall = None, None, None, None
out = []
for row in csv.reader(raw.splitlines()):
    if not row:
        continue
    if row[0] == "Kernel Name":
        if kern:
            out.append((kern, mix, stall))
        kern, hdr, mix, stall = row[1], None, collections.Counter(), collections.Counter()
        continue
    if row[0] == "Address":
        hdr = row
        continue
    if hdr is None or len(row) < len(hdr):
        continue
    src = row[hdr.index("Source")].strip()
    toks = src.split()
    if toks and toks[0].startswith("@"):
        toks = toks[1:]
    op = toks[0].rstrip(";") if toks else "?"
    op = ".".join(op.split(".")[:2])
    mix[op] += int(row[hdr.index("Instructions Executed")] or 0)
    stall[op] += int(row[hdr.index("# Samples")] or 0)
if kern:
    out.append((kern, mix, stall))
dedup = []
for o in out:                       # the source page repeats every launch (SASS view, then source view)
    if not dedup or (dedup[-1][0], sum(dedup[-1][1].values())) != (o[0], sum(o[1].values())):
        dedup.append(o)
for n, (kern, mix, stall) in enumerate(dedup):
    tot = sum(mix.values())
    if tot == 0:
        continue
    st = sum(stall.values()) or 1
    p1 = (px[n] if n < len(px) else px[-1]) if px else None
    print("== %s: %d warp instructions%s" % (kern[:70], tot, ("  (%.1f thread-inst/px over %d px)" % (tot * 32 / p1, p1)) if p1 else ""))
    for op, n in mix.most_common(28):
        print("  %-22s %12d  %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * stall[op] / st))
