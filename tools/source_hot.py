"""Per-source-line executed warp instructions and stall samples of one kernel from an .ncu-rep (ncu --set full
--import-source on):  python tools/source_hot.py <rep> [top]"""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        if hdr is not None:
            break          # first captured launch only
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():      # per-source-line aggregate rows
        lines.append(r)
ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples")
tot_i = sum(int(r[ci] or 0) for r in lines); tot_s = sum(int(r[cs] or 0) for r in lines)
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
acc = 0
for r in sorted(lines, key=lambda r: -int(r[ci] or 0))[:top]:
    print("%5s %6.2f%% inst %6.2f%% smp  %s" % (r[0], 100 * int(r[ci] or 0) / tot_i, 100 * int(r[cs] or 0) / max(tot_s, 1), r[1].strip()[:110]))
