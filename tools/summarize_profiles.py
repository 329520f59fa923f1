#!/usr/bin/env python3
"""Turn the ncu outputs of tools/profile.sh (gpurun_out/*_<tag>.*) into the committed summaries under
profiles/: the launch list (per-kernel totals and shares) and the key counters of every --set full
capture.  Runs in the build container (ncu -i works without a GPU).

    python tools/summarize_profiles.py r1
"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
dst = ROOT / "profiles"
dst.mkdir(exist_ok=True)

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]

# ---- launch list ------------------------------------------------------------------------------------
lf = OUT / ("launches_%s.csv" % tag)
if lf.exists():
    rows = [r for r in csv.reader(open(lf)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values())
    with open(dst / ("launches_%s.md" % tag), "w") as f:
        f.write("# ncu launch list, %s\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` on "
                "`python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --skip-extra-configs --skip-e2e` (cold-cache, serialised: compare shares).\n\n"
                "| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n" % tag)
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.1f | %.3f |\n" % (k, n, v / 1e3, v / 1e3 / n, v / tot))
    (dst / ("launches_%s.csv" % tag)).write_text(lf.read_text())
    print("wrote", dst / ("launches_%s.md" % tag))

# ---- full captures ----------------------------------------------------------------------------------
with open(dst / ("ncu_full_%s.md" % tag), "w") as f:
    f.write("# ncu --set full captures, %s\n\nOne row block per captured launch; values straight from "
            "`ncu -i <rep> --page raw --csv`.\n" % tag)
    for rep in sorted(OUT.glob("prof_*_%s.ncu-rep" % tag)):
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            f.write("\n## %s\n\n| metric | value | unit |\n|---|---:|---|\n" % r[hdr.index("Kernel Name")].split("(")[0])
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write("| %s | %s | %s |\n" % (w, r[i], units[i]))
print("wrote", dst / ("ncu_full_%s.md" % tag))
