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


# ---- full-batch evidence (tools/profile_r2.sh): DRAM traffic per step, opcode mix, instructions per pixel by phase ----
import json

W8, H8 = [640, 533, 444, 370, 309, 257, 214, 179], [480, 400, 333, 278, 231, 193, 161, 134]
PX = [w * h * 256 for w, h in zip(W8, H8)]
SUM_PX = sum(PX)
KP = 513000
PHASES = {}     # kernel -> (source file, pixels of the launch or None, ["phase:lo-hi", ...]), filled below


def marker_phases(cu, start, sections):
    """Source-line ranges of a kernel's phases, located by its own section comments (so edits do not stale them):
    `start` = text on the kernel's signature line, sections = [(phase name, text that opens it), ...] in file order; a
    phase runs to the line before the next one, the last to the closing brace of the kernel."""
    lines = (ROOT / "slam-module_b200" / "csrc" / cu).read_text().splitlines()
    def at(text, after):
        for i in range(after, len(lines)):
            if text in lines[i]:
                return i + 1
        raise SystemExit("%s: marker %r not found" % (cu, text))
    k = at(start, 0)
    pos, cur = [], k
    for name, text in sections:
        cur = at(text, cur) if text else k
        pos.append((name, cur))
    end = next(i + 1 for i in range(pos[-1][1], len(lines)) if lines[i] == "}")
    return ["%s:%d-%d" % (name, lo, (pos[i + 1][1] - 1) if i + 1 < len(pos) else end) for i, (name, lo) in enumerate(pos)]


PHASES["pyr_fast_kernel"] = ("pyramid.cu", None, marker_phases("pyramid.cu", "pyr_fast_kernel(const PyrArgs a", [
    ("setup, taps, TMA wait", None), ("resize", "auto resize_px"), ("blur-only load", "// blur only: the"),
    ("reflect-101", "// ---- reflect-101"), ("blur horizontal", "// ---- horizontal pass"),
    ("blur vertical + plane store", "// ---- vertical pass")]))
PHASES["fast_cells_kernel"] = ("detect.cu", SUM_PX, marker_phases("detect.cu", "fast_cells_kernel(const __grid_constant__", [
    ("setup+TMA", None), ("pair words", "// ---- pair words"), ("stage 1 (antipodal test, queue)", "// Two passes at most"),
    ("stage 2 (exact score)", "// ---- stage 2"), ("NMS", "// ---- cell-local NMS"), ("output", "if (nkeep == 0) return;")]))
PHASES["distribute_kernel"] = ("detect.cu", None, marker_phases("detect.cu", "constexpr int DIST_THREADS", [
    ("helpers (block scans, candidate store, quadrant counts)", None), ("initial nodes", "// ---- initial nodes"),
    ("processing order (rank loop)", "// ---- rounds"), ("children scan, stop point", "// D: children created"),
    ("new node table", "// F: new node table"), ("move candidates + next quadrant counts", "// G: move the candidates"),
    ("termination", "// H: termination"), ("strongest per node", "// ---- strongest candidate"),
    ("kernel entry", "distribute_kernel(const __grid_constant__")]))
PHASES["describe_kernel"] = ("describe.cu", None, marker_phases("describe.cu", "describe_kernel(const __grid_constant__", [
    ("tables, pattern", None), ("slot lookup", "const int groups_per_frame"), ("moments (TMA + IDP.4A)", "// ---- phase A"),
    ("angle, trig, outputs", "// ---- phase B"), ("rBRIEF (TMA + sampling)", "// ---- phase C")]))
traffic = {"tag": tag, "how": "ncu --set full --clock-control none on tools/full_batch_pass.py (256-frame single launches, profiling "
                              "layout of bench.py's stage timing); dram__bytes_read.sum + dram__bytes_write.sum per launch"}
mix_md = ["# SASS opcode mix and per-phase instruction budget, %s\n\nFull-batch launches (256 frames of 640x480, 8 levels, 2000 keypoints; "
          "2048 keyframe pairs), `ncu --set full --import-source on`, read with `tools/sass_mix.py` / `tools/phase_mix.py`.  "
          "Thread-instructions per pixel = 32 x warp instructions / pixels of the level(s) the launch covers.\n" % tag]
for rep in sorted(OUT.glob("prof_*_%s.ncu-rep" % tag)):
    kern = rep.name[len("prof_"):-len("_%s.ncu-rep" % tag)]
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr = rows[0]
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    units = rows[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = []
    for r in rows[2:]:
        b = float(r[ir].replace(",", "")) * scale.get(units[ir], 1) + float(r[iw].replace(",", "")) * scale.get(units[iw], 1)
        per.append({"bytes": b, "duration_us": float(r[it].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[it], 1)})
    if kern == "pyr_fast_kernel":
        traffic["pyr_fast_kernel (all 8 launches)"] = {"bytes_per_launch": sum(p["bytes"] for p in per), "launches": len(per),
                                                       "per_launch": per, "algorithmic_bytes": 2208264 * 256}
    else:
        traffic[kern] = {"bytes_per_launch": per[0]["bytes"], "duration_us": per[0]["duration_us"]}
        if kern == "fast_cells_kernel":
            traffic[kern]["algorithmic_bytes"] = 950532 * 256
    # opcode mix
    px_arg = ",".join(str(v) for v in PX) if kern == "pyr_fast_kernel" else (str(SUM_PX) if kern == "fast_cells_kernel" else None)
    cmd = [sys.executable, str(ROOT / "tools" / "sass_mix.py"), str(rep)] + ([px_arg] if px_arg else [])
    mix = subprocess.run(cmd, capture_output=True, text=True).stdout
    blocks = mix.split("== ")[1:]
    keep = blocks[:2] if kern == "pyr_fast_kernel" else blocks[:1]      # level 0 (blur only) and level 1 (resize + blur)
    mix_md.append("\n## %s\n" % kern)
    if kern == "pyr_fast_kernel":
        mix_md.append("Warp instructions of the 8 launches of one step (level 0 = blur only, levels 1-7 = resize + blur):\n")
        for n, b in enumerate(blocks):
            mix_md.append("* level %d: %s" % (n, b.splitlines()[0].split(": ", 1)[1]))
        mix_md.append("")
    for b in keep:
        lines = b.splitlines()
        mix_md.append("```\n" + lines[0] + "\n" + "\n".join(lines[1:19]) + "\n```")
    if kern in PHASES:
        cu, px, ranges = PHASES[kern]
        launches = [(1, PX[1])] if kern == "pyr_fast_kernel" else [(0, px)]
        for launch, lpx in launches:
            cmd = [sys.executable, str(ROOT / "tools" / "phase_mix.py"), str(rep), cu] + ranges + ["--launch", str(launch)]
            if kern == "pyr_fast_kernel":
                cmd += ["--launches", "8"]
            if lpx:
                cmd += ["--px", str(lpx)]
            ph = subprocess.run(cmd, capture_output=True, text=True).stdout
            mix_md.append("Per phase (source-line ranges of `csrc/%s`)%s:\n\n```\n%s```" % (cu, " of the level-1 launch" if kern == "pyr_fast_kernel" else "", ph))
(dst / ("traffic_%s.json" % tag)).write_text(json.dumps(traffic, indent=1))
(dst / ("sass_mix_%s.md" % tag)).write_text("\n".join(mix_md) + "\n")
print("wrote", dst / ("traffic_%s.json" % tag), dst / ("sass_mix_%s.md" % tag))
