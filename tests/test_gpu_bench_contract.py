"""bench.py's JSON line carries every key of the measurement contract (metric / value / e2e / roofline / cpu_baseline /
clocks / gpu_launches), for our arm and for the reference arm (the oracle port on the host cores)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _line(args):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py")] + args, capture_output=True, text=True, cwd=str(ROOT), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "bench.py must print exactly one line on stdout"
    return json.loads(lines[0])


@pytest.mark.gpu
def test_bench_line_contract():
    d = _line(["--steps", "3", "--warmup", "3", "--skip-extra-configs"])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["scaling"] == "weak" and d["dtype"] == "u8"
    assert d["vs_baseline"] is None and d["higher_is_better"] is True and d["unit"] == "frames/s"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 1e4 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["unit"] == "frames/s" and e["h2d_bytes_per_step"] == 256 * 640 * 480 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]                      # the copies are inside the timed region
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    par = c["parity"]                                       # the north star asks for the mismatch counts to be reported
    assert par["keypoint_sets_equal"] and par["descriptor_mismatches"] == 0 and par["match_index_mismatches"] == 0
    assert par["max_angle_diff_rad"] <= 1e-4 and par["keypoints_checked"] > 1000
    assert d["clocks"]["sm_mhz"] > 0 and isinstance(d["clocks"]["reasons"], list)
    m = d["matching"]
    assert m["roofline"]["unit"] == "POPC.b32/s" and m["value"] > 1e10


@pytest.mark.gpu
def test_bench_reference_arm_contract():
    d = _line(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
