"""CPU tests of the boundary and the host logic: the C-ABI library loads and exports every symbol
include/slamgpu.h declares, fails loudly without a GPU, and its host-side pieces (angle-bin order,
shard bookkeeping of bench.py) agree with the oracle.  No GPU compute is called here."""
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_exported(slamgpu):
    header = (ROOT / "include" / "slamgpu.h").read_text()
    declared = sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    L = slamgpu.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(slamgpu.ABI_SYMBOLS) == declared
    assert L.sg_abi_version() == 1


def test_header_compiles_as_c():
    src = '#include "slamgpu.h"\nint main(void){ sg_params p; (void)p; return SG_OK; }\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-x", "c", "-", "-fsyntax-only"],
                       input=src.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()


def test_no_cpu_fallback(slamgpu):
    if slamgpu.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(slamgpu.SlamGpuError) as e:
        slamgpu.Context(640, 480)
    assert e.value.code == slamgpu.SG_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_touch_oracle():
    """The product path must not import, link or call anything under oracle/."""
    for f in list((ROOT / "slam-module_b200").rglob("*.py")) + list((ROOT / "slam-module_b200").rglob("*.cu")) \
            + list((ROOT / "slam-module_b200").rglob("*.cpp")) + list((ROOT / "slam-module_b200").rglob("*.h")) \
            + list((ROOT / "slam-module_b200").rglob("*.hpp")) + list((ROOT / "slam-module_b200").rglob("Makefile")):
        text = f.read_text()
        assert "pyoracle" not in text and "orb_oracle" not in text and "liborb_oracle" not in text, f
    r = subprocess.run(["ldd", str(ROOT / "slam-module_b200" / "csrc" / "libslamgpu.so")], capture_output=True, text=True)
    assert "oracle" not in r.stdout


def test_angle_bin_order_matches_std_sort(slamgpu, oracle):
    rng = np.random.default_rng(1)
    for trial in range(3000):
        kind = trial % 4
        if kind == 0:
            sizes = rng.integers(0, 4, 30)
        elif kind == 1:
            sizes = np.zeros(30, np.int64)
            sizes[:13] = rng.integers(0, 6, 13)          # what the matcher produces: bins 0..12 only
        elif kind == 2:
            sizes = rng.integers(0, 1000, 30)
        else:
            sizes = np.repeat(rng.integers(0, 3, 6), 5)[rng.permutation(30)]
        sizes = sizes.astype(np.uint32)
        assert np.array_equal(slamgpu.angle_bin_order(sizes), oracle.bin_order(sizes)), sizes
        assert np.array_equal(slamgpu.angle_bin_order(sizes, 0), oracle.bin_order_heap(sizes)), sizes


def test_angle_bin_matches_reference(slamgpu, oracle, golden):
    for d in np.concatenate([np.linspace(-359.9, 359.9, 2001), [15.0, 45.0, 75.0, 345.0, -15.0, 359.9, 0.0]]).astype(np.float32):
        assert slamgpu.angle_bin(float(d)) == oracle.angle_bin(float(d))
    assert slamgpu.angle_bin(359.9) == 12 and slamgpu.angle_bin(15.0) == 0 and slamgpu.angle_bin(45.0) == 2


def test_synthetic_inputs_are_deterministic(synth):
    a = synth.frame(640, 480, 1000)
    b = synth.frame(640, 480, 1000)
    assert np.array_equal(a, b) and a.dtype == np.uint8 and a.shape == (480, 640)
    assert not np.array_equal(a, synth.frame(640, 480, 1001))


def test_integration_doc_covers_every_entry_point(slamgpu):
    """INTEGRATION.md names, for every exported entry point, the reference interface it replaces (or says it is a
    harness helper); `sg_x[_device]` shorthands are expanded."""
    import re
    text = (ROOT / "INTEGRATION.md").read_text()
    named = set(re.findall(r"sg_[a-z0-9_]+", text))
    for m in re.finditer(r"(sg_[a-z0-9_]+)\[(_[a-z_]+)\]", text):
        named.add(m.group(1) + m.group(2))
    for m in re.finditer(r"(sg_[a-z0-9_]+)\[(_[a-z]+)\[(_[a-z]+)\]\]", text):
        named.add(m.group(1) + m.group(2))
        named.add(m.group(1) + m.group(2) + m.group(3))
    missing = [s for s in slamgpu.ABI_SYMBOLS if s not in named]
    assert not missing, missing

