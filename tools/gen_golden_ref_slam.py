#!/usr/bin/env python3
"""Write tests/golden/golden_ref_slam.npz from THE REFERENCE'S OWN SOURCES compiled verbatim (oracle/_ref/libref_slam.so,
built from /root/reference by `make -C oracle ref`): every case of tests/refcases.py / tests/test_reference_parity.py run
with backend "ref".  Run in the BUILD container only (the reference tree does not exist on the GPU box).

    python tools/gen_golden_ref_slam.py
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import refcases as rc  # noqa: E402
import test_reference_parity as trp  # noqa: E402


def main():
    assert rc.pr.available(), "needs /root/reference (make -C oracle ref)"
    out = {}
    for name, (fn, _) in sorted(trp.CASES.items()):
        for k, v in fn("ref").items():
            out["%s/%s" % (name, k)] = np.asarray(v)
    for k, v in rc.case_bow("ref", Path(tempfile.mkdtemp())).items():
        out["bow/%s" % k] = np.asarray(v)
    path = ROOT / "tests" / "golden" / "golden_ref_slam.npz"
    np.savez_compressed(path, **out)
    print(path.name, path.stat().st_size, "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
