"""Generates tests/golden/traceback_<case>.npz: the reference's final traceback-stage results (BlastHSPResults after
Blast_RunTracebackSearch, via oracle/ref_driver.c) for the adversarial list cases of tests/cases.py.  Run from the
repository root where oracle/_ref/libblastref.so is built:  python tests/golden/make_traceback_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import cases                      # noqa: E402
from oracle import refdriver as R            # noqa: E402

for name in cases.TRACEBACK_LIST_CASES:
    task, cfgkw, vol, qs = cases.make_case(name)
    r = R.search(qs, vol, R.default_config(task, taps=R.TAP_TRACEBACK, prelim_only=0, **cfgkw))
    assert r["status"] == 0
    out = os.path.join(ROOT, "tests", "golden", f"traceback_{name}.npz")
    np.savez_compressed(out, tb_final=r["tb_final"], tb_ops=r["tb_ops"], prelim_final=r["final"],
                        gap_x_dropoff_final=np.int32(r["gap_x_dropoff_final"]))
    print(out, r["tb_final"].shape, r["tb_ops"].shape)
