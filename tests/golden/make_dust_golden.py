"""Writes tests/golden/dust_golden.json: the reference's symmetric DUST (oracle/_ref/libdustref.so, built from
/root/reference by `make -C oracle dust`) on the seeded sequences of tests/test_dust.py::dust_cases, default parameters
(level 20, window 64, linker 1).  Run from the repository root: python tests/golden/make_dust_golden.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.test_dust import dust_cases, ref_dust  # noqa: E402

out = {"params": [20, 64, 1], "intervals": [ref_dust(q) for q in dust_cases()]}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dust_golden.json"), "w"))
print(sum(len(x) for x in out["intervals"]), "intervals")
