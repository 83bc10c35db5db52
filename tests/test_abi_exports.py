"""The C-ABI library loads on a CPU-only box, exports every symbol include/gblastn_b200.h declares,
and fails loudly (BN_ERR_NO_DEVICE) instead of falling back when no GPU is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gblastn_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bn_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_are_exported(built):
    from gblastn_b200 import engine
    lib = engine.lib()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gblastn_b200.h but not exported"
    assert set(syms) == set(engine.EXPORTS)


def test_no_silent_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gblastn_b200 import engine, abi
    lib = engine.lib()
    rc = lib.bn_init(C.c_int(0), None)
    assert rc == abi.BN_ERR_NO_DEVICE
    assert b"no usable CUDA device" in lib.bn_last_error()


def test_product_does_not_import_oracle():
    """No Python import, C include or dlopen of anything under oracle/ inside the product package
    (build.py may *compile* the oracles: building the checker is not using it)."""
    pkg = os.path.join(ROOT, "gblastn_b200")
    bad = re.compile(r"^\s*(from\s+oracle|import\s+oracle)|#include\s*[\"<][^\">]*oracle|CDLL\([^)]*oracle|liboracle|libblastref", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) and f != "build.py":
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), f"{f} uses the oracle: the product path must not depend on it"


def test_replay_per_strand_equals_single_tree(built):
    """The containment replay keeps one interval tree per query strand; the reference keeps one per subject
    chunk.  Seeded random init-HSP sets (nested boxes, shared end points, equal scores, 1-12 strands):
    identical outputs in every case.  Host-only."""
    from gblastn_b200 import engine
    lib = engine.lib()
    for seed in (1, 2, 3):
        n = C.c_int64(-1)
        assert lib.bn_selftest_replay(C.c_uint64(seed), C.c_int32(20000), C.byref(n)) == 0
        assert n.value == 0
