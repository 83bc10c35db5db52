"""GPU parity: the CUDA path (through the C ABI) against the oracle port and, where the prebuilt
reference library travelled with the snapshot, against the reference engine itself.

Bit-exact bar: init-HSPs, per-chunk gapped lists, final HSP lists and E-value bit patterns.
"""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def _setup(name):
    from oracle import refdriver as R, portdriver as P
    task, cfgkw, vol, qs = cases.make_case(name)
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so did not travel to this box")
    cfg = R.default_config(task, taps=R.TAP_SCAN | R.TAP_INIT | R.TAP_GAPPED | R.TAP_LUT, **cfgkw)
    r = R.search(qs, vol, cfg)
    assert r["status"] == 0
    h = P.batch_from_reference(r, task=task, cfg=cfg)
    return r, h, vol


@pytest.mark.parametrize("name", cases.ALL)
def test_gpu_matches_reference_and_port(name):
    from gblastn_b200 import engine as E, abi
    from oracle import portdriver as P
    r, h, vol = _setup(name)
    p = P.search(h, vol, taps=P.TAP_INIT | P.TAP_GAPPED)
    V = E.Volume(vol)
    Q = E.Query(h)
    try:
        g = E.prelim_search(V, Q, taps=abi.BN_TAP_INIT | abi.BN_TAP_GAPPED)
        # stage taps
        assert np.array_equal(P.init_table(g["init"]), r["init"]), "init-HSPs differ from reference"
        assert np.array_equal(P.gapped_table(g["gapped"]), r["gapped"]), "gapped lists differ from reference"
        assert np.array_equal(P.final_table(g["hsps"]), r["final"]), "final lists / E-value bits differ"
        assert np.array_equal(P.final_table(g["hsps"]), P.final_table(p["hsps"])), "differs from oracle port"
        st = g["stats"]
        assert st["lookup_hits"] == r["lookup_hits"]
        assert st["good_init_extends"] == r["good_init_extends"]
        assert st["gap_extensions"] == r["gap_extensions"]
        assert st["good_extensions"] == r["good_extensions"]
        # scan tap, per subject
        for oid in range(len(vol.seq_len)):
            if vol.seq_len[oid] > 200_000_000:
                continue
            pairs = E.scan_subject(V, Q, oid)
            ref = r["scan"][r["scan"][:, 0] == oid][:, 2:4].astype(np.uint32)
            got = np.stack([pairs["q_off"], pairs["s_off"]], axis=1) if pairs.size else np.zeros((0, 2), np.uint32)
            assert np.array_equal(got, ref), f"scan pairs differ for oid {oid}"
    finally:
        Q.free()
        V.free()


def test_host_buffer_entry_point():
    from gblastn_b200 import engine as E
    from oracle import portdriver as P
    r, h, vol = _setup("mb_lut11_hash_indels")
    g = E.prelim_search_host(h, vol)
    assert np.array_equal(P.final_table(g["hsps"]), r["final"])
