"""The C port (oracle/port) against the reference engine itself (oracle/_ref), tap by tap.

This is what pins the oracle: scan pairs, init-HSPs, per-chunk gapped lists and the final
per-subject lists with E-value bit patterns must all be identical.
Skipped where the reference sources are not available (oracle/_ref is built from /root/reference).
"""
import numpy as np
import pytest

from tests import cases


def _both(name, built):
    from oracle import refdriver as R, portdriver as P
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so not built (needs /root/reference)")
    task, cfgkw, vol, qs = cases.make_case(name)
    cfg = R.default_config(task, taps=R.TAP_SCAN | R.TAP_INIT | R.TAP_GAPPED | R.TAP_LUT, **cfgkw)
    r = R.search(qs, vol, cfg)
    assert r["status"] == 0
    h = P.batch_from_reference(r, task=task, cfg=cfg)
    p = P.search(h, vol, taps=P.TAP_SCAN | P.TAP_INIT | P.TAP_GAPPED)
    assert p["status"] == 0
    return r, p


@pytest.mark.parametrize("name", cases.ALL)
def test_port_matches_reference(name, built):
    from oracle import portdriver as P
    r, p = _both(name, built)
    scan = np.stack([p["scan_oid"], p["scan_chunk"], p["scan"]["q_off"].astype(np.int32),
                     p["scan"]["s_off"].astype(np.int32)], axis=1) if p["scan"].size else np.zeros((0, 4), np.int32)
    assert np.array_equal(r["scan"], scan), "scan tap differs"
    assert np.array_equal(r["init"], P.init_table(p["init"])), "init-HSP tap differs"
    assert np.array_equal(r["gapped"], P.gapped_table(p["gapped"])), "gapped tap differs"
    assert np.array_equal(r["final"], P.final_table(p["hsps"])), "final lists / E-value bits differ"
    st = p["stats"]
    assert (st["lookup_hits"], st["init_extends"], st["good_init_extends"], st["gap_extensions"],
            st["good_extensions"]) == (r["lookup_hits"], r["init_extends"], r["good_init_extends"],
                                       r["gap_extensions"], r["good_extensions"])


def test_cases_exercise_every_table_shape(built):
    """The case list must keep covering MB lut 11/12, small table, both containers, both aligners."""
    from oracle import refdriver as R
    if not R.available():
        pytest.skip("reference not built")
    seen = set()
    for name in cases.ALL:
        task, cfgkw, vol, qs = cases.make_case(name)
        r = R.search(qs, vol, R.default_config(task, **cfgkw))
        seen.add((r["lut_type"], r["lut_word_length"], r["scan_step"] % 4 == 0, r["container_type"]))
    assert any(s[0] == 0 and s[1] == 11 for s in seen)
    assert any(s[0] == 0 and s[1] == 12 for s in seen)
    assert any(s[0] == 1 for s in seen)
    assert {s[3] for s in seen} == {0, 1}
    assert any(s[2] for s in seen) and any(not s[2] for s in seen)


def _db_masks(vol, rng, k):
    out = []
    for L in vol.seq_len:
        L = int(L)
        m, pos = [], int(rng.integers(0, max(1, L // 10)))
        if rng.random() < 0.3:
            pos = 0
        while pos < L and len(m) < k:
            end = min(L, pos + int(rng.integers(5, max(6, L // 8))))
            m.append((pos, end))
            pos = end + int(rng.integers(1, max(2, L // 6)))
        out.append(m)
    return out


@pytest.mark.parametrize("name", ["mb_lut11_hash_indels", "blastn_mb11_dp", "mb_smallna_diagarray", "blastn_ws7_na_table"])
@pytest.mark.parametrize("mask_type", [1, 2])
def test_port_matches_reference_with_database_masks(name, mask_type, built):
    """Soft / hard database masks (BLAST_SequenceBlk::seq_ranges, core/blast_engine.c:136-301,
    core/masksubj.inl): init-HSPs, gapped lists, final lists and the lookup-hit count of the port == reference."""
    from oracle import refdriver as R, portdriver as P
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so not built (needs /root/reference)")
    task, cfgkw, vol, qs = cases.make_case(name)
    sm = _db_masks(vol, np.random.default_rng(17 * mask_type + len(name)), 6)
    cfg = R.default_config(task, taps=R.TAP_INIT | R.TAP_GAPPED | R.TAP_LUT, **cfgkw)
    r = R.search(qs, vol, cfg, subject_masks=sm, subject_mask_type=mask_type)
    assert r["status"] == 0
    h = P.batch_from_reference(r, task=task, cfg=cfg)
    p = P.search(h, vol, taps=P.TAP_INIT | P.TAP_GAPPED, subject_masks=sm, subject_mask_type=mask_type)
    assert np.array_equal(r["init"], P.init_table(p["init"]))
    assert np.array_equal(r["gapped"], P.gapped_table(p["gapped"]))
    assert np.array_equal(r["final"], P.final_table(p["hsps"]))
    assert p["stats"]["lookup_hits"] == r["lookup_hits"]


@pytest.mark.parametrize("name", ["mb_bridged_segments", "blastn_bridged_segments"])
def test_traceback_golden_fixture_is_current(name):
    """tests/golden/traceback_<case>.npz equals what the reference built here produces today."""
    import os
    from tests import cases
    from oracle import refdriver as R
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so not built")
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"traceback_{name}.npz"))
    task, cfgkw, vol, qs = cases.make_case(name)
    r = R.search(qs, vol, R.default_config(task, taps=R.TAP_TRACEBACK, prelim_only=0, **cfgkw))
    assert r["status"] == 0
    assert np.array_equal(r["tb_final"], gold["tb_final"]) and np.array_equal(r["tb_ops"], gold["tb_ops"])
    assert np.array_equal(r["final"], gold["prelim_final"])
    # the adversarial cases must keep exercising the list logic: fewer alignments than preliminary HSPs, fewer results still
    assert r["tb_calls"].shape[0] < r["final"].shape[0]
    assert r["tb_final"].shape[0] <= r["tb_calls"].shape[0]
