"""Host set-up mirror (bn_setup_*) against the reference's own set-up (no GPU needed).

Lookup-table arrays, query layout and every integer parameter must be identical; Karlin-Altschul
doubles that come from published tables must be bit-identical, the composition-dependent ungapped
block must agree to 1e-12 relative (the reference is built with -ffast-math).
"""
import numpy as np
import pytest

from tests import cases


def _pair(name, masks=None, **extra):
    from oracle import refdriver as R
    from gblastn_b200 import setup as S
    if not R.available():
        pytest.skip("reference not built")
    task, cfgkw, vol, qs = cases.make_case(name)
    cfgkw.update(extra)
    cfg = R.default_config(task, taps=R.TAP_LUT, **cfgkw)
    r = R.search(qs, vol, cfg, masks=masks)
    assert r["status"] == 0
    s = S.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs, masks=masks, **cfgkw)
    return r, s


def _compare(r, s):
    b = s.batch
    assert b.lut_type == r["lut_type"]
    assert (b.word_length, b.lut_word_length, b.scan_step) == (r["word_length"], r["lut_word_length"], r["scan_step"])
    assert b.hashsize == r["hashsize"]
    assert np.array_equal(s.concat_query, r["concat_query"])
    if b.lut_type == 0:
        assert np.array_equal(s.hashtable, r["hashtable"])
        assert np.array_equal(s.next_pos, r["next_pos"])
        assert b.pv_array_bts == r["pv_array_bts"]
        assert np.array_equal(s.pv_array, r["pv_array"])
    elif b.lut_type == 2:
        # eNaLookupTable: thick backbone cells {num_used, entries[3] | overflow_cursor} and overflow
        assert np.array_equal(s.na_backbone, r["na_backbone"])
        if r["na_overflow"] is not None and r["na_overflow"].size:
            assert np.array_equal(s.na_overflow[: r["na_overflow"].size], r["na_overflow"])
    else:
        assert np.array_equal(s.backbone, r["backbone"])
        # overflow[0..1] are never written by the reference (malloc garbage): compare from 2
        assert np.array_equal(s.overflow[2:], r["overflow"][2:])
    assert s.longest_chain() == r["longest_chain"]
    assert b.container_type == r["container_type"]
    assert b.gap_x_dropoff == r["gap_x_dropoff"]
    assert s.gap_x_dropoff_final() == r["gap_x_dropoff_final"]
    assert np.array_equal(np.array(b.nucl_score_table), r["nucl_score_table"])
    assert np.array_equal(np.array(b.matrix).reshape(16, 16), r["matrix"])
    if r["n_masked_locations"] < 0:
        assert not b.masked_locations
    else:
        assert b.masked_locations
        assert np.array_equal(s.masked_locations, r["masked_locations"])
    ctx = s.contexts()
    assert len(ctx) == r["num_contexts"]
    for i, c in enumerate(ctx):
        assert c.query_offset == r["ctx_query_offset"][i]
        assert c.query_length == r["ctx_query_length"][i]
        assert c.length_adjustment == r["ctx_length_adjustment"][i]
        assert c.eff_searchsp == r["ctx_eff_searchsp"][i]
        assert c.x_dropoff == r["ctx_x_dropoff"][i]
        assert c.cutoff_score == r["ctx_cutoff_score"][i]
        assert c.reduced_cutoff == r["ctx_reduced_cutoff"][i]
        assert c.gapped_cutoff == r["ctx_gapped_cutoff"][i]
    # The ungapped Karlin-Altschul block comes out of a Newton-Raphson iteration that the reference
    # compiles with -ffast-math (core/Makefile.blast.lib:19): equal to 1e-12 relative, not bit for bit.
    assert np.allclose(s.kbp_std(), r["ctx_kbp_std"], rtol=1e-12, atol=0)
    if np.array_equal(r["ctx_kbp_gap"], r["ctx_kbp_std"]):
        # gap costs beyond the tabulated ones: Blast_KarlinBlkNuclGappedCalc copies the ungapped block
        # (core/blast_stat.c:3868-3871), so the same tolerance applies
        assert np.allclose(s.kbp_gap(), r["ctx_kbp_gap"], rtol=1e-12, atol=0)
    else:
        for i, c in enumerate(ctx):
            assert c.gap_lambda == r["ctx_kbp_gap"][i, 0]
            assert c.gap_logK == r["ctx_kbp_gap"][i, 2]
        assert np.array_equal(s.kbp_gap(), r["ctx_kbp_gap"])


@pytest.mark.parametrize("name", cases.ALL)
def test_setup_matches_reference(name, built):
    r, s = _pair(name)
    try:
        _compare(r, s)
    finally:
        s.free()


@pytest.mark.parametrize("mask_at_hash", [1, 0])
def test_setup_with_query_masks(mask_at_hash, built):
    name = "mb_lut11_hash_indels"
    _, _, _, qs = cases.make_case(name)
    rng = np.random.default_rng(5)
    masks = []
    for q in qs:
        m = []
        if rng.random() < 0.6:
            a = int(rng.integers(0, len(q) - 80))
            m.append((a, a + int(rng.integers(10, 70))))
            if rng.random() < 0.5:
                b0 = int(rng.integers(0, len(q) - 40))
                m.append((b0, b0 + int(rng.integers(5, 35))))
        if rng.random() < 0.1:
            m.append((0, 25))
        if rng.random() < 0.1:
            m.append((len(q) - 30, len(q) - 1))
        masks.append(m)
    r, s = _pair(name, masks=masks, mask_at_hash=mask_at_hash)
    try:
        _compare(r, s)
    finally:
        s.free()


@pytest.mark.parametrize("scores", [(1, -3, 2, 2), (1, -1, 3, 2), (2, -3, 4, 4), (1, -2, 2, 2), (4, -5, 6, 5)])
def test_setup_other_scoring_systems(scores, built):
    rw, pn, go, ge = scores
    r, s = _pair("blastn_mb11_dp", reward=rw, penalty=pn, gap_open=go, gap_extend=ge)
    try:
        _compare(r, s)
    finally:
        s.free()
