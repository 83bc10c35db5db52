"""Known-answer test taken from the reference's own unit tests:
MegablastGreedyTraceback2, c++/src/algo/blast/unit_tests/api/bl2seq_unit_test.cpp:1620-1690 —
greedy1a.fsa vs greedy1b.fsa with megablast defaults scores 619, and 6034 with reward 10 / penalty -25 /
gapped X-dropoff 100 (the second value needs the odd-score rounding of sbp->round_down).
Checked for the reference engine built here, the C port, the product set-up and (on a GPU) the CUDA path."""
import os

import numpy as np
import pytest

from gblastn_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KAT = [({}, 619), (dict(reward=10, penalty=-25, xdrop_gap=100.0, xdrop_gap_final=100.0), 6034)]


def _fasta(path):
    s = "".join(line.strip() for line in open(path) if not line.startswith(">"))
    m = {"A": 0, "C": 1, "G": 2, "T": 3}
    return np.array([m.get(c.upper(), 14) for c in s], dtype=np.uint8)


def _inputs():
    a, b = _fasta(os.path.join(GOLD, "greedy1a.fsa")), _fasta(os.path.join(GOLD, "greedy1b.fsa"))
    return [a], synth.make_volume_from_bases([b])


@pytest.mark.parametrize("kw,score", KAT)
def test_kat_port_and_reference(kw, score):
    from oracle import refdriver as R, portdriver as P
    qs, vol = _inputs()
    if R.available():
        cfg = R.default_config("megablast", taps=R.TAP_LUT, **kw)
        r = R.search(qs, vol, cfg)
        assert r["status"] == 0 and r["final"].shape[0] == 1 and int(r["final"][0, 6]) == score
        h = P.batch_from_reference(r, task="megablast", cfg=cfg)
        p = P.search(h, vol)
        assert np.array_equal(P.final_table(p["hsps"]), r["final"])
    from gblastn_b200 import setup as S, abi
    s = S.Setup(qs, task="megablast", db_length=vol.total_bases, db_num_seqs=vol.n_seqs, **kw)
    try:
        h2 = abi.BatchHolder()
        h2.batch = s.batch
        p2 = P.search(h2, vol)
        assert p2["hsps"].size == 1 and int(p2["hsps"]["score"][0]) == score, "port + product set-up"
    finally:
        s.free()


@pytest.mark.gpu
@pytest.mark.parametrize("kw,score", KAT)
def test_kat_gpu(kw, score):
    from gblastn_b200 import engine as E, setup as S
    qs, vol = _inputs()
    s = S.Setup(qs, task="megablast", db_length=vol.total_bases, db_num_seqs=vol.n_seqs, **kw)
    V, Q = E.Volume(vol), E.Query(s.batch)
    try:
        g = E.prelim_search(V, Q)
        assert g["hsps"].size == 1 and int(g["hsps"]["score"][0]) == score
        assert (int(g["hsps"]["q_off"][0]), int(g["hsps"]["q_end"][0]), int(g["hsps"]["s_off"][0]),
                int(g["hsps"]["s_end"][0])) == (159, 874, 30, 739)
    finally:
        Q.free(); V.free(); s.free()


@pytest.mark.parametrize("kw,score", KAT)
def test_kat_reference_traceback_stage(kw, score):
    """The KAT is the score AFTER traceback (the unit test reads it from the Seq-align): the reference's
    traceback stage, run by oracle/ref_driver.c over the in-memory seqsrc, must report it for its single HSP."""
    from oracle import refdriver as R
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so not built")
    qs, vol = _inputs()
    r = R.search(qs, vol, R.default_config("megablast", taps=R.TAP_TRACEBACK, prelim_only=0, **kw))
    assert r["status"] == 0 and r["tb_final"].shape[0] == 1 and int(r["tb_final"][0, 7]) == score
    c = r["tb_calls"]
    # one greedy traceback call; its raw score is rounded down to even afterwards when reward is even
    # (Blast_HSPListAdjustOddBlastnScores in s_HSPListPostTracebackUpdate): 6035 -> 6034
    assert c.shape[0] == 1 and int(c[0, 0]) == 1 and int(c[0, 8]) in (score, score + 1)
    ops = r["tb_ops"][c[0, 13]:c[0, 13] + c[0, 14]]
    # the edit script spans the alignment: substitutions + insertions = query extent, + deletions = subject extent
    assert ops[ops[:, 0] != 0][:, 1].sum() == c[0, 10] - c[0, 9]
    assert ops[ops[:, 0] != 6][:, 1].sum() == c[0, 12] - c[0, 11]


@pytest.mark.gpu
@pytest.mark.parametrize("kw,score", KAT)
def test_kat_gpu_traceback(kw, score):
    """bn_gapped_traceback on the KAT's traceback call: 619 / 6034 and the reference's edit script."""
    from gblastn_b200 import engine as E, abi
    from oracle import refdriver as R, portdriver as P
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so did not travel to this box")
    qs, vol = _inputs()
    cfg = R.default_config("megablast", taps=R.TAP_LUT | R.TAP_TRACEBACK, prelim_only=0, **kw)
    r = R.search(qs, vol, cfg)
    c = r["tb_calls"]
    h = P.batch_from_reference(r, task="megablast", cfg=cfg)
    V, Q = E.Volume(vol), E.Query(h)
    try:
        items = np.zeros(1, dtype=abi.TB_ITEM_DTYPE)
        items["oid"], items["context"], items["s_shift"] = c[:, 1], c[:, 2], c[:, 3]
        items["q_start"], items["s_start"], items["s_length"] = c[:, 4], c[:, 5], c[:, 7]
        res, ops = E.gapped_traceback(V, Q, int(r["gap_x_dropoff_final"]), items)
        assert int(res["score"][0]) == int(c[0, 8]) and int(c[0, 8]) in (score, score + 1)
        assert int(r["tb_final"][0, 7]) == score
        want = r["tb_ops"][c[0, 13]:c[0, 13] + c[0, 14]]
        assert np.array_equal(ops["op_type"], want[:, 0]) and np.array_equal(ops["num"], want[:, 1])
    finally:
        Q.free(); V.free()


# ---- NucleotideBlastWordSize4 / _EOS (bl2seq_unit_test.cpp:2238-2301): invariants of the reference's test -------
SIZE4 = [("blastn_size4a.fsa", "blastn_size4b.fsa"), ("blastn_size4c.fsa", "blastn_size4d.fsa")]


def _well_formed(h):
    assert h.size > 0
    assert (h["q_off"] < h["q_end"]).all() and (h["s_off"] < h["s_end"]).all()
    assert ((h["q_gapped_start"] >= h["q_off"]) & (h["q_gapped_start"] < h["q_end"])).all()
    assert ((h["s_gapped_start"] >= h["s_off"]) & (h["s_gapped_start"] < h["s_end"])).all()
    assert (h["q_end"] - h["q_off"] >= 4).all() and (h["s_end"] - h["s_off"] >= 4).all()


@pytest.mark.parametrize("qf,sf", SIZE4)
def test_wordsize4_port(qf, sf):
    from oracle import refdriver as R, portdriver as P
    from gblastn_b200 import setup as S, abi
    qs, vol = [_fasta(os.path.join(GOLD, qf))], synth.make_volume_from_bases([_fasta(os.path.join(GOLD, sf))])
    s = S.Setup(qs, task="blastn", db_length=vol.total_bases, db_num_seqs=vol.n_seqs, word_size=4)
    try:
        h = abi.BatchHolder()
        h.batch = s.batch
        p = P.search(h, vol)
        _well_formed(p["hsps"])
        if R.available():
            r = R.search(qs, vol, R.default_config("blastn", word_size=4))
            assert np.array_equal(P.final_table(p["hsps"]), r["final"])
    finally:
        s.free()


@pytest.mark.gpu
@pytest.mark.parametrize("qf,sf", SIZE4)
def test_wordsize4_gpu(qf, sf):
    from oracle import refdriver as R, portdriver as P
    from gblastn_b200 import engine as E, setup as S
    qs, vol = [_fasta(os.path.join(GOLD, qf))], synth.make_volume_from_bases([_fasta(os.path.join(GOLD, sf))])
    s = S.Setup(qs, task="blastn", db_length=vol.total_bases, db_num_seqs=vol.n_seqs, word_size=4)
    V, Q = E.Volume(vol), E.Query(s.batch)
    try:
        g = E.prelim_search(V, Q)
        _well_formed(g["hsps"])
        if R.available():
            r = R.search(qs, vol, R.default_config("blastn", word_size=4))
            assert np.array_equal(P.final_table(g["hsps"]), r["final"])
    finally:
        Q.free(); V.free(); s.free()
