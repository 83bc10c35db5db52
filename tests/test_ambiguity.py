"""Subject ambiguity data in the traceback stage.

The preliminary stage reads a database sequence as its 2-bit bases (ambiguous positions hold random bases); the traceback
stage fetches it in blastna with the volume's ambiguity runs restored (CSeqDBVol::x_GetAmbigSeq,
objtools/blast/seqdb_reader/seqdbvol.cpp:832-870, 1565-1640) and aligns, re-evaluates and counts identities on that.
bn_db_load_files parses the runs of the .nsq; the traceback kernels lay them over the packed bases.  Checked on the
reference's own fixture seqdb_reader/data/seqn (100 sequences, 63 of them with ambiguity runs; tests/golden/seqn.*) and on
synthetic volumes with overlapping / long runs, against the reference engine fed the same runs.
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_ambiguity_runs_of_the_fixture(built):
    """CPU: the .nsq ambiguity sections parse into runs inside their sequences; ntshort has none."""
    from gblastn_b200 import engine as E
    nin, nsq = os.path.join(GOLD, "seqn.nin"), os.path.join(GOLD, "seqn.nsq")
    info, off, ln = E.dbfile_index(nin, nsq)
    first, runs = E.dbfile_ambiguity(nin, nsq)
    assert info["n_seq"] == 100 and first.shape[0] == 101 and runs.shape[0] == first[-1] == 779
    with_amb = int((np.diff(first) > 0).sum())
    assert with_amb == 63
    for i in range(100):
        r = runs[first[i]:first[i + 1]]
        assert (r[:, 0] >= 0).all() and (r[:, 0] + r[:, 1] <= ln[i]).all() and (r[:, 1] >= 1).all()
        assert ((r[:, 2] >= 4) & (r[:, 2] <= 14)).all()
    f2, r2 = E.dbfile_ambiguity(os.path.join(GOLD, "ntshort.nin"), os.path.join(GOLD, "ntshort.nsq"))
    assert r2.shape[0] == 0 and (f2 == 0).all()


def _restored(vol, oid, first, runs):
    b = vol.bases(oid).copy()
    for a, n, code in runs[first[oid]:first[oid + 1]]:
        b[a:a + n] = code
    return b


def _queries_over_runs(vol, first, runs, rng, n_q, flank):
    """Queries cut around ambiguity runs: mutated copies of the 2-bit bases, every third one with the ambiguity codes
    themselves in the query (N against N)."""
    qs = []
    oids = [i for i in range(len(vol.seq_len)) if first[i + 1] > first[i] and vol.seq_len[i] > 2 * flank + 20]
    for k in range(n_q):
        oid = oids[int(rng.integers(0, len(oids)))]
        r = runs[first[oid] + int(rng.integers(0, first[oid + 1] - first[oid]))]
        L = int(vol.seq_len[oid])
        a = max(0, int(r[0]) - flank + int(rng.integers(-20, 20)))
        e = min(L, a + 2 * flank)
        src = _restored(vol, oid, first, runs) if k % 3 == 0 else vol.bases(oid)
        q = src[a:e].copy()
        mut = (rng.random(q.size) < 0.03) & (q < 4)
        q[mut] = (q[mut] + 1 + rng.integers(0, 3, int(mut.sum()))) % 4
        if k % 2:
            comp = np.array([3, 2, 1, 0, 5, 4, 7, 6, 8, 9, 13, 12, 11, 10, 14, 15], np.uint8)
            q = comp[q[::-1]]
        qs.append(np.ascontiguousarray(q, dtype=np.uint8))
    return qs


def _compare_final(got, ops, r):
    want, ref_ops = r["tb_final"], r["tb_ops"]
    assert got.shape[0] == want.shape[0] and want.shape[0] > 0
    for k, col in enumerate(("query_index", "oid", "context", "q_off", "q_end", "s_off", "s_end", "score", "num_ident")):
        bad = np.flatnonzero(got[col] != want[:, k])
        assert bad.size == 0, f"{col} differs at {bad[:5]}: {got[col][bad[:5]]} vs {want[bad[:5], k]}"
    ev = want[:, 9].astype(np.uint32).astype(np.uint64) | (want[:, 10].astype(np.uint32).astype(np.uint64) << np.uint64(32))
    bs = want[:, 11].astype(np.uint32).astype(np.uint64) | (want[:, 12].astype(np.uint32).astype(np.uint64) << np.uint64(32))
    assert np.array_equal(got["evalue"].view(np.uint64), ev) and np.array_equal(got["bit_score"].view(np.uint64), bs)
    flat = np.concatenate([ref_ops[want[i, 13]:want[i, 13] + want[i, 14]] for i in range(want.shape[0])])
    mine = np.concatenate([np.stack([ops["op_type"][a:a + n], ops["num"][a:a + n]], axis=1)
                           for a, n in zip(got["esp_off"], got["esp_n"])])
    assert np.array_equal(flat, mine), "edit scripts differ"


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["megablast", "blastn"])
def test_traceback_with_ambiguity_on_blast_db(task):
    """The reference's seqn volume loaded from its files: preliminary + traceback stage on the product path == the
    reference engine with the same ambiguity runs, and NOT equal to a run that ignores them."""
    from gblastn_b200 import engine as E, setup as S, synth, abi
    from oracle import refdriver as R
    if not R.available():
        pytest.skip("reference library not present")
    nin, nsq = os.path.join(GOLD, "seqn.nin"), os.path.join(GOLD, "seqn.nsq")
    info, off, ln = E.dbfile_index(nin, nsq)
    first, runs = E.dbfile_ambiguity(nin, nsq)
    raw = np.fromfile(nsq, dtype=np.uint8)
    vol = synth.Volume(packed=np.concatenate([raw, np.zeros(32, np.uint8)]), byte_off=off, seq_len=ln)
    qs = _queries_over_runs(vol, first, runs, np.random.default_rng(11), 40, 110)
    cfg = R.default_config(task, taps=R.TAP_TRACEBACK, prelim_only=0)
    r = R.search(qs, vol, cfg, ambiguity=(first, runs))
    r_plain = R.search(qs, vol, R.default_config(task, taps=R.TAP_TRACEBACK, prelim_only=0))
    assert r["status"] == 0 and r["tb_final"].shape[0] > 10
    assert not np.array_equal(r["tb_final"][:, :9], r_plain["tb_final"][:, :9]), "the ambiguity runs change nothing here"
    s = S.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs)
    V, Q = E.FileVolume(nin, nsq), E.Query(s.batch)
    try:
        g = E.prelim_search(V, Q)
        got, ops = E.traceback_search(V, Q, s.gap_x_dropoff_final(), g["hsps"])
        _compare_final(got, ops, r)
        # every alignment call the reference made, on its own inputs
        calls = r["tb_calls"]
        items = np.zeros(calls.shape[0], dtype=abi.TB_ITEM_DTYPE)
        items["oid"], items["context"], items["s_shift"] = calls[:, 1], calls[:, 2], calls[:, 3]
        items["q_start"], items["s_start"], items["s_length"] = calls[:, 4], calls[:, 5], calls[:, 7]
        res, cops = E.gapped_traceback(V, Q, int(r["gap_x_dropoff_final"]), items)
        for k, col in ((8, "score"), (9, "query_start"), (10, "query_stop"), (11, "subject_start"), (12, "subject_stop"),
                       (14, "esp_n")):
            assert np.array_equal(res[col], calls[:, k]), col
        for i in range(calls.shape[0]):
            w = r["tb_ops"][calls[i, 13]:calls[i, 13] + calls[i, 14]]
            c = cops[res["esp_off"][i]:res["esp_off"][i] + res["esp_n"][i]]
            assert np.array_equal(c["op_type"], w[:, 0]) and np.array_equal(c["num"], w[:, 1])
    finally:
        Q.free(); V.free(); s.free()


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["megablast", "blastn"])
def test_traceback_with_synthetic_ambiguity(task):
    """In-memory volume + bn_db_set_ambiguity: runs of every ambiguity code, long runs, runs that overlap (later ones
    win, as in the reference), runs at both ends of a sequence; removing the runs restores the plain answer."""
    from gblastn_b200 import engine as E, setup as S, synth
    from oracle import refdriver as R
    if not R.available():
        pytest.skip("reference library not present")
    rng = np.random.default_rng(21)
    vol = synth.random_volume([60_000, 25_000, 9_000, 300], seed=22)
    first, rows = [0], []
    for oid, L in enumerate(vol.seq_len):
        L = int(L)
        k = 0 if oid == 3 else 60
        starts = np.sort(rng.integers(0, L - 40, size=k))
        for j, a in enumerate(starts):
            n = int(rng.choice([1, 1, 2, 5, 30])) if j % 10 else 300
            rows.append((int(a), min(n, L - int(a)), int(rng.integers(4, 15))))
            if j % 7 == 0:                                  # an overlapping run right behind it
                rows.append((int(a) + n // 2, min(3, L - int(a) - n // 2), int(rng.integers(4, 15))))
        if oid == 0:
            rows.append((0, 4, 14)); rows.append((L - 5, 5, 14))
        first.append(len(rows))
    first, runs = np.array(first, np.int64), np.array(rows, np.int32).reshape(-1, 3)
    qs = _queries_over_runs(vol, first, runs, rng, 50, 160)
    r = R.search(qs, vol, R.default_config(task, taps=R.TAP_TRACEBACK, prelim_only=0), ambiguity=(first, runs))
    r_plain = R.search(qs, vol, R.default_config(task, taps=R.TAP_TRACEBACK, prelim_only=0))
    assert r["status"] == 0 and r["tb_final"].shape[0] > 10
    s = S.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs)
    V, Q = E.Volume(vol), E.Query(s.batch)
    try:
        V.set_ambiguity(first, runs)
        g = E.prelim_search(V, Q)
        got, ops = E.traceback_search(V, Q, s.gap_x_dropoff_final(), g["hsps"])
        _compare_final(got, ops, r)
        V.set_ambiguity(None, None)
        got, ops = E.traceback_search(V, Q, s.gap_x_dropoff_final(), g["hsps"])
        _compare_final(got, ops, r_plain)
    finally:
        Q.free(); V.free(); s.free()
