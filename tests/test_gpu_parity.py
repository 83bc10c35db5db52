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
        oids = list(range(len(vol.seq_len)))
        if len(oids) > 12:      # many-subject volumes: the 6 longest + 6 seeded picks
            longest = np.argsort(vol.seq_len)[-6:].tolist()
            oids = sorted(set(longest + np.random.default_rng(1).choice(len(vol.seq_len), 6, replace=False).tolist()))
        for oid in oids:
            if vol.seq_len[oid] > 200_000_000:
                continue
            pairs = E.scan_subject(V, Q, oid)
            ref = r["scan"][r["scan"][:, 0] == oid][:, 2:4].astype(np.uint32)
            got = np.stack([pairs["q_off"], pairs["s_off"]], axis=1) if pairs.size else np.zeros((0, 2), np.uint32)
            assert np.array_equal(got, ref), f"scan pairs differ for oid {oid}"
    finally:
        Q.free()
        V.free()


@pytest.mark.parametrize("name", ["mb_lut11_hash_indels", "blastn_mb11_dp", "mb_long_divergent_tier2",
                                  "mb_smallna_diagarray"])
def test_get_gapped_score_drop_in(name):
    """bn_get_gapped_score == BLAST_GetGappedScore: fed the REFERENCE's init-hit list of every subject
    chunk, it returns the reference's gapped list of that chunk."""
    from gblastn_b200 import engine as E, abi
    r, h, vol = _setup(name)
    V, Q = E.Volume(vol), E.Query(h)
    try:
        init, gapped = r["init"], r["gapped"]
        chunks = sorted(set(map(tuple, init[:, 0:2].tolist())))
        assert chunks
        for oid, chunk_off in chunks:
            rows = init[(init[:, 0] == oid) & (init[:, 1] == chunk_off)]
            arr = np.zeros(rows.shape[0], dtype=abi.INIT_DTYPE)
            for k, col in enumerate(("oid", "chunk_off", "q_off", "s_off", "q_start", "s_start", "length", "score")):
                arr[col] = rows[:, k]
            got = E.get_gapped_score(V, Q, oid, chunk_off, arr)
            want = gapped[(gapped[:, 0] == oid) & (gapped[:, 1] == chunk_off)]
            from oracle import portdriver as P
            assert np.array_equal(P.gapped_table(got), want), f"gapped list differs for oid {oid} chunk {chunk_off}"
    finally:
        Q.free(); V.free()


TRACEBACK_DP_CASES = ["blastn_mb11_dp", "blastn_smallna_dp", "blastn_ws7_array", "blastn_ntlike_many_subjects",
                      "c3_scaled_blastn_10kb", "blastn_bridged_segments", "blastn_ws7_na_table"]
TRACEBACK_GREEDY_CASES = ["c1_megablast_10kb_vs_1mb", "mb_lut11_hash_indels", "mb_smallna_diagarray", "mb_ws16", "mb_with_N",
                          "blastn_ws11_greedy", "mb_ntlike_many_subjects", "mb_long_divergent_tier2", "mb_lut12_stride17",
                          "c4_scaled_short_reads", "c5_scaled_ntlike_5kb", "mb_bridged_segments", "mb_two_hit_w40_hash",
                          "mb_affine_greedy_5_2", "blastn_affine_greedy_5_2", "mb_affine_greedy_long_tier2", "mb_affine_greedy_0_2"]


@pytest.mark.parametrize("name", TRACEBACK_DP_CASES + TRACEBACK_GREEDY_CASES)
def test_gapped_traceback_drop_in(name):
    """bn_gapped_traceback == BLAST_GappedAlignmentWithTraceback (dynamic programming) /
    BLAST_GreedyGappedAlignment with traceback (megablast): fed every call the REFERENCE's own traceback
    stage makes (Blast_TracebackFromHSPList, tapped in oracle/ref_driver.c), the device returns the same score,
    alignment bounds and edit script, operation for operation."""
    from gblastn_b200 import engine as E, abi
    from oracle import refdriver as R, portdriver as P
    task, cfgkw, vol, qs = cases.make_case(name)
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so did not travel to this box")
    cfg = R.default_config(task, taps=R.TAP_LUT | R.TAP_TRACEBACK, prelim_only=0, **cfgkw)
    r = R.search(qs, vol, cfg)
    assert r["status"] == 0
    calls = r["tb_calls"]
    assert calls.shape[0] > 0
    assert (calls[:, 0] == (1 if name in TRACEBACK_GREEDY_CASES else 0)).all()
    h = P.batch_from_reference(r, task=task, cfg=cfg)
    V, Q = E.Volume(vol), E.Query(h)
    try:
        items = np.zeros(calls.shape[0], dtype=abi.TB_ITEM_DTYPE)
        items["oid"], items["context"], items["s_shift"] = calls[:, 1], calls[:, 2], calls[:, 3]
        items["q_start"], items["s_start"], items["s_length"] = calls[:, 4], calls[:, 5], calls[:, 7]
        assert np.array_equal(calls[:, 6], r["ctx_query_length"][calls[:, 2]])
        res, ops = E.gapped_traceback(V, Q, int(r["gap_x_dropoff_final"]), items)
        assert (res["status"] == 0).all()
        for k, col in ((8, "score"), (9, "query_start"), (10, "query_stop"), (11, "subject_start"), (12, "subject_stop"),
                       (14, "esp_n")):
            bad = np.flatnonzero(res[col] != calls[:, k])
            assert bad.size == 0, f"{col} differs for {bad.size} of {calls.shape[0]} calls, first {bad[:3]}: " \
                                  f"{res[col][bad[:3]]} vs {calls[bad[:3], k]}"
        ref_ops = r["tb_ops"]
        for i in range(calls.shape[0]):
            want = ref_ops[calls[i, 13]:calls[i, 13] + calls[i, 14]]
            got = ops[res["esp_off"][i]:res["esp_off"][i] + res["esp_n"][i]]
            assert np.array_equal(got["op_type"], want[:, 0]) and np.array_equal(got["num"], want[:, 1]), \
                f"edit script differs for call {i}"
    finally:
        Q.free(); V.free()


@pytest.mark.parametrize("name", TRACEBACK_DP_CASES + TRACEBACK_GREEDY_CASES)
def test_traceback_hsps_from_prelim_lists(name):
    """bn_traceback_hsps: fed the reference's preliminary HSP lists (what its traceback stage reads from the HSP
    stream), the device derives for every HSP the start point and subject window of the call
    Blast_TracebackFromHSPList would make and aligns it.  Every call the reference really made (it skips HSPs
    contained in better ones) must be among them with identical inputs, score, bounds and edit script."""
    from gblastn_b200 import engine as E, abi
    from oracle import refdriver as R, portdriver as P
    task, cfgkw, vol, qs = cases.make_case(name)
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so did not travel to this box")
    cfg = R.default_config(task, taps=R.TAP_LUT | R.TAP_TRACEBACK, prelim_only=0, **cfgkw)
    r = R.search(qs, vol, cfg)
    assert r["status"] == 0
    calls, fin = r["tb_calls"], r["final"]
    assert calls.shape[0] > 0 and fin.shape[0] >= calls.shape[0]
    h = P.batch_from_reference(r, task=task, cfg=cfg)
    V, Q = E.Volume(vol), E.Query(h)
    try:
        hsps = np.zeros(fin.shape[0], dtype=abi.HSP_DTYPE)
        for k, col in enumerate(("oid", "context", "q_off", "q_end", "s_off", "s_end", "score", "q_gapped_start",
                                 "s_gapped_start")):
            hsps[col] = fin[:, k]
        items, res, ops = E.traceback_hsps(V, Q, int(r["gap_x_dropoff_final"]), hsps)
        index = {}
        for i in range(items.shape[0]):
            if items["oid"][i] >= 0:
                key = tuple(int(items[c][i]) for c in ("oid", "context", "s_shift", "q_start", "s_start", "s_length"))
                index.setdefault(key, i)
        ref_ops = r["tb_ops"]
        for j in range(calls.shape[0]):
            c = calls[j]
            key = (int(c[1]), int(c[2]), int(c[3]), int(c[4]), int(c[5]), int(c[7]))
            assert key in index, f"reference call {j} {key} has no counterpart among the device's start points"
            i = index[key]
            got = tuple(int(res[f][i]) for f in ("score", "query_start", "query_stop", "subject_start", "subject_stop"))
            assert got == tuple(int(x) for x in c[8:13]), f"call {j}: {got} vs {c[8:13]}"
            want = ref_ops[c[13]:c[13] + c[14]]
            g = ops[res["esp_off"][i]:res["esp_off"][i] + res["esp_n"][i]]
            assert np.array_equal(g["op_type"], want[:, 0]) and np.array_equal(g["num"], want[:, 1]), f"edit script {j}"
    finally:
        Q.free(); V.free()


@pytest.mark.parametrize("name", TRACEBACK_DP_CASES + TRACEBACK_GREEDY_CASES)
def test_traceback_search_matches_reference(name):
    """bn_traceback_search == the reference's whole traceback stage (Blast_RunTracebackSearch): fed the reference's
    preliminary lists it returns the reference's final results — every HSP of every hit list in BlastHSPResults order:
    coordinates, score, number of identities, E-value and bit-score bit patterns, edit scripts."""
    from gblastn_b200 import engine as E, abi
    from oracle import refdriver as R, portdriver as P
    task, cfgkw, vol, qs = cases.make_case(name)
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so did not travel to this box")
    cfg = R.default_config(task, taps=R.TAP_LUT | R.TAP_TRACEBACK, prelim_only=0, **cfgkw)
    r = R.search(qs, vol, cfg)
    assert r["status"] == 0
    fin, want = r["final"], r["tb_final"]
    assert want.shape[0] > 0
    h = P.batch_from_reference(r, task=task, cfg=cfg)
    V, Q = E.Volume(vol), E.Query(h)
    try:
        hsps = np.zeros(fin.shape[0], dtype=abi.HSP_DTYPE)
        for k, col in enumerate(("oid", "context", "q_off", "q_end", "s_off", "s_end", "score", "q_gapped_start",
                                 "s_gapped_start")):
            hsps[col] = fin[:, k]
        got, ops = E.traceback_search(V, Q, int(r["gap_x_dropoff_final"]), hsps)
        assert got.shape[0] == want.shape[0], f"{got.shape[0]} HSPs, reference has {want.shape[0]}"
        for k, col in enumerate(("query_index", "oid", "context", "q_off", "q_end", "s_off", "s_end", "score", "num_ident")):
            bad = np.flatnonzero(got[col] != want[:, k])
            assert bad.size == 0, f"{col} differs at {bad[:5]}: {got[col][bad[:5]]} vs {want[bad[:5], k]}"
        ev = want[:, 9].astype(np.uint32).astype(np.uint64) | (want[:, 10].astype(np.uint32).astype(np.uint64) << np.uint64(32))
        bs = want[:, 11].astype(np.uint32).astype(np.uint64) | (want[:, 12].astype(np.uint32).astype(np.uint64) << np.uint64(32))
        assert np.array_equal(got["evalue"].view(np.uint64), ev), "E-value bit patterns differ"
        assert np.array_equal(got["bit_score"].view(np.uint64), bs), "bit-score bit patterns differ"
        ref_ops = r["tb_ops"]
        for i in range(want.shape[0]):
            w = ref_ops[want[i, 13]:want[i, 13] + want[i, 14]]
            g = ops[got["esp_off"][i]:got["esp_off"][i] + got["esp_n"][i]]
            assert np.array_equal(g["op_type"], w[:, 0]) and np.array_equal(g["num"], w[:, 1]), f"edit script of HSP {i}"
    finally:
        Q.free(); V.free()


@pytest.mark.parametrize("name", cases.IDENTITY_FILTER_CASES)
def test_traceback_identity_and_length_filter(name):
    """hit_options->percent_identity / min_hit_length: the traceback stage drops the same HSPs as the reference
    (Blast_HSPTest) and the filter is not vacuous in these cases (the unfiltered run keeps more)."""
    from oracle import refdriver as R
    task, cfgkw, vol, qs = cases.make_case(name)
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so did not travel to this box")
    test_traceback_search_matches_reference(name)
    with_filter = R.search(qs, vol, R.default_config(task, taps=R.TAP_TRACEBACK, prelim_only=0, **cfgkw))
    plain = {k: v for k, v in cfgkw.items() if k not in ("percent_identity", "min_hit_length")}
    without = R.search(qs, vol, R.default_config(task, taps=R.TAP_TRACEBACK, prelim_only=0, **plain))
    assert 0 < with_filter["tb_final"].shape[0] < without["tb_final"].shape[0]


@pytest.mark.parametrize("name", ["mb_bridged_segments", "blastn_bridged_segments", "c4_scaled_short_reads",
                                  "blastn_bridged_perc_identity", "mb_bridged_perc_identity",
                                  "c5_scaled_ntlike_5kb", "c3_scaled_blastn_10kb", "mb_with_N"])
def test_full_search_product_path(name):
    """Preliminary stage + traceback stage, both on the GPU path with the product's own set-up (no reference data in
    the loop): the final results equal the reference's Blast_RunPreliminarySearch + Blast_RunTracebackSearch."""
    from gblastn_b200 import engine as E, setup as S
    from oracle import refdriver as R
    task, cfgkw, vol, qs = cases.make_case(name)
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so did not travel to this box")
    r = R.search(qs, vol, R.default_config(task, taps=R.TAP_TRACEBACK, prelim_only=0, **cfgkw))
    assert r["status"] == 0
    want = r["tb_final"]
    s = S.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs, device_lookup=1, **cfgkw)
    V, Q = E.Volume(vol), E.Query(s.batch)
    try:
        assert s.gap_x_dropoff_final() == r["gap_x_dropoff_final"]
        g = E.prelim_search(V, Q)
        got, ops = E.traceback_search(V, Q, s.gap_x_dropoff_final(), g["hsps"])
        assert got.shape[0] == want.shape[0] and want.shape[0] > 0
        for k, col in enumerate(("query_index", "oid", "context", "q_off", "q_end", "s_off", "s_end", "score", "num_ident")):
            assert np.array_equal(got[col], want[:, k]), col
        ev = want[:, 9].astype(np.uint32).astype(np.uint64) | (want[:, 10].astype(np.uint32).astype(np.uint64) << np.uint64(32))
        bs = want[:, 11].astype(np.uint32).astype(np.uint64) | (want[:, 12].astype(np.uint32).astype(np.uint64) << np.uint64(32))
        assert np.array_equal(got["evalue"].view(np.uint64), ev) and np.array_equal(got["bit_score"].view(np.uint64), bs)
        ref_ops = r["tb_ops"]
        flat = np.concatenate([ref_ops[want[i, 13]:want[i, 13] + want[i, 14]] for i in range(want.shape[0])])
        mine = np.concatenate([np.stack([ops["op_type"][a:a + n], ops["num"][a:a + n]], axis=1)
                               for a, n in zip(got["esp_off"], got["esp_n"])])
        assert np.array_equal(flat, mine), "edit scripts differ"
    finally:
        Q.free(); V.free(); s.free()


@pytest.mark.parametrize("name", cases.TRACEBACK_LIST_CASES)
def test_full_search_against_golden_fixture(name):
    """The same product path against the committed fixture tests/golden/traceback_<case>.npz (the reference's final
    results, written by tests/golden/make_traceback_golden.py): needs neither /root/reference nor oracle/_ref."""
    import os
    from gblastn_b200 import engine as E, setup as S
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"traceback_{name}.npz"))
    want, ref_ops = gold["tb_final"], gold["tb_ops"]
    task, cfgkw, vol, qs = cases.make_case(name)
    s = S.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs, device_lookup=1, **cfgkw)
    V, Q = E.Volume(vol), E.Query(s.batch)
    try:
        assert s.gap_x_dropoff_final() == int(gold["gap_x_dropoff_final"])
        g = E.prelim_search(V, Q)
        from oracle import portdriver as P
        assert np.array_equal(P.final_table(g["hsps"]), gold["prelim_final"])
        got, ops = E.traceback_search(V, Q, s.gap_x_dropoff_final(), g["hsps"])
        assert got.shape[0] == want.shape[0]
        for k, col in enumerate(("query_index", "oid", "context", "q_off", "q_end", "s_off", "s_end", "score", "num_ident")):
            assert np.array_equal(got[col], want[:, k]), col
        ev = want[:, 9].astype(np.uint32).astype(np.uint64) | (want[:, 10].astype(np.uint32).astype(np.uint64) << np.uint64(32))
        bs = want[:, 11].astype(np.uint32).astype(np.uint64) | (want[:, 12].astype(np.uint32).astype(np.uint64) << np.uint64(32))
        assert np.array_equal(got["evalue"].view(np.uint64), ev) and np.array_equal(got["bit_score"].view(np.uint64), bs)
        flat = np.concatenate([ref_ops[want[i, 13]:want[i, 13] + want[i, 14]] for i in range(want.shape[0])])
        mine = np.concatenate([np.stack([ops["op_type"][a:a + n], ops["num"][a:a + n]], axis=1)
                               for a, n in zip(got["esp_off"], got["esp_n"])])
        assert np.array_equal(flat, mine), "edit scripts differ"
    finally:
        Q.free(); V.free(); s.free()


def _random_start_items(r, vol, rng, per_hsp=3, max_hsps=60):
    """Start points for the differential test: inside real HSPs (with a small diagonal jitter), with and
    without a subject window, plus the corners of the sequences."""
    fin = r["final"]
    qlen = r["ctx_query_length"]
    rows = []
    pick = rng.permutation(fin.shape[0])[:max_hsps]
    for k in pick:
        oid, ctx, q_off, q_end, s_off = (int(x) for x in fin[k, :5])
        slen = int(vol.seq_len[oid])
        for _ in range(per_hsp):
            q = int(rng.integers(q_off, q_end))
            s = min(max(s_off + (q - q_off) + int(rng.integers(-2, 3)), 0), slen - 1)
            if rng.random() < 0.5:
                rows.append((oid, ctx, 0, slen, q, s))
            else:
                shift = max(0, s - int(rng.integers(50, 3000)))
                length = min(slen - shift, (s - shift) + int(rng.integers(50, 3000)))
                rows.append((oid, ctx, shift, length, q, s - shift))
        ql = int(qlen[ctx])
        rows.append((oid, ctx, 0, slen, 0, int(rng.integers(0, slen))))
        rows.append((oid, ctx, 0, slen, ql - 1, int(rng.integers(0, slen))))
        rows.append((oid, ctx, 0, slen, int(rng.integers(0, ql)), 0))
        rows.append((oid, ctx, 0, slen, int(rng.integers(0, ql)), slen - 1))
        rows.append((oid, ctx, 0, slen, 0, 0))
        rows.append((oid, ctx, 0, slen, ql - 1, slen - 1))
    return np.array(rows, dtype=np.int32)


@pytest.mark.parametrize("name", ["blastn_mb11_dp", "c3_scaled_blastn_10kb", "blastn_ws7_array", "mb_lut11_hash_indels",
                                  "c5_scaled_ntlike_5kb", "blastn_ws11_greedy", "mb_with_N", "mb_affine_greedy_5_2",
                                  "blastn_affine_greedy_5_2", "mb_affine_greedy_0_2"])
def test_gapped_traceback_random_starts(name):
    """Differential test on start points the search itself never produces: points anywhere inside real HSPs
    (off the optimal diagonal, in narrow subject windows) and the corners of both sequences, against the
    reference's BLAST_GappedAlignmentWithTraceback / BLAST_GreedyGappedAlignment called directly."""
    from gblastn_b200 import engine as E, abi
    from oracle import refdriver as R, portdriver as P
    task, cfgkw, vol, qs = cases.make_case(name)
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so did not travel to this box")
    cfg = R.default_config(task, taps=R.TAP_LUT, **cfgkw)
    r = R.search(qs, vol, cfg)
    assert r["status"] == 0 and r["final"].shape[0] > 0
    it = _random_start_items(r, vol, np.random.default_rng(7))
    rc = R.traceback_calls(qs, vol, it, cfg)
    assert rc["status"] == 0
    calls = rc["tb_calls"]
    assert calls.shape[0] == it.shape[0]
    h = P.batch_from_reference(r, task=task, cfg=cfg)
    V, Q = E.Volume(vol), E.Query(h)
    try:
        items = np.zeros(it.shape[0], dtype=abi.TB_ITEM_DTYPE)
        for k, col in enumerate(("oid", "context", "s_shift", "s_length", "q_start", "s_start")):
            items[col] = it[:, k]
        res, ops = E.gapped_traceback(V, Q, int(r["gap_x_dropoff_final"]), items)
        for k, col in ((8, "score"), (9, "query_start"), (10, "query_stop"), (11, "subject_start"), (12, "subject_stop"),
                       (14, "esp_n")):
            bad = np.flatnonzero(res[col] != calls[:, k])
            assert bad.size == 0, f"{col} differs for {bad.size} of {calls.shape[0]} calls, first {bad[:3]}: " \
                                  f"{res[col][bad[:3]]} vs {calls[bad[:3], k]} items {it[bad[:3]]}"
        ref_ops = rc["tb_ops"]
        for i in range(calls.shape[0]):
            want = ref_ops[calls[i, 13]:calls[i, 13] + calls[i, 14]]
            got = ops[res["esp_off"][i]:res["esp_off"][i] + res["esp_n"][i]]
            assert np.array_equal(got["op_type"], want[:, 0]) and np.array_equal(got["num"], want[:, 1]), \
                f"edit script differs for call {i} {it[i]}"
    finally:
        Q.free(); V.free()


def test_file_volume_equals_memory_volume(tmp_path):
    """bn_db_load_files: a volume written as .nin/.nsq and loaded from the files gives the same bytes
    of results as the same volume loaded from memory (ragged lengths, many subjects)."""
    from gblastn_b200 import engine as E, setup as S
    task, cfgkw, vol, qs = cases.make_case("mb_ntlike_many_subjects")
    nin, nsq = str(tmp_path / "v.nin"), str(tmp_path / "v.nsq")
    E.dbfile_write(nin, nsq, vol)
    s = S.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs, device_lookup=1, **cfgkw)
    V1, V2, Q = E.Volume(vol), E.FileVolume(nin, nsq), E.Query(s.batch)
    try:
        a, b = E.prelim_search(V1, Q), E.prelim_search(V2, Q)
        assert a["hsps"].size > 0 and a["hsps"].tobytes() == b["hsps"].tobytes()
        assert a["stats"]["lookup_hits"] == b["stats"]["lookup_hits"]
    finally:
        Q.free(); V1.free(); V2.free(); s.free()


@pytest.mark.parametrize("task", ["megablast", "blastn"])
def test_real_blast_db_volume(task):
    """A real BLAST DB volume from the reference's test data (tests/golden/ntshort, 7 sequences with
    ambiguity data): queries cut from its sequences, GPU search on the file-loaded volume == reference
    engine on the same bytes."""
    import os
    from gblastn_b200 import engine as E, setup as S, synth
    from oracle import refdriver as R, portdriver as P
    if not R.available():
        pytest.skip("reference library not present")
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    nin, nsq = os.path.join(gold, "ntshort.nin"), os.path.join(gold, "ntshort.nsq")
    info, off, ln = E.dbfile_index(nin, nsq)
    raw = np.fromfile(nsq, dtype=np.uint8)
    vol = synth.Volume(packed=np.concatenate([raw, np.zeros(32, np.uint8)]), byte_off=off, seq_len=ln)
    rng = np.random.default_rng(3)
    qs = []
    for i in range(info["n_seq"]):
        b = vol.bases(i)
        a = int(rng.integers(0, max(1, len(b) - 200)))
        q = b[a: a + 200].copy()
        mut = rng.random(q.size) < 0.03
        q[mut] = (q[mut] + 1 + rng.integers(0, 3, int(mut.sum()))) % 4
        qs.append(q if i % 2 == 0 else (3 - q[::-1]))
    r = R.search(qs, vol, R.default_config(task))
    assert r["status"] == 0 and r["final"].shape[0] > 0
    s = S.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs)
    V, Q = E.FileVolume(nin, nsq), E.Query(s.batch)
    try:
        g = E.prelim_search(V, Q)
        assert np.array_equal(P.final_table(g["hsps"]), r["final"])
    finally:
        Q.free(); V.free(); s.free()


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


@pytest.mark.parametrize("name,kw", [("mb_lut11_hash_indels", {}), ("mb_lut12_stride17", {}), ("blastn_mb11_dp", {}),
                                     ("mb_smallna_diagarray", {}), ("mb_ntlike_many_subjects", {}),
                                     ("blastn_ws7_na_table", {}), ("mb_two_hit_w40_hash", {})])
@pytest.mark.parametrize("mask_type", [1, 2])
def test_gpu_with_database_masks(name, kw, mask_type):
    """Soft / hard database masks (BLAST_SequenceBlk::seq_ranges): chunking at hard masks, per-range scanning
    from left + (word - lut), range-bounded mini-extension and s_TypeOfWord == reference engine."""
    from gblastn_b200 import engine as E, abi
    from oracle import refdriver as R, portdriver as P
    if not R.available():
        pytest.skip("reference library not present")
    task, cfgkw, vol, qs = cases.make_case(name)
    sm = _db_masks(vol, np.random.default_rng(100 * mask_type + len(name)), 5)
    cfg = R.default_config(task, taps=R.TAP_INIT | R.TAP_GAPPED | R.TAP_LUT, **cfgkw)
    r = R.search(qs, vol, cfg, subject_masks=sm, subject_mask_type=mask_type)
    assert r["status"] == 0
    h = P.batch_from_reference(r, task=task, cfg=cfg)
    V, Q = E.Volume(vol), E.Query(h)
    try:
        V.set_masks(sm, mask_type)
        g = E.prelim_search(V, Q, taps=abi.BN_TAP_INIT | abi.BN_TAP_GAPPED)
        assert np.array_equal(P.init_table(g["init"]), r["init"]), "init-HSPs differ from reference"
        assert np.array_equal(P.gapped_table(g["gapped"]), r["gapped"])
        assert np.array_equal(P.final_table(g["hsps"]), r["final"])
        assert g["stats"]["lookup_hits"] == r["lookup_hits"]
        # masks off again: the unmasked answer comes back
        V.set_masks(None)
        r0 = R.search(qs, vol, R.default_config(task, **cfgkw))
        g0 = E.prelim_search(V, Q)
        assert np.array_equal(P.final_table(g0["hsps"]), r0["final"])
    finally:
        Q.free(); V.free()


@pytest.mark.parametrize("name", ["c1_megablast_10kb_vs_1mb", "mb_lut11_hash_indels", "mb_lut12_stride17",
                                  "c4_scaled_short_reads", "c5_scaled_ntlike_5kb", "mb_bridged_segments",
                                  "mb_ntlike_many_subjects", "mb_with_N", "mb_repeat_family_hitlist5"])
@pytest.mark.parametrize("mask_type", [0, 1, 2])
def test_dense_signature_scan_changes_nothing(name, mask_type, monkeypatch):
    """Megablast tables with word - lut >= 7 carry the dense signature table (DevQuery::psig: one 2-byte gather per scan
    position answers "cell occupied?" and "can its mini-extension reach the word?").  The queue-driven kernel with the
    dense probe (words formed 8 consecutive positions per thread, or position by position: BN_NO_CONSEC) and with the
    {presence, rank} + signature-by-rank probe (BN_NO_PSIG, read when the batch is loaded) give the same seed hits and
    counters, bit for bit at every tap, with and without database masks (soft: scan units with p_first > 0; hard: chunks
    split at the masks), and equal the reference."""
    from gblastn_b200 import engine as E, abi
    from oracle import refdriver as R, portdriver as P
    if not R.available():
        pytest.skip("reference library not present")
    task, cfgkw, vol, qs = cases.make_case(name)
    sm = _db_masks(vol, np.random.default_rng(7 * mask_type + len(name)), 5) if mask_type else None
    cfg = R.default_config(task, taps=R.TAP_INIT | R.TAP_GAPPED | R.TAP_LUT, **cfgkw)
    r = R.search(qs, vol, cfg, subject_masks=sm, subject_mask_type=mask_type) if mask_type else R.search(qs, vol, cfg)
    assert r["status"] == 0
    h = P.batch_from_reference(r, task=task, cfg=cfg)
    assert h.batch.lut_type == abi.BN_LUT_MB and h.batch.word_length - h.batch.lut_word_length >= 7
    monkeypatch.setenv("BN_FILT_MAX", "0")          # small batches: not the shared-memory filter kernel
    V = E.Volume(vol)
    out = []
    try:
        if mask_type:
            V.set_masks(sm, mask_type)
        for mode in ("dense", "rank_probe", "dense_strided"):
            for k in ("BN_NO_PSIG", "BN_NO_CONSEC"):
                monkeypatch.delenv(k, raising=False)
            if mode == "rank_probe":
                monkeypatch.setenv("BN_NO_PSIG", "1")
            if mode == "dense_strided":
                monkeypatch.setenv("BN_NO_CONSEC", "1")
            Q = E.Query(h)
            try:
                out.append(E.prelim_search(V, Q, taps=abi.BN_TAP_INIT | abi.BN_TAP_GAPPED))
            finally:
                Q.free()
        a = out[0]
        for b in out[1:]:
            assert a["init"].tobytes() == b["init"].tobytes() and a["gapped"].tobytes() == b["gapped"].tobytes()
            assert a["hsps"].tobytes() == b["hsps"].tobytes()
            assert a["stats"]["lookup_hits"] == b["stats"]["lookup_hits"] == r["lookup_hits"]
        assert np.array_equal(P.init_table(a["init"]), r["init"])
        assert np.array_equal(P.final_table(a["hsps"]), r["final"])
    finally:
        V.free()


def test_batch_pipeline_equals_single_searches():
    """bn_prelim_search_batches: five different query batches (mixed table shapes, one empty result, one on
    the general path) through the two-stage pipeline give byte-identical results to one search each, and each
    equals the reference engine's result for that batch."""
    from gblastn_b200 import engine as E, setup as S, synth
    from oracle import refdriver as R, portdriver as P
    vol = synth.random_volume([400_000, 150_000, 30_000, 777], seed=61)
    all_qs = []
    specs = [dict(task="megablast", nq=120, qlen=900, seed=1, planted=0.7, sub=0.02),
             dict(task="megablast", nq=3, qlen=600, seed=2, planted=1.0, sub=0.04),          # small table, diag array
             dict(task="blastn", nq=8, qlen=700, seed=3, planted=0.8, sub=0.08),             # DP, many seed hits
             dict(task="megablast", nq=5, qlen=400, seed=4, planted=0.0, sub=0.0),           # no hits
             dict(task="megablast", nq=300, qlen=1000, seed=5, planted=0.5, sub=0.02)]       # lut 12
    setups = []
    for sp in specs:
        qs = synth.planted_queries(vol, sp["nq"], sp["qlen"], seed=sp["seed"], planted_frac=sp["planted"],
                                   sub_rate=sp["sub"], indel_rate=0.003)
        all_qs.append(qs)
        setups.append(S.Setup(qs, task=sp["task"], db_length=vol.total_bases, db_num_seqs=vol.n_seqs,
                              device_lookup=1 if sp["task"] == "megablast" else 0))
    V = E.Volume(vol)
    try:
        singles = []
        for st in setups:
            Q = E.Query(st.batch)
            singles.append(E.prelim_search(V, Q))
            Q.free()
        for _ in range(2):
            piped = E.prelim_search_batches(V, [st.batch for st in setups])
            assert len(piped) == len(singles)
            for a, b in zip(singles, piped):
                assert a["hsps"].tobytes() == b["hsps"].tobytes()
            # ... and every batch equals the reference engine run on that batch alone
            if R.available():
                for sp, qs, b in zip(specs, all_qs, piped):
                    r = R.search(qs, vol, R.default_config(sp["task"]))
                    assert r["status"] == 0
                    assert np.array_equal(P.final_table(b["hsps"]), r["final"]), f"pipelined batch differs from reference: {sp}"
                    assert b["stats"]["lookup_hits"] == r["lookup_hits"]
                    assert b["stats"]["gap_extensions"] == r["gap_extensions"]
                for k in ("lookup_hits", "good_init_extends", "gap_extensions", "good_extensions"):
                    assert a["stats"][k] == b["stats"][k]
        assert sum(x["hsps"].size for x in singles) > 0 and singles[3]["hsps"].size == 0
    finally:
        V.free()
        for st in setups:
            st.free()


def _assert_tb_equals_reference(got, ops, r):
    want, ref_ops = r["tb_final"], r["tb_ops"]
    assert got.shape[0] == want.shape[0]
    for k, col in enumerate(("query_index", "oid", "context", "q_off", "q_end", "s_off", "s_end", "score", "num_ident")):
        assert np.array_equal(got[col], want[:, k]), col
    ev = want[:, 9].astype(np.uint32).astype(np.uint64) | (want[:, 10].astype(np.uint32).astype(np.uint64) << np.uint64(32))
    bs = want[:, 11].astype(np.uint32).astype(np.uint64) | (want[:, 12].astype(np.uint32).astype(np.uint64) << np.uint64(32))
    assert np.array_equal(got["evalue"].view(np.uint64), ev) and np.array_equal(got["bit_score"].view(np.uint64), bs)
    if want.shape[0]:
        flat = np.concatenate([ref_ops[want[i, 13]:want[i, 13] + want[i, 14]] for i in range(want.shape[0])])
        mine = np.concatenate([np.stack([ops["op_type"][a:a + n], ops["num"][a:a + n]], axis=1)
                               for a, n in zip(got["esp_off"], got["esp_n"])])
        assert np.array_equal(flat, mine), "edit scripts differ"


def test_job_pipeline_equals_reference():
    """bn_prelim_search_jobs (prepare -> device -> host replay -> traceback, software-pipelined over two lanes): a mixed
    stream of jobs — resident and host-side volumes, resident and host-side batches, megablast and blastn, fused and
    general path, an empty result — returns per job exactly what the reference's preliminary + traceback stages return
    for that (volume, batch) pair alone; run twice (steady state re-uses the lanes' buffers)."""
    from gblastn_b200 import engine as E, setup as S
    from oracle import refdriver as R, portdriver as P
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so did not travel to this box")
    names = ["mb_lut12_stride17", "blastn_mb11_dp", "mb_smallna_diagarray", "mb_no_hits", "c4_scaled_short_reads",
             "mb_bridged_segments", "blastn_bridged_perc_identity", "mb_lut11_hash_indels"]
    jobs, refs, keep = [], [], []
    try:
        for k, name in enumerate(names):
            task, cfgkw, vol, qs = cases.make_case(name)
            r = R.search(qs, vol, R.default_config(task, taps=R.TAP_TRACEBACK, prelim_only=0, **cfgkw))
            assert r["status"] == 0
            st = S.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs,
                         device_lookup=1 if task == "megablast" else 0, **cfgkw)
            keep.append(st)
            j = {"gap_x_dropoff_final": st.gap_x_dropoff_final()}
            if k % 2 == 0:
                j["host_volume"] = vol
            else:
                V = E.Volume(vol); keep.append(V); j["volume"] = V
            if k % 3 == 0:
                Q = E.Query(st.batch); keep.append(Q); j["query"] = Q
            else:
                j["batch"] = st.batch
            jobs.append(j); refs.append(r)
        for _ in range(2):
            out = E.prelim_search_jobs(jobs, traceback=True)
            assert len(out) == len(jobs)
            for name, o, r in zip(names, out, refs):
                assert np.array_equal(P.final_table(o["hsps"]), r["final"]), f"preliminary lists of {name}"
                assert o["stats"]["lookup_hits"] == r["lookup_hits"] and o["stats"]["gap_extensions"] == r["gap_extensions"]
                _assert_tb_equals_reference(o["tb"][0], o["tb"][1], r)
        plain = E.prelim_search_jobs(jobs)              # without the traceback stage
        for a, b in zip(out, plain):
            assert a["hsps"].tobytes() == b["hsps"].tobytes() and "tb" not in b
        assert sum(o["hsps"].size for o in out) > 0 and out[3]["hsps"].size == 0
    finally:
        for x in keep:
            x.free()


def test_job_pipelines_on_concurrent_threads():
    """Three caller threads each run a job pipeline WITH its traceback stage on the same device (4 lanes each, 6 on the
    device): lanes are handed out all-or-nothing, so the calls queue instead of deadlocking on partial sets, and every
    job still equals the single calls."""
    import threading
    from gblastn_b200 import engine as E, setup as S
    task, cfgkw, vol, qs = cases.make_case("mb_lut11_hash_indels")
    st = S.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs, device_lookup=1, **cfgkw)
    V, Q = E.Volume(vol), E.Query(st.batch)
    try:
        want = E.prelim_search(V, Q)
        want_tb, want_ops = E.traceback_search(V, Q, st.gap_x_dropoff_final(), want["hsps"])
        jobs = [{"volume": V, "query": Q, "gap_x_dropoff_final": st.gap_x_dropoff_final()},
                {"host_volume": vol, "batch": st.batch, "gap_x_dropoff_final": st.gap_x_dropoff_final()}] * 4
        out, errs = {}, []

        def run(i):
            try:
                out[i] = E.prelim_search_jobs(jobs, traceback=True)
            except Exception as e:          # noqa: BLE001
                errs.append(e)
        threads = [threading.Thread(target=run, args=(i,)) for i in range(3)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=120)
        assert not any(t.is_alive() for t in threads), "job pipelines deadlocked"
        assert not errs, errs
        for i in range(3):
            for o in out[i]:
                assert o["hsps"].tobytes() == want["hsps"].tobytes()
                assert o["tb"][0].tobytes() == want_tb.tobytes() and o["tb"][1].tobytes() == want_ops.tobytes()
    finally:
        Q.free(); V.free(); st.free()


def test_job_pipeline_reports_errors():
    from gblastn_b200 import engine as E, abi
    r, h, vol = _setup("mb_lut11_hash_indels")
    V = E.Volume(vol)
    try:
        with pytest.raises(E.BnError):
            E.prelim_search_jobs([{"volume": V, "batch": h}, {"volume": V, "query": type("Q", (), {"handle": 9999})()}])
        ok = E.prelim_search_jobs([{"volume": V, "batch": h}])          # the device is usable afterwards
        assert ok[0]["hsps"].size == r["final"].shape[0]
    finally:
        V.free()


def test_host_buffer_entry_point():
    from gblastn_b200 import engine as E
    from oracle import portdriver as P
    r, h, vol = _setup("mb_lut11_hash_indels")
    g = E.prelim_search_host(h, vol)
    assert np.array_equal(P.final_table(g["hsps"]), r["final"])


def _masks_for(qs, seed):
    rng = np.random.default_rng(seed)
    masks = []
    for q in qs:
        m = []
        if rng.random() < 0.7:
            a = int(rng.integers(0, len(q) - 90))
            m.append((a, a + int(rng.integers(12, 80))))
        if rng.random() < 0.3:
            b0 = int(rng.integers(0, len(q) - 40))
            m.append((b0, b0 + int(rng.integers(5, 35))))
        masks.append(m)
    return masks


@pytest.mark.parametrize("name", ["mb_lut11_hash_indels", "mb_lut12_stride17", "mb_smallna_diagarray",
                                  "c5_scaled_ntlike_5kb"])
def test_gpu_with_masked_queries(name):
    """mask-at-hash query masks: lut->masked_locations != NULL switches on the lookup re-probing of
    s_TypeOfWord (core/na_ungapped.c:489-588) in the diagonal stage."""
    from gblastn_b200 import engine as E, abi
    from oracle import refdriver as R, portdriver as P
    if not R.available():
        pytest.skip("reference library not present")
    task, cfgkw, vol, qs = cases.make_case(name)
    masks = _masks_for(qs, 77)
    cfg = R.default_config(task, taps=R.TAP_INIT | R.TAP_GAPPED | R.TAP_LUT, **cfgkw)
    r = R.search(qs, vol, cfg, masks=masks)
    assert r["status"] == 0 and r["n_masked_locations"] > 0
    h = P.batch_from_reference(r, task=task, cfg=cfg)
    V, Q = E.Volume(vol), E.Query(h)
    try:
        g = E.prelim_search(V, Q, taps=abi.BN_TAP_INIT | abi.BN_TAP_GAPPED)
        assert np.array_equal(P.init_table(g["init"]), r["init"])
        assert np.array_equal(P.final_table(g["hsps"]), r["final"])
    finally:
        Q.free(); V.free()


def test_gpu_na_table_word_longer_than_lut():
    """eNaLookupTable with word 11 > lut 8 (a long, mostly masked query: few table entries but offsets beyond
    15 bits): s_BlastNaScanSubject_8_4 + s_BlastNaExtendAligned + s_TypeOfWord over s_NaLookup."""
    from gblastn_b200 import engine as E, setup as S, synth, abi
    from oracle import refdriver as R, portdriver as P
    if not R.available():
        pytest.skip("reference library not present")
    vol = synth.random_volume([80_000, 30_000], seed=23)
    qs = synth.planted_queries(vol, 1, 40_000, seed=33, planted_frac=1.0, sub_rate=0.06, indel_rate=0.005)
    masks = [[(0, 17_000), (19_000, 36_500), (38_000, 39_999)]]
    r = R.search(qs, vol, R.default_config("blastn", taps=R.TAP_INIT), masks=masks)
    assert r["status"] == 0 and r["lut_type"] == 2 and (r["lut_word_length"], r["word_length"], r["scan_step"]) == (8, 11, 4)
    s = S.Setup(qs, task="blastn", db_length=vol.total_bases, db_num_seqs=vol.n_seqs, masks=masks)
    V, Q = E.Volume(vol), E.Query(s.batch)
    try:
        g = E.prelim_search(V, Q, taps=abi.BN_TAP_INIT)
        assert np.array_equal(P.init_table(g["init"]), r["init"])
        assert np.array_equal(P.final_table(g["hsps"]), r["final"])
    finally:
        Q.free(); V.free(); s.free()


@pytest.mark.parametrize("name", ["mb_lut11_hash_indels", "blastn_mb11_dp", "mb_smallna_diagarray", "mb_lut12_stride17"])
def test_gpu_with_product_setup(name):
    """Whole product path: our own set-up (bn_setup_*) + GPU search == reference blastn engine."""
    from gblastn_b200 import engine as E, setup as S
    from oracle import refdriver as R, portdriver as P
    if not R.available():
        pytest.skip("reference library not present")
    task, cfgkw, vol, qs = cases.make_case(name)
    r = R.search(qs, vol, R.default_config(task, **cfgkw))
    s = S.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs, **cfgkw)
    V, Q = E.Volume(vol), E.Query(s.batch)
    try:
        g = E.prelim_search(V, Q)
        assert np.array_equal(P.final_table(g["hsps"]), r["final"])
    finally:
        Q.free(); V.free(); s.free()


@pytest.mark.parametrize("name", ["c1_megablast_10kb_vs_1mb", "mb_lut11_hash_indels", "mb_lut12_stride17",
                                  "mb_with_N", "blastn_mb11_dp", "mb_ws16", "mb_ntlike_many_subjects"])
def test_gpu_device_lookup_fill(name):
    """s_FillContigMBTable on the device (bn_query_load without hashtable/next_pos): the table is the
    reference's bit for bit, with and without query masks, and the search results do not change."""
    from gblastn_b200 import engine as E, setup as S
    from oracle import refdriver as R, portdriver as P
    if not R.available():
        pytest.skip("reference library not present")
    task, cfgkw, vol, qs = cases.make_case(name)
    for masks in (None, _masks_for(qs, 5)):
        cfg = R.default_config(task, taps=R.TAP_LUT, **cfgkw)
        r = R.search(qs, vol, cfg, masks=masks)
        assert r["status"] == 0
        s = S.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs, masks=masks,
                    device_lookup=1, **cfgkw)
        assert s.hashtable is None and s.batch.n_lookup_segments > 0
        V, Q = E.Volume(vol), E.Query(s.batch)
        try:
            ht, nx = E.download_lookup(Q)
            assert np.array_equal(ht, r["hashtable"]), "device-built hashtable differs from the reference"
            assert np.array_equal(nx, r["next_pos"]), "device-built next_pos differs from the reference"
            g = E.prelim_search(V, Q)
            assert np.array_equal(P.final_table(g["hsps"]), r["final"])
        finally:
            Q.free(); V.free(); s.free()


def test_gpu_full_size_properties():
    """BASELINE configs[1] at full size (1000 x 1 kb vs 250 Mb): size-independent properties.
    * determinism / idempotence: two searches of the same resident inputs give identical bytes
    * every HSP is a real local alignment: end points inside the sequences, score >= cutoff,
      the ungapped seed region around (q_gapped_start, s_gapped_start) matches exactly
    * planted queries are found: >= 95 % of planted queries report an HSP covering >= 90 % of the query
    * subject coordinates are absolute (second 50 Mb chunk starts at 199 999 900)"""
    from gblastn_b200 import engine as E, setup as S, synth
    vol = synth.random_volume([250_000_000], seed=2)
    qs = synth.planted_queries(vol, 1000, 1000, seed=22, planted_frac=0.8, sub_rate=0.02)
    s = S.Setup(qs, task="megablast", db_length=vol.total_bases, db_num_seqs=vol.n_seqs)
    V, Q = E.Volume(vol), E.Query(s.batch)
    try:
        g1 = E.prelim_search(V, Q)
        g2 = E.prelim_search(V, Q)
        h = g1["hsps"]
        assert h.tobytes() == g2["hsps"].tobytes()
        assert g1["stats"]["subject_bases_scanned"] == 250_000_100       # 200 Mb + 50 000 100 (100-base overlap)
        ctx = s.contexts()
        cutoff = np.array([c.gapped_cutoff for c in ctx])
        qlen = np.array([c.query_length for c in ctx])
        assert (h["score"] >= cutoff[h["context"]]).all()
        assert (h["q_off"] >= 0).all() and (h["q_end"] <= qlen[h["context"]]).all()
        assert (h["s_off"] >= 0).all() and (h["s_end"] <= 250_000_000).all()
        assert (h["q_off"] < h["q_end"]).all() and (h["s_off"] < h["s_end"]).all()
        assert (h["s_end"] > 200_000_000).any(), "no HSP in the second subject chunk"
        cq = s.concat_query[1:]
        subj = vol.bases(0) if False else None
        # exact-match check of 8 bases at the gapped start point (the greedy seed estimate lies
        # inside the longest run of matches)
        ok = 0
        for rec in h[:200]:
            c = ctx[int(rec["context"])]
            qpos = c.query_offset + int(rec["q_gapped_start"])
            spos = int(rec["s_gapped_start"])
            sb = vol.packed[spos // 4: spos // 4 + 4]
            un = np.stack([sb >> 6, (sb >> 4) & 3, (sb >> 2) & 3, sb & 3], axis=1).reshape(-1)[spos % 4: spos % 4 + 8]
            ok += int(np.array_equal(cq[qpos: qpos + 8], un))
        assert ok >= 190
        covered = set()
        for rec in h:
            if rec["q_end"] - rec["q_off"] >= 900:
                covered.add(int(rec["context"]) // 2)
        assert len(covered) >= 0.95 * 0.8 * 1000 * 0.98
    finally:
        Q.free(); V.free(); s.free()


def test_gpu_full_size_c2_equals_reference():
    """BASELINE configs[1] at full size (1000 x 1 kb vs one 250 Mb sequence, two subject chunks): the final lists
    with E-value bit patterns and the diagnostics counters equal the reference engine's, bit for bit."""
    from gblastn_b200 import engine as E, setup as S, synth
    from oracle import refdriver as R, portdriver as P
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so did not travel to this box")
    vol = synth.random_volume([250_000_000], seed=2)
    qs = synth.planted_queries(vol, 1000, 1000, seed=22, planted_frac=0.8, sub_rate=0.02)
    r = R.search(qs, vol, R.default_config("megablast", taps=R.TAP_INIT))
    assert r["status"] == 0 and r["final"].shape[0] > 700
    s = S.Setup(qs, task="megablast", db_length=vol.total_bases, db_num_seqs=vol.n_seqs, device_lookup=1)
    V, Q = E.Volume(vol), E.Query(s.batch)
    try:
        from gblastn_b200 import abi
        g = E.prelim_search(V, Q, taps=abi.BN_TAP_INIT)
        assert np.array_equal(P.init_table(g["init"]), r["init"]), "init-HSPs differ from reference"
        assert np.array_equal(P.final_table(g["hsps"]), r["final"]), "final lists / E-value bits differ"
        st = g["stats"]
        assert (st["lookup_hits"], st["good_init_extends"], st["gap_extensions"], st["good_extensions"]) == \
               (r["lookup_hits"], r["good_init_extends"], r["gap_extensions"], r["good_extensions"])
        assert (r["final"][:, 4] > 200_000_000).any(), "no HSP in the second subject chunk"
    finally:
        Q.free(); V.free(); s.free()


def _split_volume(vol, n_vol):
    """Contiguous OID ranges of a volume as volumes of their own (what a multi-volume database is)."""
    from gblastn_b200 import synth
    n = len(vol.seq_len)
    cuts = [n * k // n_vol for k in range(n_vol + 1)]
    out = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        lo = int(vol.byte_off[a])
        hi = int(vol.byte_off[b - 1]) + int(vol.seq_len[b - 1]) // 4 + 1
        packed = np.concatenate([vol.packed[lo:hi], np.zeros(synth.PAD_BYTES, np.uint8)])
        out.append(synth.Volume(packed, (vol.byte_off[a:b] - lo).astype(np.int64), vol.seq_len[a:b].copy()))
    return out


@pytest.mark.parametrize("name", ["mb_repeat_family_hitlist5", "mb_repeat_family_hitlist20", "blastn_repeat_family_hitlist5",
                                  "mb_ntlike_many_subjects", "blastn_ntlike_many_subjects"])
@pytest.mark.parametrize("n_vol", [2, 4])
def test_volumes_equal_one_reference_pass(name, n_vol):
    """bn_prelim_search_volumes: the database split into n_vol volumes (spread over the visible devices) and gathered
    on the host == ONE reference pass over the concatenated database, at every tap — the low_score feedback
    (core/blast_engine.c:1313-1320) crosses volume boundaries — and with prune_hitlists the lists returned are the
    ones the reference's HSP stream holds when the stage ends (prelim_hitlist_size applied once, after the gather)."""
    from gblastn_b200 import engine as E, abi
    from oracle import portdriver as P
    r, h, vol = _setup(name)
    parts = _split_volume(vol, n_vol)
    n_dev = E.device_count()
    Vs = [E.Volume(p, device=k % n_dev) for k, p in enumerate(parts)]
    Q = E.Query(h)
    try:
        g = E.prelim_search_volumes(Vs, Q, taps=abi.BN_TAP_INIT | abi.BN_TAP_GAPPED)
        assert np.array_equal(P.init_table(g["init"]), r["init"]), "init-HSPs differ from the single pass"
        assert np.array_equal(P.gapped_table(g["gapped"]), r["gapped"]), "gapped lists differ from the single pass"
        assert np.array_equal(P.final_table(g["hsps"]), r["final"]), "final lists differ from the single pass"
        st = g["stats"]
        assert (st["lookup_hits"], st["good_init_extends"], st["gap_extensions"], st["good_extensions"]) == \
               (r["lookup_hits"], r["good_init_extends"], r["gap_extensions"], r["good_extensions"])
        # the lists that reach the traceback stage
        gp = E.prelim_search_volumes(Vs, Q, prune_hitlists=True)
        ctx_q = np.arange(r["num_contexts"]) // 2
        got = gp["hsps"]
        kept = {(int(q), int(o)) for q, o in r["kept"][:, :2]}
        assert {(int(ctx_q[c]), int(o)) for c, o in zip(got["context"], got["oid"])} == kept
        fin = r["final"]
        sel = np.array([(int(ctx_q[c]), int(o)) in kept for o, c in zip(fin[:, 0], fin[:, 1])], dtype=bool)
        assert np.array_equal(P.final_table(got), fin[sel])
        if name.startswith(("mb_repeat", "blastn_repeat")):
            assert sel.sum() < fin.shape[0], "the case no longer overflows a hit list"
    finally:
        Q.free()
        for V in Vs:
            V.free()


def test_concurrent_callers_one_device():
    """Re-entrancy per GPU (SURVEY.md 8(b) threading; api/prelim_search_runner.hpp:94-113 runs the engine from
    num_threads threads): four threads search one resident volume with four different query batches at the same time,
    while a fifth keeps loading and freeing batches; every result equals the serial one."""
    import threading
    from gblastn_b200 import engine as E, setup as S, synth
    vol = synth.random_volume([600_000, 250_000, 40_000, 999], seed=81)
    specs = [dict(task="megablast", nq=150, qlen=800, seed=1), dict(task="blastn", nq=6, qlen=700, seed=2),
             dict(task="megablast", nq=3, qlen=600, seed=3), dict(task="megablast", nq=400, qlen=1000, seed=4)]
    setups = []
    for sp in specs:
        qs = synth.planted_queries(vol, sp["nq"], sp["qlen"], seed=sp["seed"], planted_frac=0.8, sub_rate=0.03, indel_rate=0.003)
        setups.append(S.Setup(qs, task=sp["task"], db_length=vol.total_bases, db_num_seqs=vol.n_seqs,
                              device_lookup=1 if sp["task"] == "megablast" else 0))
    V = E.Volume(vol)
    Qs = [E.Query(st.batch) for st in setups]
    try:
        serial = [E.prelim_search(V, Q) for Q in Qs]
        assert sum(x["hsps"].size for x in serial) > 0
        out = [[None] * 6 for _ in Qs]
        errors = []
        stop = threading.Event()

        def run(k):
            try:
                for it in range(6):
                    out[k][it] = E.prelim_search(V, Qs[k])
            except Exception as e:      # noqa: BLE001
                errors.append(e)

        def churn():
            try:
                while not stop.is_set():
                    q = E.Query(setups[0].batch)
                    q.free()
            except Exception as e:      # noqa: BLE001
                errors.append(e)

        ts = [threading.Thread(target=run, args=(k,)) for k in range(len(Qs))]
        tc = threading.Thread(target=churn)
        tc.start()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        stop.set()
        tc.join()
        assert not errors, errors
        for k in range(len(Qs)):
            for it in range(6):
                assert out[k][it]["hsps"].tobytes() == serial[k]["hsps"].tobytes()
                assert out[k][it]["stats"]["lookup_hits"] == serial[k]["stats"]["lookup_hits"]
    finally:
        for Q in Qs:
            Q.free()
        V.free()
        for st in setups:
            st.free()


def test_stage_entry_points_validate_their_input():
    """bn_get_gapped_score rejects init hits outside the query context / subject chunk (they would be read by the
    gapped kernels); bn_scan_subject selects one chunk of a split subject or rejects a chunk that does not exist."""
    from gblastn_b200 import engine as E, abi
    r, h, vol = _setup("mb_lut11_hash_indels")
    V, Q = E.Volume(vol), E.Query(h)
    try:
        rows = r["init"][(r["init"][:, 0] == 0) & (r["init"][:, 1] == 0)]
        arr = np.zeros(rows.shape[0], dtype=abi.INIT_DTYPE)
        for k, col in enumerate(("oid", "chunk_off", "q_off", "s_off", "q_start", "s_start", "length", "score")):
            arr[col] = rows[:, k]
        assert E.get_gapped_score(V, Q, 0, 0, arr).size > 0
        for field, value in (("s_off", int(vol.seq_len[0]) + 5), ("q_off", -1), ("q_off", 10 ** 9), ("length", 10 ** 7),
                             ("s_start", -3), ("q_start", int(r["ctx_query_offset"][1]) - 2)):
            bad = arr.copy()
            bad[field][0] = value
            with pytest.raises(E.BnError) as e:
                E.get_gapped_score(V, Q, 0, 0, bad)
            assert e.value.code == abi.BN_ERR_INVALID
        whole = E.scan_subject(V, Q, 0)
        one = E.scan_subject(V, Q, 0, 0, int(vol.seq_len[0]))
        assert whole.tobytes() == one.tobytes() and whole.size > 0
        with pytest.raises(E.BnError):
            E.scan_subject(V, Q, 0, 0, int(vol.seq_len[0]) - 1)
        with pytest.raises(E.BnError):
            E.scan_subject(V, Q, 0, 4, 0)
    finally:
        Q.free(); V.free()


@pytest.mark.parametrize("name", ["blastn_mb11_dp", "c3_scaled_blastn_10kb", "blastn_direct_mixed_lengths_N",
                                  "blastn_bridged_segments", "blastn_ws11_greedy"])
def test_direct_filter_changes_nothing(name, monkeypatch):
    """blastn mode (lookup word == word): the scan kernel drops the lookup hits whose ungapped extension cannot reach
    the cutoff before they are sorted and replayed.  With the filter switched off (BN_NO_DIRECT_FILTER, read when the
    batch is loaded) every tap is the same, bit for bit — and equal to the reference."""
    from gblastn_b200 import engine as E, abi
    from oracle import portdriver as P
    r, h, vol = _setup(name)
    assert r["lut_word_length"] == r["word_length"], "the case must run in direct mode"
    V = E.Volume(vol)
    out = []
    try:
        for mode in ("dense", "off", "inline", "warp_leaders"):
            monkeypatch.delenv("BN_NO_DIRECT_FILTER", raising=False)
            monkeypatch.delenv("BN_NO_DIRECT_DENSE", raising=False)
            monkeypatch.delenv("BN_NO_SCALAR_LEADERS", raising=False)
            if mode == "warp_leaders":                          # speculative ungapped extensions by whole warps (read per search)
                monkeypatch.setenv("BN_NO_SCALAR_LEADERS", "1")
            monkeypatch.setenv("BN_FILT_MAX", "0")              # the queue-driven kernel: the one with the dense evaluation
            if mode == "off":
                monkeypatch.setenv("BN_NO_DIRECT_FILTER", "1")
            if mode == "inline":                                # evaluations where the chain walk meets the hits (read per search)
                monkeypatch.setenv("BN_NO_DIRECT_DENSE", "1")
            Q = E.Query(h)
            try:
                out.append(E.prelim_search(V, Q, taps=abi.BN_TAP_INIT | abi.BN_TAP_GAPPED))
            finally:
                Q.free()
        a, b, c, d = out
        assert a["init"].tobytes() == b["init"].tobytes() and a["gapped"].tobytes() == b["gapped"].tobytes()
        assert a["hsps"].tobytes() == b["hsps"].tobytes()
        assert a["init"].tobytes() == c["init"].tobytes() and a["hsps"].tobytes() == c["hsps"].tobytes()
        assert a["init"].tobytes() == d["init"].tobytes() and a["hsps"].tobytes() == d["hsps"].tobytes()
        assert a["stats"]["lookup_hits"] == c["stats"]["lookup_hits"]
        assert np.array_equal(P.init_table(a["init"]), r["init"])
        assert np.array_equal(P.final_table(a["hsps"]), r["final"])
        assert a["stats"]["lookup_hits"] == b["stats"]["lookup_hits"] == r["lookup_hits"]
    finally:
        V.free()


@pytest.mark.parametrize("name", ["blastn_mb11_dp", "c3_scaled_blastn_10kb", "blastn_direct_mixed_lengths_N",
                                  "blastn_bridged_segments", "mb_lut12_stride17", "c4_scaled_short_reads",
                                  "mb_bridged_segments", "blastn_smallna_dp", "mb_long_divergent_tier2"])
def test_device_triage_changes_nothing(name, monkeypatch):
    """Large result sets are triaged on the device after the gapped stage: winners (extension >= cutoff) and the losers
    a winner's box could contain go to the host replay, every other loser arrives as a count.  Forced on here at small
    size (BN_FORCE_GENERAL + BN_TRIAGE_MIN) and compared with the untriaged path and with the reference: lists,
    E-value bits and the extension counters."""
    from gblastn_b200 import engine as E
    from oracle import portdriver as P
    r, h, vol = _setup(name)
    V, Q = E.Volume(vol), E.Query(h)
    try:
        monkeypatch.setenv("BN_FORCE_GENERAL", "1")
        monkeypatch.setenv("BN_TRIAGE_MIN", "1")
        monkeypatch.delenv("BN_NO_TRIAGE", raising=False)
        a = E.prelim_search(V, Q)
        monkeypatch.setenv("BN_NO_TRIAGE", "1")
        b = E.prelim_search(V, Q)
        assert a["hsps"].tobytes() == b["hsps"].tobytes()
        assert np.array_equal(P.final_table(a["hsps"]), r["final"])
        for st in (a["stats"], b["stats"]):
            assert (st["lookup_hits"], st["good_init_extends"], st["gap_extensions"], st["good_extensions"]) == \
                   (r["lookup_hits"], r["good_init_extends"], r["gap_extensions"], r["good_extensions"])
        assert a["stats"]["kernel_launches"] > b["stats"]["kernel_launches"], "the triage kernels did not run"
    finally:
        Q.free(); V.free()


@pytest.mark.parametrize("name", ["c3_scaled_blastn_10kb", "blastn_bridged_segments", "blastn_mb11_dp"])
@pytest.mark.parametrize("mode", ["rounds", "set_aside_all", "no_rounds"])
def test_long_extensions_in_rounds(name, mode, monkeypatch):
    """blastn mode: the long gapped extensions are made in rounds — per (chunk, context) the best pending one, the ones
    inside a box made so far set aside — and the host replay asks for any set-aside extension it needs after all.
    `set_aside_all` makes the prediction deliberately wrong (everything but the first box of a group is set aside) so the
    replay has to ask; `no_rounds` extends them all.  Same lists and counters as the reference every time."""
    from gblastn_b200 import engine as E
    from oracle import portdriver as P
    r, h, vol = _setup(name)
    V, Q = E.Volume(vol), E.Query(h)
    try:
        monkeypatch.setenv("BN_FORCE_GENERAL", "1")
        monkeypatch.setenv("BN_TRIAGE_MIN", "1")
        if mode == "set_aside_all":
            monkeypatch.setenv("BN_LONG_SET_ASIDE_ALL", "1")
        if mode == "no_rounds":
            monkeypatch.setenv("BN_NO_LONG_ROUNDS", "1")
        a = E.prelim_search(V, Q)
        assert np.array_equal(P.final_table(a["hsps"]), r["final"])
        st = a["stats"]
        assert (st["lookup_hits"], st["good_init_extends"], st["gap_extensions"], st["good_extensions"]) == \
               (r["lookup_hits"], r["good_init_extends"], r["gap_extensions"], r["good_extensions"])
    finally:
        Q.free(); V.free()


@pytest.mark.parametrize("name", ["c1_megablast_10kb_vs_1mb", "mb_lut11_hash_indels", "mb_lut12_stride17", "blastn_mb11_dp",
                                  "c3_scaled_blastn_10kb", "c4_scaled_short_reads", "c5_scaled_ntlike_5kb", "mb_bridged_segments",
                                  "blastn_direct_mixed_lengths_N", "mb_ntlike_many_subjects", "mb_with_N", "mb_ws16",
                                  "mb_repeat_family_hitlist5", "mb_two_hit_w40_hash"])
def test_filtered_scan_changes_nothing(name, monkeypatch):
    """Small megablast tables are scanned by scan_kernel_filtered (hashed presence filter in shared memory, one
    persistent CTA per SM).  Same seed hits as the queue-driven kernel (filter off: BN_FILT_MAX=0, read when the batch
    is loaded) and as the folded 2^19-bit map (BN_FILT_HALF): every tap bit for bit, the scan tap included, and equal
    to the reference."""
    from gblastn_b200 import engine as E, abi
    from oracle import portdriver as P
    r, h, vol = _setup(name)
    assert h.batch.lut_type == abi.BN_LUT_MB
    if h.batch.concat_len > (1 << 16):
        monkeypatch.setenv("BN_FILT_MAX", str(1 << 30))     # larger batches: forced, the filter is merely denser
    V = E.Volume(vol)
    out = []
    try:
        forced = h.batch.concat_len > (1 << 16)
        for mode in ("filtered", "half", "queue", "filtered_strided", "queue_strided"):
            if not forced:
                monkeypatch.delenv("BN_FILT_MAX", raising=False)
            monkeypatch.delenv("BN_FILT_HALF", raising=False)
            monkeypatch.delenv("BN_NO_CONSEC", raising=False)
            if mode.endswith("strided"):        # words formed position by position instead of 8 consecutive ones per thread
                monkeypatch.setenv("BN_NO_CONSEC", "1")
            if mode.startswith("queue"):
                monkeypatch.setenv("BN_FILT_MAX", "0")
            if mode == "half":
                monkeypatch.setenv("BN_FILT_HALF", "1")
            Q = E.Query(h)
            try:
                g = E.prelim_search(V, Q, taps=abi.BN_TAP_INIT | abi.BN_TAP_GAPPED)
                oid = int(np.argmax(vol.seq_len))
                g["pairs"] = E.scan_subject(V, Q, oid) if vol.seq_len[oid] <= 200_000_000 else None
                out.append(g)
            finally:
                Q.free()
        a = out[0]
        for b in out[1:]:
            assert a["init"].tobytes() == b["init"].tobytes() and a["gapped"].tobytes() == b["gapped"].tobytes()
            assert a["hsps"].tobytes() == b["hsps"].tobytes()
            assert a["stats"]["lookup_hits"] == b["stats"]["lookup_hits"] == r["lookup_hits"]
            if a["pairs"] is not None:
                assert a["pairs"].tobytes() == b["pairs"].tobytes()
        assert np.array_equal(P.init_table(a["init"]), r["init"])
        assert np.array_equal(P.final_table(a["hsps"]), r["final"])
    finally:
        V.free()


@pytest.mark.parametrize("n,bits", [(1, 8), (31, 9), (2047, 17), (2048, 24), (2049, 25), (65_537, 33), (1_000_003, 41),
                                    (6_000_000, 39), (300_000, 64)])
def test_own_radix_sort_and_prefix_sum(n, bits):
    """csrc/radix_sort.cu (the path's own device-wide sort and scans, no library kernels): seeded random pairs with many
    equal keys against a stable host sort, inclusive prefix sum against a host loop; sizes around the warp-tile and
    scan-tile boundaries and C3's seed-hit count."""
    import ctypes as C
    from gblastn_b200 import engine as E
    E.init(1)
    bad = C.c_int64(-1)
    rc = E.lib().bn_selftest_sort(C.c_int(0), C.c_int64(n), C.c_int(bits), C.c_uint64(7 + n), C.byref(bad))
    assert rc == 0 and bad.value == 0


def test_gapped_traceback_band_wider_than_shared_ring():
    """ALIGN_EX with a final X-drop whose band (2 X / gap_extend cells) exceeds the DP kernel's shared-memory ring of
    1024 cells: the batch is re-run with the rings in global memory (traceback_dp_kernel<32768, true>); same scores, end
    points and edit scripts as the reference's BLAST_GappedAlignmentWithTraceback, which has no such limit."""
    from gblastn_b200 import engine as E, abi
    from oracle import refdriver as R, portdriver as P
    task, cfgkw, vol, qs = cases.make_case("blastn_mb11_dp")
    if not R.available():
        pytest.skip("oracle/_ref/libblastref.so did not travel to this box")
    cfgkw = dict(cfgkw, gap_open=2, gap_extend=2, xdrop_gap_final=1300.0)
    cfg = R.default_config(task, taps=R.TAP_LUT, **cfgkw)
    r = R.search(qs, vol, cfg)
    assert r["status"] == 0 and r["final"].shape[0] > 0
    x_final = int(r["gap_x_dropoff_final"])
    assert 2 * x_final // 2 > 1100, f"X_final {x_final}: the band would fit the shared ring"
    it = _random_start_items(r, vol, np.random.default_rng(9), per_hsp=1, max_hsps=12)
    rc = R.traceback_calls(qs, vol, it, cfg)
    assert rc["status"] == 0
    calls = rc["tb_calls"]
    h = P.batch_from_reference(r, task=task, cfg=cfg)
    V, Q = E.Volume(vol), E.Query(h)
    try:
        items = np.zeros(it.shape[0], dtype=abi.TB_ITEM_DTYPE)
        for k, col in enumerate(("oid", "context", "s_shift", "s_length", "q_start", "s_start")):
            items[col] = it[:, k]
        res, ops = E.gapped_traceback(V, Q, x_final, items)
        for k, col in ((8, "score"), (9, "query_start"), (10, "query_stop"), (11, "subject_start"), (12, "subject_stop"),
                       (14, "esp_n")):
            assert np.array_equal(res[col], calls[:, k]), col
        ref_ops = rc["tb_ops"]
        for i in range(calls.shape[0]):
            want = ref_ops[calls[i, 13]:calls[i, 13] + calls[i, 14]]
            got = ops[res["esp_off"][i]:res["esp_off"][i] + res["esp_n"][i]]
            assert np.array_equal(got["op_type"], want[:, 0]) and np.array_equal(got["num"], want[:, 1])
    finally:
        Q.free(); V.free()
