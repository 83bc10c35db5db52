"""Seeded parity cases shared by the oracle tests (CPU) and the GPU parity tests.

Each case = (task, reference-config overrides, volume spec, query spec).  Sizes are chosen so the
reference engine and the C port finish in well under a second; the table types, strides, diagonal
containers and gapped aligners they exercise are listed per case.
"""
from __future__ import annotations

import numpy as np

from gblastn_b200 import synth

CASES = {
    # name: dict(task, cfg, seq_lens, vol_seed, nq, qlen, q_seed, sub, indel, planted)
    # C1 of BASELINE.json: megablast 1 x 10 kb vs 1 Mb -> MB lut 11 / stride 18, hash, greedy
    "c1_megablast_10kb_vs_1mb": dict(task="megablast", cfg={}, seq_lens=[1_000_000], vol_seed=1,
                                     nq=1, qlen=10_000, q_seed=11, sub=0.02, indel=0.0, planted=1.0),
    # many short queries, indels, ragged subject lengths incl. shorter-than-word sequences
    "mb_lut11_hash_indels": dict(task="megablast", cfg={}, seq_lens=[300_000, 50_000, 777, 120_001, 13, 27, 28, 29],
                                 vol_seed=2, nq=40, qlen=500, q_seed=12, sub=0.03, indel=0.004, planted=0.8),
    # small total query -> eSmallNaLookupTable lut 8 / stride 21, diag ARRAY container
    "mb_smallna_diagarray": dict(task="megablast", cfg={}, seq_lens=[300_000, 50_000, 777, 120_001],
                                 vol_seed=2, nq=3, qlen=700, q_seed=13, sub=0.05, indel=0.01, planted=1.0),
    # blastn ws 11 -> MB lut 11 / stride 1 (no mini-extension), hash, packed DP, odd-score rounding
    "blastn_mb11_dp": dict(task="blastn", cfg={}, seq_lens=[300_000, 50_000, 777, 120_001, 13],
                           vol_seed=2, nq=30, qlen=800, q_seed=14, sub=0.08, indel=0.01, planted=0.8),
    # blastn, tiny query -> small table lut 8 / stride 4 (AlignedOneByte), diag array, DP
    "blastn_smallna_dp": dict(task="blastn", cfg={}, seq_lens=[300_000, 50_000, 777, 120_001],
                              vol_seed=2, nq=4, qlen=600, q_seed=15, sub=0.08, indel=0.01, planted=1.0),
    # word size 16 megablast: MB lut 11 / stride 6; word size 12 blastn-like
    "mb_ws16": dict(task="megablast", cfg={"word_size": 16}, seq_lens=[200_000, 90_000], vol_seed=5,
                    nq=25, qlen=600, q_seed=16, sub=0.06, indel=0.005, planted=0.8),
    # word size 27 + large query batch -> lut 12 / stride 16 (byte-aligned scan + ExtendAligned)
    "mb_ws27_lut12_aligned": dict(task="megablast", cfg={"word_size": 27}, seq_lens=[400_000, 100_000], vol_seed=6,
                                  nq=160, qlen=1000, q_seed=17, sub=0.02, indel=0.002, planted=0.5),
    # default megablast with > 300 k table entries -> lut 12 / stride 17 (config C2's table shape)
    "mb_lut12_stride17": dict(task="megablast", cfg={}, seq_lens=[500_000, 250_000], vol_seed=7,
                              nq=170, qlen=1000, q_seed=18, sub=0.02, indel=0.002, planted=0.5),
    # word size 11 megablast scoring with DP turned on / blastn with greedy
    "blastn_ws11_greedy": dict(task="blastn", cfg={"greedy": 1, "gap_open": 0, "gap_extend": 0},
                               seq_lens=[150_000, 60_000], vol_seed=8, nq=20, qlen=700, q_seed=19,
                               sub=0.05, indel=0.005, planted=0.8),
    # small word sizes on the array container: word_length < 11 exact-extension branch
    "blastn_ws7_array": dict(task="blastn", cfg={"word_size": 7}, seq_lens=[20_000, 5_000], vol_seed=9,
                             nq=2, qlen=300, q_seed=20, sub=0.10, indel=0.01, planted=1.0),
    # queries with ambiguity codes (N) : words containing them are not indexed
    "mb_with_N": dict(task="megablast", cfg={}, seq_lens=[200_000, 100_000], vol_seed=10,
                      nq=30, qlen=600, q_seed=21, sub=0.02, indel=0.002, planted=0.9, n_frac=0.004),
    # long, divergent queries: greedy distance > 254 -> exercises the tier-2 (global scratch) path on the GPU
    "mb_long_divergent_tier2": dict(task="megablast", cfg={}, seq_lens=[600_000, 200_000], vol_seed=12,
                                    nq=3, qlen=30_000, q_seed=23, sub=0.03, indel=0.002, planted=1.0),
    # nt-like volume: thousands of short sequences (log-normal lengths, many shorter than a word),
    # one launch must cover them all and hits must not leak across sequence boundaries
    "mb_ntlike_many_subjects": dict(task="megablast", cfg={}, seq_lens="lognormal:4000:7:1500:1.0", vol_seed=13,
                                    nq=60, qlen=400, q_seed=24, sub=0.02, indel=0.002, planted=0.9),
    "blastn_ntlike_many_subjects": dict(task="blastn", cfg={}, seq_lens="lognormal:1500:8:1200:0.9", vol_seed=14,
                                        nq=12, qlen=500, q_seed=25, sub=0.06, indel=0.008, planted=0.9),
    # two-hit mode (window_size > 0): s_TypeOfWord's double-word test, hit_len / hit_saved bookkeeping
    "mb_two_hit_w40_hash": dict(task="megablast", cfg={"word_size": 16, "window_size": 40}, seq_lens=[200_000, 90_000],
                                vol_seed=15, nq=25, qlen=600, q_seed=26, sub=0.06, indel=0.005, planted=0.8),
    "blastn_two_hit_w40_direct": dict(task="blastn", cfg={"window_size": 40}, seq_lens=[150_000, 60_000, 900],
                                      vol_seed=16, nq=20, qlen=700, q_seed=27, sub=0.08, indel=0.01, planted=0.8),
    "mb_two_hit_smallna_array": dict(task="megablast", cfg={"word_size": 20, "window_size": 50},
                                     seq_lens=[300_000, 50_000, 777], vol_seed=17, nq=3, qlen=700, q_seed=28,
                                     sub=0.05, indel=0.01, planted=1.0),
    "blastn_two_hit_array_ws7": dict(task="blastn", cfg={"word_size": 7, "window_size": 30}, seq_lens=[20_000, 5_000],
                                     vol_seed=18, nq=2, qlen=300, q_seed=29, sub=0.10, indel=0.01, planted=1.0),
    # two-hit mode with an off-diagonal search (-off_diagonal_range, scan_range > 0): a single word pairs with an
    # unsaved hit on a neighbouring diagonal (core/na_ungapped.c:697-726, :853-884); serial replay on the GPU
    "mb_two_hit_offdiag_hash": dict(task="megablast", cfg={"word_size": 16, "window_size": 40, "scan_range": 4},
                                    seq_lens=[200_000, 90_000], vol_seed=15, nq=25, qlen=600, q_seed=26, sub=0.06,
                                    indel=0.02, planted=0.8),
    "blastn_two_hit_offdiag_direct": dict(task="blastn", cfg={"window_size": 40, "scan_range": 6},
                                          seq_lens=[150_000, 60_000, 900], vol_seed=16, nq=20, qlen=700, q_seed=27,
                                          sub=0.08, indel=0.02, planted=0.8),
    "mb_two_hit_offdiag_smallna_array": dict(task="megablast", cfg={"word_size": 20, "window_size": 50, "scan_range": 5},
                                             seq_lens=[300_000, 50_000, 777], vol_seed=17, nq=3, qlen=700, q_seed=28,
                                             sub=0.05, indel=0.02, planted=1.0),
    "blastn_two_hit_offdiag_array_ws7": dict(task="blastn", cfg={"word_size": 7, "window_size": 30, "scan_range": 3},
                                             seq_lens=[20_000, 5_000], vol_seed=18, nq=2, qlen=300, q_seed=29, sub=0.10,
                                             indel=0.02, planted=1.0),
    "blastn_ws8_na_table_offdiag": dict(task="blastn", cfg={"word_size": 8, "window_size": 40, "scan_range": 4},
                                        seq_lens=[60_000, 20_000], vol_seed=22, nq=60, qlen=700, q_seed=32, sub=0.08,
                                        indel=0.02, planted=0.8),
    # affine greedy (BLAST_AffineGreedyAlign body): -greedy with explicit gap costs
    "mb_affine_greedy_5_2": dict(task="megablast", cfg={"greedy": 1, "gap_open": 5, "gap_extend": 2},
                                 seq_lens=[300_000, 50_000, 777, 120_001], vol_seed=2, nq=40, qlen=500, q_seed=12,
                                 sub=0.03, indel=0.004, planted=0.8),
    "blastn_affine_greedy_5_2": dict(task="blastn", cfg={"greedy": 1, "gap_open": 5, "gap_extend": 2},
                                     seq_lens=[300_000, 50_000, 777, 120_001, 13], vol_seed=2, nq=30, qlen=800,
                                     q_seed=14, sub=0.08, indel=0.01, planted=0.8),
    "mb_affine_greedy_long_tier2": dict(task="megablast", cfg={"greedy": 1, "gap_open": 3, "gap_extend": 1},
                                        seq_lens=[600_000, 200_000], vol_seed=12, nq=3, qlen=30_000, q_seed=23,
                                        sub=0.03, indel=0.002, planted=1.0),
    "mb_affine_greedy_0_2": dict(task="megablast", cfg={"word_size": 16, "greedy": 1, "gap_open": 0, "gap_extend": 2},
                                 seq_lens=[200_000, 90_000], vol_seed=5, nq=25, qlen=600, q_seed=16, sub=0.06,
                                 indel=0.005, planted=0.8),
    # BASELINE configs[2..4] at reduced size (same table shapes, containers and aligners as the full configs)
    # C3: blastn ws 11, 100 x 10 kb vs 1 Gb (10 x 100 Mb)  ->  6 x 10 kb vs 6 x 1 Mb: MB lut 11 / stride 1, DP heavy
    "c3_scaled_blastn_10kb": dict(task="blastn", cfg={}, seq_lens=[1_000_000] * 6, vol_seed=3, nq=6, qlen=10_000,
                                  q_seed=33, sub=0.08, indel=0.01, planted=0.8),
    # C4: megablast 100 k x 150 bp short reads vs 3 Gb  ->  3 000 x 150 bp vs 7 x 400 kb: lut 12 / stride 17
    "c4_scaled_short_reads": dict(task="megablast", cfg={}, seq_lens=[400_000] * 7, vol_seed=40, nq=3000, qlen=150,
                                  q_seed=44, sub=0.02, indel=0.0, planted=0.8),
    # C5: megablast 1 000 x 5 kb (+ masks) vs nt-like volume  ->  80 x 5 kb vs 3 000 log-normal sequences
    "c5_scaled_ntlike_5kb": dict(task="megablast", cfg={}, seq_lens="lognormal:3000:50:2000:1.1", vol_seed=50,
                                 nq=80, qlen=5000, q_seed=55, sub=0.02, indel=0.002, planted=0.8),
    # eNaLookupTable: word size < 9 with a query batch beyond the small table's 15-bit offsets (e.g. a batch of
    # short-word searches); lut == word -> s_BlastNaExtendDirect, stride 1
    "blastn_ws7_na_table": dict(task="blastn", cfg={"word_size": 7}, seq_lens=[60_000, 20_000, 900], vol_seed=21,
                                nq=45, qlen=800, q_seed=31, sub=0.08, indel=0.01, planted=0.8),
    "blastn_ws8_na_table_two_hit": dict(task="blastn", cfg={"word_size": 8, "window_size": 40}, seq_lens=[60_000, 20_000],
                                        vol_seed=22, nq=60, qlen=700, q_seed=32, sub=0.08, indel=0.01, planted=0.8),
    # empty result: random queries only
    "mb_no_hits": dict(task="megablast", cfg={}, seq_lens=[100_000], vol_seed=11,
                       nq=5, qlen=400, q_seed=22, sub=0.0, indel=0.0, planted=0.0),
}

# Traceback-stage list logic (containment in a better HSP's final alignment, common-endpoint trimming, second
# containment pass): queries made of several homologous segments separated by short junk / small indels, which the
# preliminary X-drop stops at and the final X-drop (100 bits) bridges, plus queries with an internal tandem copy.
CASES["mb_bridged_segments"] = dict(task="megablast", cfg={}, seq_lens=[300_000, 120_000, 40_000], vol_seed=61,
                                    nq=60, qlen=900, q_seed=62, sub=0.02, indel=0.0, planted=1.0, bridged=True)
CASES["blastn_bridged_segments"] = dict(task="blastn", cfg={}, seq_lens=[300_000, 120_000, 40_000], vol_seed=63,
                                        nq=40, qlen=900, q_seed=64, sub=0.06, indel=0.0, planted=1.0, bridged=True)
TRACEBACK_LIST_CASES = ["mb_bridged_segments", "blastn_bridged_segments"]

# Repeat-family volumes: every query hits dozens of subjects, so the per-query hit lists of the HSP stream fill up
# (prelim_hitlist_size = max(min(2 h, h + 50), 10), core/hspfilter_collector.c:335-342), turn into heaps
# (Blast_HitListUpdate, core/blast_hits.c:2924-2981) and feed hit_params->low_score back into BLAST_GetGappedScore
# (core/blast_engine.c:1313-1320, core/blast_gapalign.c:3340-3375): later subjects that only carry short fragments
# of the family are skipped without a gapped extension.
CASES["mb_repeat_family_hitlist5"] = dict(task="megablast", cfg={"hitlist_size": 5}, repeat_family=dict(
    seed=71, n_subj=70, n_fam=6, elem_len=1200, frag_from=0.45), nq=18, q_seed=72, sub=0.01)
CASES["mb_repeat_family_hitlist20"] = dict(task="megablast", cfg={"hitlist_size": 20}, repeat_family=dict(
    seed=73, n_subj=90, n_fam=4, elem_len=900, frag_from=0.6), nq=12, q_seed=74, sub=0.015)
CASES["blastn_repeat_family_hitlist5"] = dict(task="blastn", cfg={"hitlist_size": 5}, repeat_family=dict(
    seed=75, n_subj=60, n_fam=4, elem_len=700, frag_from=0.5), nq=8, q_seed=76, sub=0.04)
# hit_options->hsp_num_max is ignored by gapped searches (BlastHspNumMax, core/blast_hits.c:169-191): same lists
CASES["mb_repeat_family_hsp_num_max2"] = dict(task="megablast", cfg={"hitlist_size": 5, "hsp_num_max": 2},
                                              repeat_family=dict(seed=71, n_subj=70, n_fam=6, elem_len=1200,
                                                                 frag_from=0.45), nq=18, q_seed=72, sub=0.01)
# blastn direct mode (lut == word) with queries of very different lengths (per-context cutoffs differ) and ambiguity
# codes: the scan kernel's direct filter takes its per-context path and its keep-on-ambiguity path
CASES["blastn_direct_mixed_lengths_N"] = dict(task="blastn", cfg={}, seq_lens=[400_000, 90_000, 5_000], vol_seed=77,
                                              qlen_list=[(6, 120), (8, 700), (4, 4000), (2, 9000)], q_seed=78,
                                              sub=0.07, indel=0.008, planted=0.85, n_frac=0.003)
# hit_options->percent_identity / min_hit_length (blastn -perc_identity): Blast_HSPTest in the traceback stage
# (core/blast_traceback.c:658-669 for DP tracebacks, :727-735 after the re-evaluation); thresholds chosen inside the
# spread of the planted alignments' identities so that part of the HSPs goes and part stays
CASES["blastn_bridged_perc_identity"] = dict(task="blastn", cfg={"percent_identity": 92.0, "min_hit_length": 120},
                                             seq_lens=[300_000, 120_000, 40_000], vol_seed=63, nq=40, qlen=900,
                                             q_seed=64, sub=0.06, indel=0.0, planted=1.0, bridged=True)
CASES["mb_bridged_perc_identity"] = dict(task="megablast", cfg={"percent_identity": 96.5, "min_hit_length": 150},
                                         seq_lens=[300_000, 120_000, 40_000], vol_seed=61, nq=60, qlen=900,
                                         q_seed=62, sub=0.02, indel=0.0, planted=1.0, bridged=True)
CASES["blastn_dp_perc_identity"] = dict(task="blastn", cfg={"percent_identity": 91.0}, seq_lens=[300_000, 50_000, 777, 120_001, 13],
                                        vol_seed=2, nq=30, qlen=800, q_seed=14, sub=0.08, indel=0.01, planted=0.8)
IDENTITY_FILTER_CASES = ["blastn_bridged_perc_identity", "mb_bridged_perc_identity", "blastn_dp_perc_identity"]
HITLIST_CASES = ["mb_repeat_family_hitlist5", "mb_repeat_family_hitlist20", "blastn_repeat_family_hitlist5",
                 "mb_repeat_family_hsp_num_max2"]

FAST = ["c1_megablast_10kb_vs_1mb", "mb_lut11_hash_indels", "mb_smallna_diagarray", "blastn_mb11_dp",
        "blastn_smallna_dp", "mb_ws16", "blastn_ws11_greedy", "blastn_ws7_array", "mb_with_N", "mb_no_hits"]
ALL = list(CASES.keys())


def _seq_lens(spec, seed):
    if not isinstance(spec, str):
        return spec
    _, n, seed2, median, sigma = spec.split(":")
    rng = np.random.default_rng(int(seed2) * 7919 + seed)
    lens = np.exp(rng.normal(np.log(float(median)), float(sigma), size=int(n))).astype(np.int64)
    lens = np.clip(lens, 1, 400_000)
    lens[::97] = rng.integers(1, 30, size=lens[::97].shape[0])      # sprinkle sequences shorter than a word
    return lens


def _bridged_queries(vol, nq, qlen, seed, sub_rate):
    """Each query = 2-4 segments copied from consecutive stretches of one subject, separated by junk of 8-70 random
    bases that replaces a subject stretch of the same length +- a small indel; some queries repeat their first
    segment at the end (two alignments to the same subject region)."""
    rng = np.random.default_rng(seed)
    out = []
    lens = vol.seq_len.astype(np.int64)
    ok = np.nonzero(lens >= 4 * qlen)[0]
    for _ in range(nq):
        oid = int(ok[rng.integers(0, ok.size)])
        L = int(lens[oid])
        start = int(rng.integers(0, L - 3 * qlen))
        b0 = int(vol.byte_off[oid])
        raw = vol.packed[b0 + start // 4: b0 + (start + 3 * qlen) // 4 + 2]
        un = np.stack([raw >> 6, (raw >> 4) & 3, (raw >> 2) & 3, raw & 3], axis=1).reshape(-1)[start % 4:]
        nseg = int(rng.integers(2, 5))
        seglen = qlen // nseg - 40
        parts, pos = [], 0
        for k in range(nseg):
            seg = un[pos:pos + seglen].copy()
            m = rng.random(seglen) < sub_rate
            seg[m] = (seg[m] + rng.integers(1, 4, size=int(m.sum()))) % 4
            parts.append(seg.astype(np.uint8))
            pos += seglen
            if k + 1 < nseg:
                junk = int(rng.integers(8, 70))
                parts.append(rng.integers(0, 4, size=junk, dtype=np.uint8))
                pos += junk + int(rng.integers(-12, 13))          # the subject skips a slightly different length
        if rng.random() < 0.3:
            parts.append(parts[0].copy())
        q = np.concatenate(parts)[:qlen + 200]
        if rng.random() < 0.5:
            q = synth.revcomp(q)
        out.append(np.ascontiguousarray(q, dtype=np.uint8))
    return out


def _repeat_family(spec, nq, q_seed, sub):
    """Subjects = random background with mutated copies of a small family of elements; the first `frag_from` of the
    subjects carry whole copies at 1-9 % divergence (a later, better copy displaces the worst entry of a full hit
    list), the rest mostly short fragments (40-160 bases: their ungapped scores stay below 15 % of the worst kept
    score, so low_score skips them) with an occasional whole copy.  Queries = family members, lightly mutated."""
    rng = np.random.default_rng(spec["seed"])
    fam = [rng.integers(0, 4, size=spec["elem_len"], dtype=np.uint8) for _ in range(spec["n_fam"])]
    seqs = []
    for i in range(spec["n_subj"]):
        L = int(rng.integers(2, 5)) * spec["elem_len"] * spec["n_fam"] // 2
        s = rng.integers(0, 4, size=L, dtype=np.uint8)
        pos = int(rng.integers(0, 200))
        for f in rng.permutation(spec["n_fam"]):
            if rng.random() < 0.15:
                continue
            whole = i < spec["frag_from"] * spec["n_subj"] or rng.random() < 0.12
            e = synth.mutate(fam[f], rng, float(rng.uniform(0.01, 0.09)), 0.002)
            if not whole:
                a = int(rng.integers(0, e.shape[0] - 160))
                e = e[a:a + int(rng.integers(40, 160))]
            if rng.random() < 0.5:
                e = synth.revcomp(e)
            if pos + e.shape[0] >= L:
                break
            s[pos:pos + e.shape[0]] = e
            pos += e.shape[0] + int(rng.integers(20, 400))
        seqs.append(s)
    vol = synth.make_volume_from_bases(seqs)
    qrng = np.random.default_rng(q_seed)
    qs = []
    for k in range(nq):
        q = synth.mutate(fam[k % spec["n_fam"]], qrng, sub, 0.0)
        qs.append(np.ascontiguousarray(synth.revcomp(q) if k % 3 == 1 else q, dtype=np.uint8))
    return vol, qs


def make_case(name):
    c = CASES[name]
    if c.get("repeat_family"):
        vol, qs = _repeat_family(c["repeat_family"], c["nq"], c["q_seed"], c["sub"])
        return c["task"], dict(c["cfg"]), vol, qs
    vol = synth.random_volume(_seq_lens(c["seq_lens"], c["vol_seed"]), seed=c["vol_seed"])
    if c.get("bridged"):
        return c["task"], dict(c["cfg"]), vol, _bridged_queries(vol, c["nq"], c["qlen"], c["q_seed"], c["sub"])
    if c.get("qlen_list"):
        qs = []
        for k, (n, ql) in enumerate(c["qlen_list"]):
            qs += synth.planted_queries(vol, n, ql, seed=c["q_seed"] + k, planted_frac=c["planted"],
                                        sub_rate=c["sub"], indel_rate=c["indel"], n_frac=c.get("n_frac", 0.0))
        return c["task"], dict(c["cfg"]), vol, qs
    qs = synth.planted_queries(vol, c["nq"], c["qlen"], seed=c["q_seed"], planted_frac=c["planted"],
                               sub_rate=c["sub"], indel_rate=c["indel"], n_frac=c.get("n_frac", 0.0))
    return c["task"], dict(c["cfg"]), vol, qs
