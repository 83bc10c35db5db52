"""Experiment (not the bench): whole search = preliminary stage + traceback stage on the GPU path for a C2 batch
(megablast 1000x1 kb vs 250 Mb) and a C3-shaped batch (blastn 20x10 kb vs 100 Mb), next to the reference's two
stages on one host core; checks the final results against the reference's.  Needs oracle/_ref."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gblastn_b200 import synth, engine as E, setup as S
from oracle import refdriver as R


def run(tag, task, vol, qs):
    r = R.search(qs, vol, R.default_config(task, taps=R.TAP_TRACEBACK, prelim_only=0))
    want = r["tb_final"]
    s = S.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs, device_lookup=1)
    V, Q = E.Volume(vol), E.Query(s.batch)
    xf = s.gap_x_dropoff_final()
    tp, tt = [], []
    for it in range(5):
        t0 = time.perf_counter(); g = E.prelim_search(V, Q); t1 = time.perf_counter()
        got, ops = E.traceback_search(V, Q, xf, g["hsps"]); t2 = time.perf_counter()
        tp.append(t1 - t0); tt.append(t2 - t1)
    ok = got.shape[0] == want.shape[0]
    if ok:
        for k, col in enumerate(("query_index", "oid", "context", "q_off", "q_end", "s_off", "s_end", "score", "num_ident")):
            ok = ok and np.array_equal(got[col], want[:, k])
        ev = want[:, 9].astype(np.uint32).astype(np.uint64) | (want[:, 10].astype(np.uint32).astype(np.uint64) << np.uint64(32))
        ok = ok and np.array_equal(got["evalue"].view(np.uint64), ev)
    print(json.dumps({"case": tag, "prelim_hsps": int(g["hsps"].shape[0]), "final_hsps": int(got.shape[0]),
                      "edit_ops": int(ops.shape[0]),
                      "gpu_prelim_ms": round(min(tp) * 1e3, 3), "gpu_traceback_stage_ms": round(min(tt) * 1e3, 3),
                      "reference_prelim_ms_1core": round(r["seconds_prelim"] * 1e3, 1),
                      "reference_traceback_ms_1core": round(r["seconds_traceback"] * 1e3, 1),
                      "identical_to_reference": bool(ok)}), flush=True)
    Q.free(); V.free(); s.free()


E.init(1)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "c2"):
    vol = synth.random_volume([250_000_000], seed=2)
    qs = synth.planted_queries(vol, 1000, 1000, seed=22, planted_frac=0.8, sub_rate=0.02, rc_frac=0.5)
    run("C2 megablast 1000x1kb vs 250Mb", "megablast", vol, qs)
if which in ("all", "c3"):
    vol = synth.random_volume([25_000_000] * 4, seed=3)
    qs = synth.planted_queries(vol, 20, 10_000, seed=33, planted_frac=0.8, sub_rate=0.08, indel_rate=0.01)
    run("C3-shaped blastn 20x10kb vs 100Mb", "blastn", vol, qs)
if which in ("c4",):
    vol = synth.random_volume([50_000_000] * 4, seed=40)
    qs = synth.planted_queries(vol, 20_000, 150, seed=44, planted_frac=0.8, sub_rate=0.02, indel_rate=0.0)
    run("C4-shaped megablast 20000x150bp vs 200Mb", "megablast", vol, qs)
if which in ("c5",):
    vol = synth.random_volume(np.clip(np.exp(np.random.default_rng(5).normal(np.log(2000), 1.1, size=60_000)), 50, 400_000).astype(np.int64), seed=50)
    qs = synth.planted_queries(vol, 500, 5000, seed=55, planted_frac=0.8, sub_rate=0.02, indel_rate=0.002)
    run("C5-shaped megablast 500x5kb vs nt-like 60k sequences", "megablast", vol, qs)
