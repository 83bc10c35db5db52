"""Experiment (not the bench): the traceback-stage alignment kernels (bn_gapped_traceback) on the calls the
reference's own traceback stage makes for a C2-shaped megablast batch (greedy traceback) and a C3-shaped blastn
batch (ALIGN_EX), timed next to the reference's whole traceback stage on one host core.  Needs oracle/_ref."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gblastn_b200 import synth, engine as E, abi
from oracle import refdriver as R, portdriver as P


def run(tag, task, vol, qs, **cfgkw):
    cfg = R.default_config(task, taps=R.TAP_LUT | R.TAP_TRACEBACK, prelim_only=0, **cfgkw)
    r = R.search(qs, vol, cfg)
    assert r["status"] == 0
    calls = r["tb_calls"]
    h = P.batch_from_reference(r, task=task, cfg=cfg)
    V, Q = E.Volume(vol), E.Query(h)
    items = np.zeros(calls.shape[0], dtype=abi.TB_ITEM_DTYPE)
    items["oid"], items["context"], items["s_shift"] = calls[:, 1], calls[:, 2], calls[:, 3]
    items["q_start"], items["s_start"], items["s_length"] = calls[:, 4], calls[:, 5], calls[:, 7]
    times = []
    for it in range(5):
        t = time.perf_counter()
        res, ops = E.gapped_traceback(V, Q, int(r["gap_x_dropoff_final"]), items)
        times.append(time.perf_counter() - t)
    ok = all(np.array_equal(res[c], calls[:, k]) for k, c in ((8, "score"), (9, "query_start"), (10, "query_stop"),
                                                                (11, "subject_start"), (12, "subject_stop"), (14, "esp_n")))
    ref_ops = r["tb_ops"]
    for i in range(calls.shape[0]):
        w = ref_ops[calls[i, 13]:calls[i, 13] + calls[i, 14]]
        g = ops[res["esp_off"][i]:res["esp_off"][i] + res["esp_n"][i]]
        ok = ok and np.array_equal(g["op_type"], w[:, 0]) and np.array_equal(g["num"], w[:, 1])
    aligned = int((calls[:, 10] - calls[:, 9]).sum())
    out = {"case": tag, "alignments": int(calls.shape[0]), "kind": "greedy" if calls[0, 0] == 1 else "dp",
           "aligned_query_bases": aligned, "edit_ops": int(ops.shape[0]),
           "gpu_ms_best": round(min(times) * 1e3, 3), "gpu_ms_all": [round(t * 1e3, 3) for t in times],
           "reference_traceback_stage_ms_1core": round(r["seconds_traceback"] * 1e3, 2),
           "reference_prelim_stage_ms_1core": round(r["seconds_prelim"] * 1e3, 2),
           "identical_to_reference": bool(ok)}
    print(json.dumps(out), flush=True)
    Q.free(); V.free()


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
