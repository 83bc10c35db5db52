"""Scan-kernel timing for small query batches against the C2 volume (250 Mb): shared-memory filter path
(scan_kernel_filtered) against the queue-driven kernel (BN_FILT_MAX=0), same batch, same results."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gblastn_b200 import engine, setup, synth
vol, _ = bench.make_workload(0, 0)
engine.init(1)
V = engine.Volume(vol)
peak, _ = bench.peak_hbm()
shapes = [(1, 10_000), (10, 1_000), (10, 10_000), (100, 1_000), (250, 1_000), (500, 1_000)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for task in ("megablast",):
    for nq, ql in shapes:
        qs = synth.planted_queries(vol, nq, ql, seed=5, planted_frac=0.8, sub_rate=0.02, rc_frac=0.5)
        s = setup.Setup(qs, task=task, db_length=vol.total_bases, db_num_seqs=vol.n_seqs, device_lookup=1)
        row = {"task": task, "queries": f"{nq}x{ql}", "lut": s.batch.lut_word_length, "step": s.batch.scan_step}
        res = {}
        for mode in ("filtered", "queue"):
            if mode == "queue":
                os.environ["BN_FILT_MAX"] = "0"
            else:
                os.environ.pop("BN_FILT_MAX", None)
            Q = engine.Query(s.batch)
            engine.bench_scan(V, Q, 3)
            ms, bases, hits = engine.bench_scan(V, Q, 20)
            g = engine.prelim_search(V, Q)
            res[mode] = g
            row[mode + "_us"] = round(1e3 * ms, 2)
            row[mode + "_frac_hbm"] = round(0.25 * bases / (ms * 1e-3) / 1e9 / peak, 4)
            row["survivors"] = int(hits)
            row["lookup_hits"] = int(g["stats"]["lookup_hits"])
            Q.free()
        row["identical"] = res["filtered"]["hsps"].tobytes() == res["queue"]["hsps"].tobytes() and \
            res["filtered"]["stats"]["lookup_hits"] == res["queue"]["stats"]["lookup_hits"]
        print(json.dumps(row), flush=True)
        s.free()
