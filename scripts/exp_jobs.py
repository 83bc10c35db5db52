"""Timeline of the job pipeline on the C2 workload (BN_TRACE=2 prints per-job host timestamps)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from gblastn_b200 import engine, setup

n_sets = 4
sets = [bench.make_workload(0, k) for k in range(n_sets)]
for v, _ in sets:
    p = torch.empty(v.packed.shape[0], dtype=torch.uint8).pin_memory(); p.numpy()[:] = v.packed; v.packed = p.numpy()
engine.init(1)
sts = [setup.Setup(q, task="megablast", db_length=v.total_bases, db_num_seqs=v.n_seqs, device_lookup=1) for v, q in sets]
Vs = [engine.Volume(v) for v, _ in sets]
Qs = [engine.Query(s.batch) for s in sts]
mode = sys.argv[1] if len(sys.argv) > 1 else "resident"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
if mode == "resident":
    jobs = [{"volume": Vs[i % n_sets], "query": Qs[i % n_sets]} for i in range(n)]
else:
    jobs = [{"host_volume": sets[i % n_sets][0], "batch": sts[i % n_sets].batch} for i in range(n)]
tr = os.environ.pop("BN_TRACE", None)
engine.prelim_search_jobs(jobs[:4])
if tr:
    os.environ["BN_TRACE"] = tr
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
for rep in range(reps):
    if tr and rep < reps - 1:
        os.environ.pop("BN_TRACE", None)
    elif tr:
        os.environ["BN_TRACE"] = tr
    t0 = time.perf_counter()
    engine.prelim_search_jobs(jobs)
    print(f"{mode}: {1e3 * (time.perf_counter() - t0) / n:.3f} ms per job", file=sys.stderr)
