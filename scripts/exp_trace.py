"""Experiment: host-phase breakdown (BN_TRACE) of config C2 with the bench's own set-up path."""
import os, sys
os.environ["BN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gblastn_b200 import engine, setup
vol, qs = bench.make_workload(0)
engine.init(0, [0])
s = setup.Setup(qs, task="megablast", db_length=vol.total_bases, db_num_seqs=vol.n_seqs, device_lookup=1)
V = engine.Volume(vol, device=0)
Q = engine.Query(s.batch)
for _ in range(8):
    g = engine.prelim_search(V, Q)
print(g["stats"])
