"""One C2 search through the blocking call (for ncu captures of every kernel of a job)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gblastn_b200 import engine, setup
vol, qs = bench.make_workload(0, 0)
engine.init(1)
s = setup.Setup(qs, task="megablast", db_length=vol.total_bases, db_num_seqs=vol.n_seqs, device_lookup=1)
V = engine.Volume(vol); Q = engine.Query(s.batch)
for _ in range(3):
    g = engine.prelim_search(V, Q)
print(g["stats"])
