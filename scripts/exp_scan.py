"""Scan-kernel timing on the C2 workload for the library named by GBLASTN_B200_LIB (kernel experiments)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gblastn_b200 import engine, setup
vol, qs = bench.make_workload(0, 0)
engine.init(1)
s = setup.Setup(qs, task="megablast", db_length=vol.total_bases, db_num_seqs=vol.n_seqs, device_lookup=1)
V = engine.Volume(vol); Q = engine.Query(s.batch)
engine.bench_scan(V, Q, 3)
ms, bases, hits = engine.bench_scan(V, Q, 20)
g = engine.prelim_search(V, Q)
print(f"{os.environ.get('GBLASTN_B200_LIB', 'default'):40s} scan {1e3 * ms:7.2f} us  survivors {hits}  hsps {g['hsps'].size} lookup_hits {g['stats']['lookup_hits']}")
