"""Experiment (not the bench): a stream of DIFFERENT C2-shaped query batches against one resident 250 Mb volume,
one bn_query_load + bn_prelim_search per batch versus bn_prelim_search_batches (two-stage pipeline)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gblastn_b200 import synth, engine as E, setup as S

n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 8
vol = synth.random_volume([250_000_000], seed=2)
setups = []
for k in range(n_batches):
    qs = synth.planted_queries(vol, 1000, 1000, seed=100 + k, planted_frac=0.8, sub_rate=0.02)
    setups.append(S.Setup(qs, task="megablast", db_length=vol.total_bases, db_num_seqs=vol.n_seqs, device_lookup=1))
V = E.Volume(vol)
for rep in range(3):
    t = time.time()
    singles = []
    for st in setups:
        Q = E.Query(st.batch); singles.append(E.prelim_search(V, Q)); Q.free()
    t_seq = time.time() - t
    t = time.time()
    piped = E.prelim_search_batches(V, [st.batch for st in setups])
    t_pipe = time.time() - t
    same = all(a["hsps"].tobytes() == b["hsps"].tobytes() for a, b in zip(singles, piped))
    print("batches %d: sequential %.2f ms/batch (%.0f Gbases/s), pipelined %.2f ms/batch (%.0f Gbases/s), identical %s" % (
        n_batches, 1e3 * t_seq / n_batches, vol.total_bases * n_batches / t_seq / 1e9,
        1e3 * t_pipe / n_batches, vol.total_bases * n_batches / t_pipe / 1e9, same), flush=True)
