"""Experiment (not the bench): blastn-mode (C3-shaped) stage times on the GPU path, product set-up."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gblastn_b200 import synth, engine as E, setup as S

n_q = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_seq = int(sys.argv[2]) if len(sys.argv) > 2 else 4
seq_len = int(sys.argv[3]) if len(sys.argv) > 3 else 25_000_000
planted = float(sys.argv[4]) if len(sys.argv) > 4 else 0.8
vol = synth.random_volume([seq_len] * n_seq, seed=3)
qs = synth.planted_queries(vol, n_q, 10_000, seed=33, planted_frac=planted, sub_rate=0.08, indel_rate=0.01)
s = S.Setup(qs, task="blastn", db_length=vol.total_bases, db_num_seqs=vol.n_seqs)
V = E.Volume(vol); Q = E.Query(s.batch)
for it in range(3):
    t = time.time(); g = E.prelim_search(V, Q); dt = time.time() - t
    st = g["stats"]
    print("search %.1f ms: scan %.2f ext %.2f gapped %.2f host %.2f | lookup_hits %d init %d gapext %d hsps %d launches %d -> %.2f Gbases/s" % (
        dt * 1e3, st["ms_scan"], st["ms_extend"], st["ms_gapped"], st["ms_host"], st["lookup_hits"], st["good_init_extends"],
        st["gap_extensions"], g["hsps"].size, st["kernel_launches"], vol.total_bases / dt / 1e9), flush=True)
