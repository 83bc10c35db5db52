"""Experiment (not the bench): config C2 through the GPU path with reference-built tables."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gblastn_b200 import synth, engine as E, abi
from oracle import refdriver as R, portdriver as P

n_q = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
db = int(sys.argv[2]) if len(sys.argv) > 2 else 250_000_000
t = time.time()
vol = synth.random_volume([db], seed=2)
qs = synth.planted_queries(vol, n_q, 1000, seed=22)
print("gen %.2fs" % (time.time() - t))
cfg = R.default_config("megablast", taps=R.TAP_LUT | R.TAP_INIT | R.TAP_GAPPED)
t = time.time(); r = R.search(qs, vol, cfg); print("ref total %.2fs prelim %.3fs" % (time.time() - t, r["seconds_prelim"]))
print("lut", r["lut_word_length"], r["scan_step"], "lookup_hits", r["lookup_hits"], "init", r["good_init_extends"], "gapext", r["gap_extensions"], "final", r["final"].shape[0])
h = P.batch_from_reference(r, task="megablast", cfg=cfg)
t = time.time(); V = E.Volume(vol); print("db_load %.3fs" % (time.time() - t))
t = time.time(); Q = E.Query(h); print("query_load %.3fs" % (time.time() - t))
for it in range(4):
    t = time.time(); g = E.prelim_search(V, Q, taps=abi.BN_TAP_INIT | abi.BN_TAP_GAPPED); dt = time.time() - t
    s = g["stats"]
    print("search %.2f ms: scan %.3f ext %.3f gapped %.3f host %.3f total %.3f launches %d" % (dt * 1e3, s["ms_scan"], s["ms_extend"], s["ms_gapped"], s["ms_host"], s["ms_total"], s["kernel_launches"]))
print("parity init", np.array_equal(P.init_table(g["init"]), r["init"]), "gapped", np.array_equal(P.gapped_table(g["gapped"]), r["gapped"]), "final", np.array_equal(P.final_table(g["hsps"]), r["final"]))
ms, bases, hits = E.bench_scan(V, Q, 20)
print("scan kernel: %.3f ms/launch, %.1f Gbases/s, %.1f GB/s, survivors %d" % (ms, bases / ms / 1e6, bases / 4 / ms / 1e6, hits))
for nt in (1, 8):
    cfg2 = R.default_config("megablast", num_threads=nt)
    r2 = R.search(qs, vol, cfg2)
    print("ref threads=%d prelim %.3fs -> %.2f Gbases/s" % (nt, r2["seconds_prelim"], db / r2["seconds_prelim"] / 1e9))
