"""Experiment (not the bench): C4 / C5-shaped megablast runs at a fraction of the full size:
stage times of the GPU path with product set-up (device-side table fill)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gblastn_b200 import synth, engine as E, setup as S

mode = sys.argv[1] if len(sys.argv) > 1 else "c5"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
rng = np.random.default_rng(50)
t0 = time.time()
if mode == "c5":       # nt-like volume: log-normal lengths (median 2 kb), 1000 x 5 kb queries
    total = int(2_500_000_000 * scale)
    lens = np.clip(np.exp(rng.normal(np.log(2000.0), 1.2, size=int(total / 4000))).astype(np.int64), 30, 10_000_000)
    vol = synth.random_volume(lens, seed=50)
    qs = synth.planted_queries(vol, 1000, 5000, seed=55, planted_frac=0.8, sub_rate=0.02, indel_rate=0.002)
else:                  # c4: short reads vs long sequences
    vol = synth.random_volume([int(107_000_000 * scale)] * 7, seed=40)
    qs = synth.planted_queries(vol, int(100_000 * min(1.0, scale * 3)), 150, seed=44, planted_frac=0.8, sub_rate=0.02)
print("generated: %d seqs, %.1f Mb, %d queries in %.1fs" % (vol.n_seqs, vol.total_bases / 1e6, len(qs), time.time() - t0), flush=True)
t = time.time(); s = S.Setup(qs, task="megablast", db_length=vol.total_bases, db_num_seqs=vol.n_seqs, device_lookup=1)
print("setup %.3fs lut %d step %d concat %d" % (time.time() - t, s.batch.lut_word_length, s.batch.scan_step, s.batch.concat_len), flush=True)
t = time.time(); V = E.Volume(vol); print("db_load %.3fs" % (time.time() - t), flush=True)
t = time.time(); Q = E.Query(s.batch); print("query_load %.3fs" % (time.time() - t), flush=True)
for it in range(4):
    t = time.time(); g = E.prelim_search(V, Q); dt = time.time() - t
    st = g["stats"]
    print("search %.2f ms: scan %.3f ext %.3f gapped %.3f host %.3f | lookup_hits %d init %d hsps %d launches %d -> %.1f Gbases/s" % (
        dt * 1e3, st["ms_scan"], st["ms_extend"], st["ms_gapped"], st["ms_host"], st["lookup_hits"], st["good_init_extends"],
        g["hsps"].size, st["kernel_launches"], vol.total_bases / dt / 1e9), flush=True)
ms, bases, hits = E.bench_scan(V, Q, 10)
print("scan kernel alone: %.3f ms -> %.1f Gbases/s, %.1f GB/s algorithmic" % (ms, bases / ms / 1e6, bases / 4 / ms / 1e6))
