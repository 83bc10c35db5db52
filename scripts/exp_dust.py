import time, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gblastn_b200 import engine as E, synth
E.init(1)
vol = synth.random_volume([3_000_000], seed=50)
qs = synth.planted_queries(vol, 1000, 5000, seed=55, planted_frac=0.8, sub_rate=0.02, indel_rate=0.002)
qs = synth.add_low_complexity(qs, seed=56, frac=0.3)
E.dust_mask_batch(qs[:4])
for _ in range(2):
    t0=time.perf_counter(); m=E.dust_mask_batch(qs); t1=time.perf_counter()
    print("device batch ms", 1e3*(t1-t0))
