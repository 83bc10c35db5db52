#!/bin/bash
TAG=r02e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "direct_filter" 2>&1 | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3_$TAG.csv python scripts/exp_c3.py 100 10 100000000 > gpurun_out/exp_c3_ncu_$TAG.txt 2>&1
tail -3 gpurun_out/exp_c3_ncu_$TAG.txt
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_c3_r02e.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); 
agg=collections.OrderedDict()
for r in rows[hdr+1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg.setdefault(r[ki][:70],[]).append(v)
for k,v in agg.items(): print(f"{k:70s} n={len(v):3d} total={sum(v)/1e6:9.3f} ms max={max(v)/1e6:9.3f} ms")
PY
