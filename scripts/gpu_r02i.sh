#!/bin/bash
TAG=r02i
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
BN_TRACE=1 timeout 900 python scripts/exp_c3.py 100 10 100000000 > gpurun_out/exp_c3_full_$TAG.txt 2>&1
tail -2 gpurun_out/exp_c3_full_$TAG.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3_$TAG.csv python scripts/exp_c3.py 100 10 100000000 > gpurun_out/exp_c3_ncu_$TAG.txt 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_c3_r02i.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[hdr+1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg.setdefault(r[ki][:60],[]).append(v)
for k,v in agg.items():
    if sum(v)>50000: print(f"{k:60s} n={len(v):3d} total={sum(v)/1e6:9.3f} ms max={max(v)/1e6:9.3f} ms")
PY
