#!/bin/bash
mkdir -p gpurun_out
for d in 3 4 5; do
  BN_JOB_DEPTH=$d python bench.py --steps 60 --warmup 20 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('depth $d value', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['ranks'][0])"
done
BN_TRACE=2 BN_JOB_DEPTH=4 python scripts/exp_jobs.py resident 16 2>&1 | tail -18 | cut -c1-200
