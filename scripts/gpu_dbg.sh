#!/bin/bash
for v in mb4 mb5 mb6; do
GBLASTN_B200_LIB=$PWD/gblastn_b200/libvar_$v.so timeout 200 python bench.py --steps 60 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v value', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'scan', round(d['roofline']['ms_per_launch'],4))"
done
