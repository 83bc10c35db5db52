#!/bin/bash
TAG=r02l
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_err_$TAG.log
python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'scan',d['roofline']['ms_per_launch'],'frac',d['roofline']['frac'],d['stage_ms_per_step'])"
BN_NO_SIG=1 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_nosig_$TAG.json 2> /dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_nosig_$TAG.json')); print('NOSIG value',d['value'],'ms',d['ms_per_step'],'scan',d['roofline']['ms_per_launch'],d['stage_ms_per_step'])"
