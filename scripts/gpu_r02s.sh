#!/bin/bash
mkdir -p gpurun_out
BN_TRACE=2 python scripts/exp_jobs.py host 12 2> gpurun_out/jobs_host.txt
tail -14 gpurun_out/jobs_host.txt
timeout 1500 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_r02s.json 2> gpurun_out/bench_err_r02s.log
python -c "
import json; d=json.load(open('gpurun_out/bench_r02s.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['ranks'][0]['e2e_ms_per_step'],'scan',d['roofline']['ms_per_launch'],d['stage_ms_per_step'], d['single_call'])"
tail -3 gpurun_out/bench_err_r02s.log
