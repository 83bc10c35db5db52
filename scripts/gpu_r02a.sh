#!/bin/bash
# round-2 first visit: parity, bench baseline, full-size C3 / C4-shard / C5-shard stage times
TAG=r02a
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
BN_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_err_$TAG.log
cat gpurun_out/bench_$TAG.json
BN_TRACE=1 timeout 900 python scripts/exp_c3.py 100 10 100000000 > gpurun_out/exp_c3_full_$TAG.txt 2>&1
tail -5 gpurun_out/exp_c3_full_$TAG.txt
BN_TRACE=1 timeout 600 python scripts/exp_c5.py c4 1.0 > gpurun_out/exp_c4_full_$TAG.txt 2>&1
tail -6 gpurun_out/exp_c4_full_$TAG.txt
BN_TRACE=1 timeout 600 python scripts/exp_c5.py c5 1.0 > gpurun_out/exp_c5_full_$TAG.txt 2>&1
tail -6 gpurun_out/exp_c5_full_$TAG.txt
nproc; free -g | head -2; lscpu | grep -i "numa\|model name" | head
