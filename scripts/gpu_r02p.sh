#!/bin/bash
TAG=r02p
mkdir -p gpurun_out
BN_TRACE=0 timeout 1500 python bench.py --steps 20 --warmup 3 --no-configs > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_err_$TAG.log
tail -c 4000 gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_err_$TAG.log
