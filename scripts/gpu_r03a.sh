#!/bin/bash
# re-entry check: parity tests + full bench line
TAG=r03a
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_err_$TAG.log
tail -c 1500 gpurun_out/bench_$TAG.json
tail -3 gpurun_out/bench_err_$TAG.log
