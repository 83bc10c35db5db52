#!/bin/bash
TAG=r02g
mkdir -p gpurun_out
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_err_$TAG.log
cat gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_err_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1
cat gpurun_out/bench_ref_$TAG.json
