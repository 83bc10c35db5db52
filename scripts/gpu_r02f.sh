#!/bin/bash
TAG=r02f
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
BN_TRACE=1 timeout 900 python scripts/exp_c3.py 100 10 100000000 > gpurun_out/exp_c3_full_$TAG.txt 2>&1
tail -4 gpurun_out/exp_c3_full_$TAG.txt
BN_TRACE=1 timeout 600 python scripts/exp_c5.py c4 1.0 > gpurun_out/exp_c4_full_$TAG.txt 2>&1
tail -3 gpurun_out/exp_c4_full_$TAG.txt
