#!/bin/bash
TAG=r02m
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
BN_TRACE=1 timeout 600 python scripts/exp_c5.py c4 1.0 > gpurun_out/exp_c4_full_$TAG.txt 2>&1
tail -4 gpurun_out/exp_c4_full_$TAG.txt
