#!/bin/bash
TAG=r03b
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
timeout 600 python scripts/exp_filt.py > gpurun_out/filt_$TAG.txt 2>&1
cat gpurun_out/filt_$TAG.txt
python scripts/exp_scan.py 2>&1 | tail -1
BN_NO_CONSEC=1 python scripts/exp_scan.py 2>&1 | tail -1
