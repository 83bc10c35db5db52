#!/bin/bash
TAG=r02j
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gapped_long_kernel -c 1 -o gpurun_out/prof_long_$TAG -f python scripts/exp_c3.py 100 10 100000000 > gpurun_out/prof_long_$TAG.log 2>&1
tail -3 gpurun_out/prof_long_$TAG.log
ls -la gpurun_out/prof_long_$TAG.ncu-rep
