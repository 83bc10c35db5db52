#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel_filtered -s 3 -c 1 -o gpurun_out/prof_filt_r03d -f python scripts/exp_filt.py 1x10000 > gpurun_out/prof_filt_r03d.log 2>&1
tail -3 gpurun_out/prof_filt_r03d.log
