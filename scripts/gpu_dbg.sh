#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4_r04z.csv \
    python scripts/exp_c5.py c4 1.0 > gpurun_out/c4_ncu_r04z.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5_r04z.csv \
    python scripts/exp_c5.py c5 1.0 > gpurun_out/c5_ncu_r04z.log 2>&1
wc -l gpurun_out/launches_c4_r04z.csv gpurun_out/launches_c5_r04z.csv
