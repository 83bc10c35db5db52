#!/bin/bash
# round-2 record: parity tests, full bench line (C2 + C3 block), launch list, --set full of the scan kernel
TAG=r02w
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_err_$TAG.log
tail -c 3000 gpurun_out/bench_$TAG.json
tail -3 gpurun_out/bench_err_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/b_ncu_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel_staged -s 4 -c 2 -o gpurun_out/prof_scan_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/b_ncu2_$TAG.log 2>&1
ls -la gpurun_out | tail -8
