#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu --set full of the scan kernel.
# usage: scripts/gpu_round.sh TAG [quick]
TAG=${1:-r01x}
MODE=${2:-full}
mkdir -p gpurun_out
if [ "$MODE" = "quick" ]; then
  python -m pytest tests -m gpu -x -q -k "c1_ or lut12 or ntlike or masked or host_buffer" 2>&1 | tail -15 > gpurun_out/tests_$TAG.log
else
  python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests_$TAG.log
fi
cat gpurun_out/tests_$TAG.log
BN_TRACE=1 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_err_$TAG.log
cat gpurun_out/bench_$TAG.json
tail -3 gpurun_out/bench_err_$TAG.log
if [ "$3" = "noncu" ]; then exit 0; fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_kernel_staged -s 4 -c 2 -o gpurun_out/prof_scan_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2_$TAG.log 2>&1
ls -la gpurun_out | tail -8
