#!/bin/bash
# round-2 final record (r04): parity tests, full bench line, launch lists (C2 job pipeline, C3-shaped run), --set full of the scan kernel
TAG=${1:-r04z}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_err_$TAG.log
python - <<P
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "parity", d.get("parity_vs_reference"), "launches", d["gpu_launches"])
print("small", d["small_batch"]["filtered"]["scan_ms"], d["small_batch"]["filtered"]["frac_of_hbm_peak"], d["small_batch"]["identical"])
for c in d.get("configs", []): print(c["name"], c.get("ms_per_pass"), c.get("gbases_per_s"), c.get("stage_ms"), c.get("parity_vs_reference",{}).get("identical"), c.get("error"))
P
tail -3 gpurun_out/bench_err_$TAG.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2>> gpurun_out/bench_err_$TAG.log
tail -c 600 gpurun_out/bench_reference_$TAG.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/b_ncu_$TAG.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_c3_$TAG.csv \
    python scripts/exp_c3.py 100 2 50000000 > gpurun_out/c3_ncu_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_kernel_staged -s 4 -c 2 -o gpurun_out/prof_scan_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/b_ncu2_$TAG.log 2>&1
ls -la gpurun_out | tail -6
