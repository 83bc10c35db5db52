#!/bin/bash
TAG=r03j
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_err_$TAG.log
python - <<P
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "parity", d.get("parity_vs_reference"))
for c in d.get("configs", []): print(c["name"], c.get("ms_per_pass"), c.get("gbases_per_s"), c.get("stage_ms"), c.get("parity_vs_reference",{}).get("identical"), c.get("error"))
P
tail -3 gpurun_out/bench_err_$TAG.log
