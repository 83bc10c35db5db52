#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r04g.json 2> gpurun_out/bench_err_r04g.log
python - <<P
import json
d=json.loads(open("gpurun_out/bench_r04g.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["roofline"]["ms_per_launch"], "parity", d.get("parity_vs_reference"), "launches", d["gpu_launches"])
print("small", d["small_batch"]["filtered"]["scan_ms"], d["small_batch"]["filtered"]["frac_of_hbm_peak"], d["small_batch"]["identical"])
for c in d.get("configs", []): print(c["name"], c.get("ms_per_pass"), c.get("gbases_per_s"), c.get("stage_ms"), c.get("parity_vs_reference",{}).get("identical"), c.get("error"))
P
tail -3 gpurun_out/bench_err_r04g.log
