#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final_check.json 2>/dev/null
python - <<P
import json
d=json.loads(open("gpurun_out/bench_final_check.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), "traffic", d["roofline"]["traffic"], "parity", d.get("parity_vs_reference"), "launches", d["gpu_launches"], d["clocks"])
for c in d.get("configs", []): print(c["name"], round(c.get("ms_per_pass",0),2), c.get("parity_vs_reference",{}).get("identical"), c.get("error"))
P
