#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
BN_TRACE=2 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_r02x_$i.json 2> gpurun_out/bench_r02x_$i.err
python - <<P
import json
for l in open("gpurun_out/bench_r02x_$i.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=1 value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"], d["ranks"][0]["e2e_ms_per_step"], d["stage_ms_per_step"])
P
done
