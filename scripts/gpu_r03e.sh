#!/bin/bash
# N=8: the driver's scaling command (full default line incl. the C5 block) and the reference arm
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r03e_n8.json 2> gpurun_out/bench_r03e_n8.err
python - <<'P'
import json
for l in open("gpurun_out/bench_r03e_n8.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=8 value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"]); [print(r) for r in d["ranks"]]
        for c in d.get("configs", []): print(c["name"], c.get("ms_per_pass"), c.get("gbases_per_s"), c.get("parity_vs_reference"))
P
tail -3 gpurun_out/bench_r03e_n8.err
