#!/bin/bash
# multi-GPU bench line (one rank per GPU under torchrun), N given as $1
N=${1:-2}
mkdir -p gpurun_out
nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r04z_n$N.json 2> gpurun_out/bench_r04z_n$N.err
python - <<P
import json
for l in open("gpurun_out/bench_r04z_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N value",round(d["value"]),"ms",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"])); print([ (r["ms_per_step"], r["e2e_ms_per_step"], r["ms_host"]) for r in d["ranks"]])
        for c in d.get("configs", []): print(c["name"], c.get("ms_per_pass"), c.get("gbases_per_s"), c.get("parity_vs_reference",{}).get("identical"), c.get("error"))
P
tail -2 gpurun_out/bench_r04z_n$N.err
