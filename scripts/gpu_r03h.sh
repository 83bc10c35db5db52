#!/bin/bash
mkdir -p gpurun_out
nproc; lscpu | grep -E "^CPU\(s\)|Thread|Socket|NUMA node\(s\)"
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-configs > gpurun_out/bench_r03h_$1.json 2> gpurun_out/bench_r03h_$1.err
  python - <<P
import json
for l in open("gpurun_out/bench_r03h_$1.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$1 N=8 value",round(d["value"]),"ms",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"])); print([ (r["ms_per_step"], r["e2e_ms_per_step"], r["ms_host"]) for r in d["ranks"]])
P
}
run bind
BN_BENCH_NO_BIND=1 run nobind
