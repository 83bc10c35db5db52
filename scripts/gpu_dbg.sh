#!/bin/bash
mkdir -p gpurun_out
BN_TRACE=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_trace_n8.json 2> gpurun_out/bench_trace_n8.err
grep -c "job " gpurun_out/bench_trace_n8.err
BN_TRACE=2 CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_trace_n1.json 2> gpurun_out/bench_trace_n1.err
grep -c "job " gpurun_out/bench_trace_n1.err
python - <<P
import json
for f in ("gpurun_out/bench_trace_n8.json","gpurun_out/bench_trace_n1.json"):
  for l in open(f):
    if l.startswith("{"):
        d=json.loads(l); print(f, "value",round(d["value"]),"ms",round(d["ms_per_step"],4)); print([ r["ms_per_step"] for r in d["ranks"]])
P
