#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-configs > gpurun_out/bench_r02v_n2.json 2> gpurun_out/bench_r02v_n2.err
python - <<'P'
import json
for l in open("gpurun_out/bench_r02v_n2.json"):
    if l.startswith("{"):
        d=json.loads(l); print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"]); [print(r) for r in d["ranks"]]
P
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_r02v.json 2> gpurun_out/bench_r02v.err
python - <<'P'
import json
for l in open("gpurun_out/bench_r02v.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=1 value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"]); [print(r) for r in d["ranks"]]
P
