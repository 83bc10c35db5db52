#!/bin/bash
mkdir -p gpurun_out
( nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|core|thread|model name"; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done; nproc ) > gpurun_out/topo.txt 2>&1
BN_TRACE=2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-configs > gpurun_out/bench_r02t_n2.json 2> gpurun_out/bench_r02t_n2.err
python - <<'P'
import json
for l in open("gpurun_out/bench_r02t_n2.json"):
    if l.startswith("{"):
        d=json.loads(l); print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"]); [print(r) for r in d["ranks"]]
P
