#!/bin/bash
# ncu --set full captures of every kernel of the hot path (one launch each), C2 bench workload + C3-shaped run.
# usage: scripts/gpu_profile.sh TAG
TAG=${1:-r01x}
mkdir -p gpurun_out
# C2: skip the launches of the warm-up steps (3 x 7) and the set-up kernels; take the 7 kernels of one timed step
ncu --set full --clock-control none --import-source on -k regex:'bn::' --launch-skip 40 --launch-count 7 \
    -o gpurun_out/prof_c2_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_c2_$TAG.log 2>&1
# query load (device-side table fill) kernels
ncu --set full --clock-control none -k regex:'mb_|build_|mark_' --launch-count 12 \
    -o gpurun_out/prof_load_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_load_$TAG.log 2>&1
# C3-shaped blastn run: extend kernels at 10 M seed hits, thread DP, warp DP
ncu --set full --clock-control none --import-source on -k regex:'gapped_kernel|gapped_warp_kernel|extend_kernel|extend_leaders_kernel|scan_kernel' \
    --launch-skip 10 --launch-count 5 -o gpurun_out/prof_c3_$TAG -f python scripts/exp_c3.py 20 4 25000000 > gpurun_out/prof_c3_$TAG.log 2>&1
./scripts/gather_probe > gpurun_out/gather_probe_$TAG.txt 2>&1
ls -la gpurun_out | tail -8
