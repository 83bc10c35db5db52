#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./scripts/gather_probe 2>&1 | grep -E "TEX|LDG.64 4MB per=8 blocks/SM=4|LDG.32 2MB per=8 blocks/SM=4" | tee gpurun_out/gather_probe_r03n.txt
