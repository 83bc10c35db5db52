#!/bin/bash
mkdir -p gpurun_out
for v in "$@"; do
  GBLASTN_B200_LIB=$PWD/gblastn_b200/libvar_$v.so python scripts/exp_scan.py 2>&1 | tail -1
done | tee gpurun_out/scanvar.txt
