#!/bin/bash
# scripts/build_variant.sh NAME "-DBN_...=..."  ->  gblastn_b200/libvar_NAME.so (scan_kernel.cu recompiled with the flags)
set -e
NAME=$1; FLAGS=$2
CSRC=gblastn_b200/csrc; OBJ=/tmp/bnobj; mkdir -p $OBJ
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
for f in radix_sort.cu dust_kernel.cu extend_kernel.cu gapped_kernel.cu traceback_kernel.cu lookup_build.cu group_sort.cu triage_kernel.cu engine.cu hostpost.cpp setup.cpp dbfile.cpp dust.cpp; do
  o=$OBJ/${f%.*}.o
  if [ ! -f $o ] || [ $CSRC/$f -nt $o ] || [ $CSRC/bn_device.cuh -nt $o ]; then $NV -c $CSRC/$f -o $o & fi
done
wait
$NV $FLAGS -c $CSRC/scan_kernel.cu -o $OBJ/scan_$NAME.o
$NV -shared -o gblastn_b200/libvar_$NAME.so $OBJ/scan_$NAME.o $OBJ/extend_kernel.o $OBJ/gapped_kernel.o $OBJ/traceback_kernel.o $OBJ/lookup_build.o $OBJ/group_sort.o $OBJ/triage_kernel.o $OBJ/engine.o $OBJ/hostpost.o $OBJ/setup.o $OBJ/dbfile.o $OBJ/dust.o $OBJ/dust_kernel.o $OBJ/radix_sort.o
echo built gblastn_b200/libvar_$NAME.so
