#!/bin/bash
# compute-sanitizer memcheck over the whole GPU suite (0 errors expected)
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests -m gpu -q > gpurun_out/memcheck_all.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/memcheck_all.log | head -10
