#!/bin/bash
mkdir -p gpurun_out
BN_TRACE=2 python scripts/exp_jobs.py resident 12 2> gpurun_out/jobs_resident.txt
tail -14 gpurun_out/jobs_resident.txt
BN_TRACE=2 python scripts/exp_jobs.py host 12 2> gpurun_out/jobs_host.txt
tail -14 gpurun_out/jobs_host.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
