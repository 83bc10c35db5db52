#!/bin/bash
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 BN_TRACE=2 python scripts/exp_jobs.py host 20 4 2> gpurun_out/jobs_host_g0.txt &
CUDA_VISIBLE_DEVICES=1 BN_TRACE=2 python scripts/exp_jobs.py host 20 4 2> gpurun_out/jobs_host_g1.txt &
wait
grep "per job" gpurun_out/jobs_host_g0.txt gpurun_out/jobs_host_g1.txt
grep "slow step" gpurun_out/jobs_host_g0.txt gpurun_out/jobs_host_g1.txt | tail -30
