#!/bin/bash
TAG=r02o
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "job_pipeline or identity or batch_pipeline or traceback_search" 2>&1 | tail -15 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
