#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/exp_c3.py 100 2 50000000 2>&1 | tail -1
