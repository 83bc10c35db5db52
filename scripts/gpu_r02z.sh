#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "dust" 2>&1 | tail -3
python scripts/exp_dust.py
