#!/bin/bash
echo "--- dense, mask-first compaction"; timeout 120 python scripts/exp_scan.py 2>&1 | tail -1
timeout 300 python scripts/exp_c5.py c5 1.0 2>&1 | grep -E "^search" | tail -1 | cut -c1-120
timeout 300 python scripts/exp_c5.py c4 1.0 2>&1 | grep -E "^search" | tail -1 | cut -c1-120
timeout 600 python -m pytest tests -m gpu -x -q -k "dense_signature or filtered_scan or full_size or masks" 2>&1 | tail -2
