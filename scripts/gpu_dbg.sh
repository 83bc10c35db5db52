#!/bin/bash
timeout 120 python scripts/exp_c3.py 100 2 50000000 2>&1 | tail -1
BN_NO_SCALAR_LEADERS=1 timeout 120 python scripts/exp_c3.py 100 2 50000000 2>&1 | tail -1
timeout 400 python -m pytest tests -m gpu -x -q -k "direct or c3 or blastn" 2>&1 | tail -4
