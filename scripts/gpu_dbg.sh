#!/bin/bash
echo "--- dense, 7 blocks/SM (32 regs)"; timeout 120 python scripts/exp_scan.py 2>&1 | tail -1
timeout 300 python scripts/exp_c5.py c5 1.0 2>&1 | grep -E "^search" | tail -1 | cut -c1-120
