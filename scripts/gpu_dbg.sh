#!/bin/bash
echo "--- dense + lane-private, 6 blocks/SM"
timeout 120 python scripts/exp_scan.py 2>&1 | tail -1
