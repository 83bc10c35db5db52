#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "test_filtered_scan or (database_masks and mb_lut11)" --tb=short 2>&1 | grep -v "^$" | tail -80
