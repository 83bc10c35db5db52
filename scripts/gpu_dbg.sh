#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q -k "radix or device_lookup" 2>&1 | tail -5
