#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q -k "band_wider" 2>&1 | tail -12
