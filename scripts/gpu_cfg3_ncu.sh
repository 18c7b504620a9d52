#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_multicrop_step.py -q -x --timeout 300 2>&1 | tail -3
python scripts/cfg3_step.py
