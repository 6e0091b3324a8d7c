#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_glue_gpu.py -q --timeout 1200 --tb=short -s 2>&1 | tail -60
