#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_plugin_gpu.py tests/test_gemm_gpu.py -x -q --timeout 900 --tb=short -s 2>&1 | tail -30
