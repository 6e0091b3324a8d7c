#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "depthwise" --timeout 600 2>&1 | tail -4
timeout 300 python tools/dw_layers.py --check | tee gpurun_out/dw_cold2.txt
timeout 300 python tools/dw_layers.py --warm | tail -2
