#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_nets_gpu.py tests/test_kernels_gpu.py -x -q --timeout 1200 --tb=short -k "out_of_memory or mapped or yolov8_decode or concat_in_place or folded" 2>&1 | tail -30
