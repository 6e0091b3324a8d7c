#!/bin/bash
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_gpu.py tests/test_nets_gpu.py -q --timeout 300 --tb=short -x -k "convolution or gemm or innerproduct or model_parity or shortcut" 2>&1 | tail -3
for wl in resnet50 mobilenet_v2 yolov8s vgg16; do
  timeout 90 python bench.py --workload $wl --layers --no-extra-legs --no-cpu-baseline 2>&1 >/dev/null | grep "layers total" | sed "s/^/$wl /"
done
