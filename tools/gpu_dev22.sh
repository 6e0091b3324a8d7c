#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_nets_gpu.py -x -q --timeout 600 --tb=short -k "convolution or model_parity or stem" 2>&1 | tail -3
python bench.py --workload mobilenet_v2 --layers --no-extra-legs --no-cpu-baseline 2>&1 >/dev/null | grep "^conv1\|block1/linear\|block3/linear\|layers total"
python bench.py --workload yolov8s --layers --no-extra-legs --no-cpu-baseline 2>&1 >/dev/null | grep "^conv_1 \|^conv_2 \|layers total"
python bench.py --layers --no-extra-legs --no-cpu-baseline 2>&1 >/dev/null | grep "layers total"
