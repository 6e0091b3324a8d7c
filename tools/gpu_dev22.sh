#!/bin/bash
timeout 600 python -m pytest tests/test_nets_gpu.py -x -q --timeout 600 --tb=short -k "projection_shortcut" 2>&1 | tail -3
python bench.py --workload yolov8s --layers --no-extra-legs --no-cpu-baseline 2>&1 >/dev/null | grep "conv_1 \|conv_2 \|head0_flat\|up2\|layers total"
python bench.py --layers --no-extra-legs --no-cpu-baseline 2>&1 >/dev/null | grep "layers total"
python bench.py --workload mobilenet_v2 --layers --no-extra-legs --no-cpu-baseline 2>&1 >/dev/null | grep "layers total"
