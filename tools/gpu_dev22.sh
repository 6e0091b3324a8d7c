#!/bin/bash
timeout 600 python -m pytest tests -q -m gpu --timeout 300 --tb=short -x 2>&1 | tail -4
for wl in resnet50 mobilenet_v2 yolov8s vgg16 squeezenet_v1_1; do
  timeout 90 python bench.py --workload $wl --layers --no-extra-legs --no-cpu-baseline 2>&1 >/dev/null | grep "layers total" | sed "s/^/$wl /"
done
