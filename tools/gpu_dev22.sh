#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu --timeout 300 --tb=short -x 2>&1 | tail -3
timeout 300 python bench.py --layers > gpurun_out/bench_resnet50.json 2> gpurun_out/bench_resnet50.layers; tail -c 300 gpurun_out/bench_resnet50.json; tail -2 gpurun_out/bench_resnet50.layers
for wl in mobilenet_v2 yolov8s vgg16 squeezenet_v1_1; do
  timeout 120 python bench.py --workload $wl --layers --no-cpu-baseline --no-extra-legs > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.layers; tail -1 gpurun_out/bench_$wl.layers
done
