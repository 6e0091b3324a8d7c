#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_nets_gpu.py -x -q --timeout 1200 --tb=short -s -k "concat or fusion or folded or model_parity" 2>&1 | tail -40
timeout 600 python bench.py --workload yolov8s --layers --no-cpu-baseline --no-extra-legs > gpurun_out/bench_yolov8s.json 2> gpurun_out/bench_yolov8s.layers; tail -c 900 gpurun_out/bench_yolov8s.json; echo; tail -2 gpurun_out/bench_yolov8s.layers
awk '{t[$2]+=$3; n[$2]++} END {for (k in t) printf "%-24s %3d %8.4f ms\n", k, n[k], t[k]}' gpurun_out/bench_yolov8s.layers | sort -k3 -n -r | head -12
