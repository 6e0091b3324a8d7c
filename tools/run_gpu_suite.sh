#!/bin/bash
# One gpurun call: parity tests, smoke, bench lines, per-launch ncu list, one full ncu capture of the top kernel.
# Outputs under gpurun_out/.  Usage: tools/run_gpu_suite.sh [quick]
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
if [ "$1" != "quick" ]; then
  timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --tb=short > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
fi
timeout 600 python bench.py --layers > gpurun_out/bench_resnet50.json 2> gpurun_out/bench_resnet50.layers; tail -1 gpurun_out/bench_resnet50.json; tail -3 gpurun_out/bench_resnet50.layers
timeout 600 python bench.py --workload mobilenet_v2 --layers --no-cpu-baseline > gpurun_out/bench_mobilenet_v2.json 2> gpurun_out/bench_mobilenet_v2.layers; tail -1 gpurun_out/bench_mobilenet_v2.json
if [ "$1" != "quick" ]; then
  timeout 600 python bench.py --workload vgg16 --layers --no-cpu-baseline > gpurun_out/bench_vgg16.json 2> gpurun_out/bench_vgg16.layers; tail -1 gpurun_out/bench_vgg16.json
  timeout 600 python bench.py --workload yolov8s --layers --no-cpu-baseline > gpurun_out/bench_yolov8s.json 2> gpurun_out/bench_yolov8s.layers; tail -1 gpurun_out/bench_yolov8s.json
  timeout 600 python bench.py --impl reference --steps 2 > gpurun_out/bench_reference.json 2>&1; tail -1 gpurun_out/bench_reference.json
fi
# ncu: launch list of the bench command (shares, not absolutes), then one full capture of the conv kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_resnet50.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 60 -c 4 -f -o gpurun_out/prof_tc_gemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dwconv -s 10 -c 4 -f -o gpurun_out/prof_dwconv \
    python bench.py --workload mobilenet_v2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_dw.log 2>&1
ls -la gpurun_out
