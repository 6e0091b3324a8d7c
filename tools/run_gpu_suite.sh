#!/bin/bash
# One gpurun call: parity tests, smoke, bench lines, per-launch ncu list.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 --tb=short > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
python bench.py --layers > gpurun_out/bench_resnet50.json 2> gpurun_out/bench_resnet50.layers; tail -1 gpurun_out/bench_resnet50.json; tail -3 gpurun_out/bench_resnet50.layers
python bench.py --workload mobilenet_v2 --layers --no-cpu-baseline > gpurun_out/bench_mobilenet_v2.json 2> gpurun_out/bench_mobilenet_v2.layers; tail -1 gpurun_out/bench_mobilenet_v2.json
