#!/bin/bash
# Round 2, closing capture (short: what fits in the GPU minutes left).  Parity tests, smoke, the default bench line + per-layer
# tables of the five workloads, the reference arm, and the ncu launch list of the default bench command.  Everything lands in
# gpurun_out/ and is copied to profiles/r2/ afterwards.  The ncu --set full captures of the round are in profiles/r2/ already
# (tools/run_gpu_round_r2.sh).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
timeout 420 python -m pytest tests -q -m gpu --timeout 300 --tb=short --deselect tests/test_graph_gpu.py > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -m pytest tests/test_graph_gpu.py -q -m gpu --timeout 150 --tb=short > gpurun_out/pytest_graph_gpu.log 2>&1; tail -15 gpurun_out/pytest_graph_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 300 python bench.py --layers > gpurun_out/bench_resnet50.json 2> gpurun_out/bench_resnet50.layers; tail -c 400 gpurun_out/bench_resnet50.json; tail -2 gpurun_out/bench_resnet50.layers
timeout 200 python bench.py --workload squeezenet_v1_1 --layers --no-cpu-baseline > gpurun_out/bench_squeezenet_v1_1.json 2> gpurun_out/bench_squeezenet_v1_1.layers; tail -c 900 gpurun_out/bench_squeezenet_v1_1.json
for wl in mobilenet_v2 vgg16 yolov8s; do
  timeout 120 python bench.py --workload $wl --layers --no-cpu-baseline --no-extra-legs > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.layers; tail -1 gpurun_out/bench_$wl.layers
done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>&1; tail -c 300 gpurun_out/bench_reference.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_resnet50.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_launches.log 2>&1
grep -c "mbarrier wait timed out" gpurun_out/*.json gpurun_out/*.log gpurun_out/*.layers | grep -v ":0$"
ls gpurun_out | wc -l
