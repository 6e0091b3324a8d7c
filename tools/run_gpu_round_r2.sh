#!/bin/bash
# Round 2: one gpurun call that leaves everything the judge reads under gpurun_out/ (copied to profiles/r2/ afterwards):
# parity tests, smoke, bench lines + per-layer tables of the five workloads, the reference arm, the ncu launch list of the
# default bench command, ncu --set full captures of the dominant kernel families reduced to CSV, per-layer conv / dw tables.
# Usage: tools/run_gpu_round_r2.sh [notests]
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
if [ "$1" != "notests" ]; then
  timeout 1200 python -m pytest tests -q -m gpu --timeout 900 --tb=short > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
fi
timeout 900 python bench.py --layers > gpurun_out/bench_resnet50.json 2> gpurun_out/bench_resnet50.layers; tail -c 600 gpurun_out/bench_resnet50.json
for wl in mobilenet_v2 vgg16 yolov8s squeezenet_v1_1; do
  timeout 600 python bench.py --workload $wl --layers --no-cpu-baseline --no-extra-legs > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.layers; tail -c 300 gpurun_out/bench_$wl.json
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>&1; tail -1 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_resnet50.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_launches.log 2>&1
prof() { # name regex skip count workload
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/prof_$1 \
      python bench.py --workload $5 --steps 1 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_full_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1.raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/prof_$1.raw.csv > gpurun_out/ncu_$1_summary.txt
  python tools/ncu_traffic.py gpurun_out/prof_$1.raw.csv > gpurun_out/traffic_$1.json
}
# one step of ResNet-50 = 48 tc_gemm convolutions + fc + the fused stem (rows_pack + stem_pool): 51 launches of these families
prof resnet50_conv "tc_gemm|stem_pool|rows_pack" 153 51 resnet50
prof mobilenet_v2_dw dwconv 51 17 mobilenet_v2
timeout 600 python tools/conv_layers.py --json gpurun_out/conv_layers_resnet50.json > gpurun_out/conv_layers_resnet50.txt 2>&1
timeout 600 python tools/dw_layers.py > gpurun_out/dw_layers_cold.txt 2>&1
du -sh gpurun_out; ls gpurun_out | wc -l
