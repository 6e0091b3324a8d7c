#!/bin/bash
# One gpurun call per round: parity tests, smoke, the bench lines (per-layer tables on stderr), the reference arm,
# the ncu launch list of the bench command and full ncu captures reduced to CSV on the box
# (.ncu-rep files are usually too big for gpurun_out's 64 MiB limit; only small ones are kept).
# Usage: tools/run_gpu_round.sh [notests]
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
if [ "$1" != "notests" ]; then
  timeout 1200 python -m pytest tests -q -m gpu --timeout 900 --tb=short > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
fi
for wl in resnet50 mobilenet_v2 vgg16 yolov8s; do
  extra="--no-cpu-baseline"; [ $wl = resnet50 ] && extra=""
  timeout 600 python bench.py --workload $wl --layers $extra > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.layers; tail -1 gpurun_out/bench_$wl.json
done
timeout 600 python bench.py --impl reference --steps 2 > gpurun_out/bench_reference.json 2>&1; tail -1 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_resnet50.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
prof() { # name regex skip count workload
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/prof_$1 \
      python bench.py --workload $5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1.raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details --csv > gpurun_out/prof_$1.details.csv 2>/dev/null
  sz=$(stat -c %s /tmp/prof_$1.ncu-rep); [ "$sz" -lt 12000000 ] && cp /tmp/prof_$1.ncu-rep gpurun_out/
}
prof tc_gemm tc_gemm 159 53 resnet50
prof dwconv dwconv 51 17 mobilenet_v2
bash tools/ncu_hot.sh rn_stage2 tc_gemm 212 6 resnet50
du -sh gpurun_out; ls -la gpurun_out
