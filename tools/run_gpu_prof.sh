#!/bin/bash
# One gpurun call: bench lines with per-layer tables, ncu launch list, full ncu captures reduced to CSV on the box
# (the .ncu-rep files are too big for gpurun_out's 64 MiB limit; only small ones are kept).
set -x
mkdir -p gpurun_out
for wl in resnet50 mobilenet_v2 vgg16 yolov8s; do
  extra="--no-cpu-baseline"; [ $wl = resnet50 ] && extra=""
  timeout 600 python bench.py --workload $wl --layers $extra > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.layers; tail -1 gpurun_out/bench_$wl.json
done
timeout 600 python bench.py --impl reference --steps 2 > gpurun_out/bench_reference.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_resnet50.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
prof() { # name regex skip count workload
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/prof_$1 \
      python bench.py --workload $5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1.raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details --csv > gpurun_out/prof_$1.details.csv 2>/dev/null
  sz=$(stat -c %s /tmp/prof_$1.ncu-rep); [ "$sz" -lt 12000000 ] && cp /tmp/prof_$1.ncu-rep gpurun_out/
}
prof tc_gemm tc_gemm 53 53 resnet50
prof dwconv dwconv 0 17 mobilenet_v2
du -sh gpurun_out; ls -la gpurun_out
