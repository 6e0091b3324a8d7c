#!/bin/bash
# Round 2: validation of the monotonic-word two-issuer protocol of tc_gemm (all GPU tests, the 3x3 per-layer table, the default bench
# with its three-stream end-to-end leg, the other full-size workloads).
mkdir -p gpurun_out/issuer
O=gpurun_out/issuer
timeout 300 python -m pytest tests -q -m gpu --timeout 200 --tb=short -x > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 60 python tools/conv_layers.py --only "3x3" > $O/conv_layers_3x3.txt 2>&1; tail -6 $O/conv_layers_3x3.txt
timeout 200 python bench.py --layers > $O/bench_resnet50.json 2> $O/bench_resnet50.layers; tail -c 200 $O/bench_resnet50.json; tail -1 $O/bench_resnet50.layers
for wl in mobilenet_v2 yolov8s vgg16; do
  timeout 100 python bench.py --workload $wl --layers --no-cpu-baseline --no-extra-legs > $O/bench_$wl.json 2> $O/bench_$wl.layers; tail -1 $O/bench_$wl.layers
done
grep -c "timed out" $O/* | grep -v ":0$"
