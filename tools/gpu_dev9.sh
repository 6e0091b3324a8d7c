#!/bin/bash
mkdir -p gpurun_out
python tools/hbm_probe.py | tee gpurun_out/hbm_probe.json
timeout 600 python tools/conv_layers.py --json gpurun_out/conv_r2d.json > gpurun_out/conv_r2d.txt 2>&1; tail -3 gpurun_out/conv_r2d.txt
timeout 600 python bench.py --workload mobilenet_v2 --layers --no-cpu-baseline --no-extra-legs > gpurun_out/bench_mbv2.json 2> gpurun_out/bench_mbv2.layers; grep -E "DepthWise|layers total" gpurun_out/bench_mbv2.layers
