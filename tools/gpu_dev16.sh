#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nets_gpu.py -x -q --timeout 600 --tb=short -k "stem_conv_maxpool or projection_shortcut or test_model_parity" 2>&1 | tail -30
python bench.py --steps 20 --warmup 5 --layers --no-extra-legs --no-cpu-baseline > gpurun_out/bench_dual.json 2> gpurun_out/bench_dual.layers
head -4 gpurun_out/bench_dual.layers; tail -2 gpurun_out/bench_dual.layers
