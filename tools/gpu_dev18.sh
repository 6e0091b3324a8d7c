#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_gpu.py -x -q --timeout 600 --tb=short -k "pool or innerproduct or gemm or linear" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_nets_gpu.py -x -q --timeout 600 --tb=short -k "model_parity or full_size" 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --layers --no-extra-legs --no-cpu-baseline > gpurun_out/bench_dual.json 2> gpurun_out/bench_dual.layers
head -3 gpurun_out/bench_dual.layers; tail -5 gpurun_out/bench_dual.layers
