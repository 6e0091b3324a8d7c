#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/conv_layers.py --check --json gpurun_out/conv_r2b.json > gpurun_out/conv_r2b.txt 2>&1; tail -27 gpurun_out/conv_r2b.txt
NCNN_B200_TC_PAIR=0 timeout 600 python tools/conv_layers.py --only "s3|s4|s5" > gpurun_out/conv_r2b_p0.txt 2>&1; tail -18 gpurun_out/conv_r2b_p0.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "convolution or innerproduct" --timeout 600 2>&1 | tail -3
