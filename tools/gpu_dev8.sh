#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "convolution or innerproduct" --timeout 600 2>&1 | tail -8
timeout 600 python tools/conv_layers.py --check --json gpurun_out/conv_r2c.json > gpurun_out/conv_r2c.txt 2>&1; tail -27 gpurun_out/conv_r2c.txt
NCNN_B200_TC_TMASTORE=0 timeout 600 python tools/conv_layers.py > gpurun_out/conv_r2c_direct.txt 2>&1; tail -3 gpurun_out/conv_r2c_direct.txt
timeout 600 python tools/conv_layers.py --workload mobilenet_v2 --check > gpurun_out/conv_mbv2_tma.txt 2>&1; tail -25 gpurun_out/conv_mbv2_tma.txt
NCNN_B200_TC_TMASTORE=0 timeout 600 python tools/conv_layers.py --workload mobilenet_v2 > gpurun_out/conv_mbv2_direct.txt 2>&1; tail -2 gpurun_out/conv_mbv2_direct.txt
