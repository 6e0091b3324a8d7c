#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/conv_layers.py --check --json gpurun_out/conv_pair1b.json > gpurun_out/conv_pair1b.txt 2>&1; tail -27 gpurun_out/conv_pair1b.txt
NCNN_B200_TC_BN=128 timeout 600 python tools/conv_layers.py --only "s4|s5" > gpurun_out/conv_pair1_bn128.txt 2>&1; tail -14 gpurun_out/conv_pair1_bn128.txt
NCNN_B200_TC_BN=128 NCNN_B200_TC_PAIR=0 timeout 600 python tools/conv_layers.py --only "s4|s5" > gpurun_out/conv_pair0_bn128.txt 2>&1; tail -14 gpurun_out/conv_pair0_bn128.txt
