#!/bin/bash
# dev call: CTA-pair tc_gemm vs the 1-CTA kernel, layer by layer, then the conv parity tests and whole-net lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
timeout 600 python tools/conv_layers.py --check --json gpurun_out/conv_pair1.json > gpurun_out/conv_pair1.txt 2>&1; tail -30 gpurun_out/conv_pair1.txt
NCNN_B200_TC_PAIR=0 timeout 600 python tools/conv_layers.py --check --json gpurun_out/conv_pair0.json > gpurun_out/conv_pair0.txt 2>&1; tail -30 gpurun_out/conv_pair0.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "convolution or innerproduct" --timeout 600 > gpurun_out/pytest_conv.log 2>&1; tail -5 gpurun_out/pytest_conv.log
for st in fp16 bf16; do
  timeout 300 python bench.py --storage $st --no-cpu-baseline --layers > gpurun_out/bench_resnet50_$st.json 2> gpurun_out/bench_resnet50_$st.layers; tail -c 600 gpurun_out/bench_resnet50_$st.json; echo
done
