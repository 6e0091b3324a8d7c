#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nets_gpu.py tests/test_kernels_gpu.py -x -q --timeout 600 --tb=short -k "stem" 2>&1 | tail -3
for d in 0 7 23; do
  NCNN_B200_STEM_DBG=$d python bench.py --steps 10 --warmup 3 --layers --no-extra-legs --no-cpu-baseline 2>&1 >/dev/null | grep "^conv1" | sed "s/^/dbg=$d /"
done
