#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nets_gpu.py -x -q --timeout 600 --tb=short -k "projection_shortcut" 2>&1 | tail -30
python bench.py --steps 20 --warmup 5 --layers --no-extra-legs --no-cpu-baseline > gpurun_out/bench_dual.json 2> gpurun_out/bench_dual.layers
grep -n "res2a\|res3a" gpurun_out/bench_dual.layers
