#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu | grep -iE "numa|socket|model name|^cpu\(s\)" >> gpurun_out/topo.txt; cat gpurun_out/topo.txt | tail -20
timeout 900 python bench.py --layers > gpurun_out/bench_default.json 2> gpurun_out/bench_default.layers; tail -c 3000 gpurun_out/bench_default.json; echo; tail -3 gpurun_out/bench_default.layers
timeout 1500 python -m pytest tests/test_nets_gpu.py -q -k "full_size" --timeout 900 --tb=short -s 2>&1 | grep -E "full size|passed|failed|Error|error" | tail -20
