#!/bin/bash
timeout 600 python -m pytest tests/test_nets_gpu.py -x -q --timeout 600 --tb=short -k "stem_conv" 2>&1 | tail -3
for i in 1 2; do python bench.py --layers --no-extra-legs --no-cpu-baseline 2>&1 >/dev/null | grep "^conv1\|layers total"; done
