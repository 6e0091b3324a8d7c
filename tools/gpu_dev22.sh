#!/bin/bash
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_kernels_gpu.py -x -q --timeout 600 --tb=short -k "model_parity or squeezenet_golden or mat_batch or stem_conv or preprocessing" 2>&1 | tail -3
python tools/e2e_probe.py
