#!/bin/bash
timeout 900 python -m pytest tests/test_nets_gpu.py -q --timeout 600 --tb=short -k "nanodet or fastest" 2>&1 | tail -30
