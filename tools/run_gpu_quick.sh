set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_nets_gpu.py -q --timeout 600 --tb=line -k "nanodet or fastestv2" > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
