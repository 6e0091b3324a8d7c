set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_nets_gpu.py -q --timeout 600 --tb=line -k "reference_benchmark" > gpurun_out/pytest_extra.log 2>&1; tail -40 gpurun_out/pytest_extra.log
