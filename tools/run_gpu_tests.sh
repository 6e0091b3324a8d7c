# the whole GPU test tier, nothing else (log under gpurun_out/)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 --tb=short > gpurun_out/pytest_gpu.log 2>&1; tail -40 gpurun_out/pytest_gpu.log
