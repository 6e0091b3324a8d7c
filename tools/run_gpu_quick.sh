set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 --tb=short -k "batchnorm or unfused or reshape" > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
