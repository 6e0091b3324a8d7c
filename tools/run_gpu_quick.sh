set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 600 --tb=short -k "convolution" > gpurun_out/pytest_kernels.log 2>&1; tail -5 gpurun_out/pytest_kernels.log
for wl in resnet50 mobilenet_v2; do
  timeout 300 python bench.py --workload $wl --layers --no-cpu-baseline > gpurun_out/bench_${wl}_q.json 2> gpurun_out/bench_${wl}_q.layers; tail -1 gpurun_out/bench_${wl}_q.json | cut -c1-1500
done
