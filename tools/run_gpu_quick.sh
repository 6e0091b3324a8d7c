set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 300 --tb=short -k "convolution or innerproduct" > gpurun_out/pytest_kernels.log 2>&1; tail -12 gpurun_out/pytest_kernels.log
timeout 600 python -m pytest tests/test_nets_gpu.py -q --timeout 500 --tb=short > gpurun_out/pytest_nets.log 2>&1; tail -4 gpurun_out/pytest_nets.log
for wl in resnet50 mobilenet_v2 vgg16 yolov8s; do
  timeout 300 python bench.py --workload $wl --layers --no-cpu-baseline > gpurun_out/bench_${wl}_q.json 2> gpurun_out/bench_${wl}_q.layers; tail -1 gpurun_out/bench_${wl}_q.json | cut -c1-200
done
