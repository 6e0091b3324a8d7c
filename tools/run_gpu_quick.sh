set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
for wl in resnet50 mobilenet_v2 vgg16 yolov8s squeezenet_v1_1; do
  timeout 300 python bench.py --workload $wl --layers --no-cpu-baseline > gpurun_out/bench_${wl}_q.json 2> gpurun_out/bench_${wl}_q.layers; tail -1 gpurun_out/bench_${wl}_q.json | cut -c1-200
done
