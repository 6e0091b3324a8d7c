set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py -q --timeout 600 --tb=short -k "depthwise" > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python -m pytest tests/test_nets_gpu.py -q --timeout 500 --tb=short -k "mobilenet or mnasnet or efficientnet or shufflenet" > gpurun_out/pytest_nets.log 2>&1; tail -4 gpurun_out/pytest_nets.log
timeout 300 python bench.py --workload mobilenet_v2 --layers --no-cpu-baseline > gpurun_out/bench_mobilenet_v2_q.json 2> gpurun_out/bench_mobilenet_v2_q.layers; tail -1 gpurun_out/bench_mobilenet_v2_q.json | cut -c1-150; grep dwise gpurun_out/bench_mobilenet_v2_q.layers
