set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 600 --tb=short -k "convolution" > gpurun_out/pytest_kernels.log 2>&1; tail -15 gpurun_out/pytest_kernels.log
NCNN_B200_CONV_SHIFT=2 timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 600 --tb=line -k "shifted_window" > gpurun_out/pytest_shift2.log 2>&1; tail -5 gpurun_out/pytest_shift2.log
for wl in resnet50 vgg16; do
  timeout 300 python bench.py --workload $wl --storage bf16 --layers --no-cpu-baseline > gpurun_out/bench_${wl}_bf16.json 2> gpurun_out/bench_${wl}_bf16.layers; tail -1 gpurun_out/bench_${wl}_bf16.json | cut -c1-120; head -8 gpurun_out/bench_${wl}_bf16.layers
done
bash tools/ncu_hot.sh vggstem tc_gemm 48 1 vgg16 --storage bf16
