set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 600 --tb=short > gpurun_out/pytest_kernels.log 2>&1; tail -15 gpurun_out/pytest_kernels.log
timeout 600 python -m pytest tests/test_nets_gpu.py -q --timeout 500 --tb=short > gpurun_out/pytest_nets.log 2>&1; tail -8 gpurun_out/pytest_nets.log
for wl in resnet50 mobilenet_v2 vgg16 yolov8s; do
 for st in bf16; do
  timeout 300 python bench.py --workload $wl --storage $st --layers --no-cpu-baseline > gpurun_out/bench_${wl}_$st.json 2> gpurun_out/bench_${wl}_$st.layers; tail -1 gpurun_out/bench_${wl}_$st.json | cut -c1-120
 done
done
NCNN_B200_STEM_UNFOLD=0 timeout 300 python bench.py --workload resnet50 --storage bf16 --layers --no-cpu-baseline > gpurun_out/bench_resnet50_nounfold.json 2> gpurun_out/bench_resnet50_nounfold.layers; head -1 gpurun_out/bench_resnet50_nounfold.layers
NCNN_B200_STEM_UNFOLD=0 timeout 300 python bench.py --workload vgg16 --storage bf16 --layers --no-cpu-baseline > gpurun_out/bench_vgg16_nounfold.json 2> gpurun_out/bench_vgg16_nounfold.layers; head -1 gpurun_out/bench_vgg16_nounfold.layers
