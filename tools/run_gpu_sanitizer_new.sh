# compute-sanitizer memcheck over the tests of the operators added late in the round (log under gpurun_out/)
set -x
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py tests/test_nets_gpu.py -q --timeout 180 -p no:cacheprovider \
    -k "deconvolution or reduction or layernorm or gelu or yolov8_decode or multihead or pixels_resize or yolov8_device" > gpurun_out/sanitizer_new.log 2>&1
echo "rc=$?" >> gpurun_out/sanitizer_new.log
tail -8 gpurun_out/sanitizer_new.log
