# compute-sanitizer over the kernels added in round 2 (stem fold, projection-shortcut fold, whole-map pooling, narrow fc tiles);
# memcheck, then racecheck and synccheck on the mbarrier / named-barrier kernels (logs under gpurun_out/)
set -x
mkdir -p gpurun_out
SEL="stem_conv_maxpool or projection_shortcut"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_nets_gpu.py -q --timeout 500 -p no:cacheprovider -k "$SEL and (fp16 or bf16)" > gpurun_out/sanitizer_r2_memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/sanitizer_r2_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_nets_gpu.py -q --timeout 800 -p no:cacheprovider -k "(stem_conv_maxpool and case0 and fp16) or (projection_shortcut and case1 and fp16)" > gpurun_out/sanitizer_r2_racecheck.log 2>&1
echo "rc=$?" >> gpurun_out/sanitizer_r2_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_nets_gpu.py tests/test_kernels_gpu.py -q --timeout 500 -p no:cacheprovider -k "(stem_conv_maxpool and fp16) or (projection_shortcut and fp16) or pooling or depthwise" > gpurun_out/sanitizer_r2_synccheck.log 2>&1
echo "rc=$?" >> gpurun_out/sanitizer_r2_synccheck.log
tail -4 gpurun_out/sanitizer_r2_memcheck.log gpurun_out/sanitizer_r2_racecheck.log gpurun_out/sanitizer_r2_synccheck.log
