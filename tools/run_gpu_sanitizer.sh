# compute-sanitizer memcheck over the kernel-level GPU tests (small shapes; a few minutes)
set -x
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -q --timeout 480 -p no:cacheprovider > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitizer.log | head -20
