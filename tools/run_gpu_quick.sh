# quick GPU check between full rounds: the whole GPU test tier + one bench line
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 --tb=line > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_resnet50_q.json 2> gpurun_out/bench_q.err; tail -1 gpurun_out/bench_resnet50_q.json | cut -c1-300
