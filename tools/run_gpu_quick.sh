set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_nets_gpu.py -q --timeout 600 --tb=short -k "preprocessing" > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_resnet50_q.json 2> gpurun_out/bench_q.err; tail -3 gpurun_out/bench_q.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_resnet50_q.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'])"
