set -x
mkdir -p gpurun_out
for t in 2 3 4; do
  timeout 300 python bench.py --workload resnet50 --no-cpu-baseline --e2e-threads $t --steps 24 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('threads',d['e2e']['host_threads'],'e2e',round(d['e2e']['value']),'serial',round(d['e2e']['serial_value']),'dev',round(d['value']))"
done
