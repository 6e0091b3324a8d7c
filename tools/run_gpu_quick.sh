set -x
mkdir -p gpurun_out
for a in 0 1 2 3 4 7; do
  NCNN_B200_EPI_ABLATE=$a timeout 200 python bench.py --workload resnet50 --layers --no-cpu-baseline --e2e-threads 1 --steps 10 > gpurun_out/abl_$a.json 2> gpurun_out/abl_$a.layers
  echo "ablate $a: $(python -c "import json;print(round(json.loads(open('gpurun_out/abl_$a.json').read().strip().splitlines()[-1])['value']))") $(grep -E '^conv1 |res2a_branch1 |res2a_branch2c |res2b_branch2a |res4b_branch2c ' gpurun_out/abl_$a.layers | awk '{printf "%s=%s ", $1, $3}')"
done
