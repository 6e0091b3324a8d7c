#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for wl in mobilenet_v2 yolov8s vgg16; do
  python bench.py --workload $wl --layers --no-extra-legs --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.layers
done
python bench.py --workload squeezenet_v1_1 --layers --no-extra-legs --no-cpu-baseline > gpurun_out/bench_squeezenet.json 2> gpurun_out/bench_squeezenet.layers
for f in default mobilenet_v2 yolov8s vgg16 squeezenet; do python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1])
print('$f', d['config']['workload'], 'value %.0f ms %.3f e2e %.0f frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']), d.get('parity',{}).get('max_norm_err'), (d['roofline'].get('depthwise') or {}).get('frac_of_hbm'))
PY
done
