#!/bin/bash
for cfg in "20 3" "60 3" "60 2" "60 4" "60 6"; do
  set -- $cfg
  python bench.py --steps $1 --warmup 3 --e2e-threads $2 --no-extra-legs --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('steps $1 threads $2: value %.0f  e2e %.0f  serial %.0f  h2d %.1f GB/s  pixels %.0f' % (d['value'], e['value'], e['serial_value'], e['h2d_gbs_per_gpu'], e['pixels_value'] or 0))"
done
