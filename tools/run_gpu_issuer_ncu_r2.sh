#!/bin/bash
# Round 2: the two-issuer protocol under ncu kernel replay (serialised, cold-cache passes: TMA loads land late and out of order, the
# regime in which the earlier parity-based variants deadlocked), then the default bench line of the final binary.
mkdir -p gpurun_out/issuer
O=gpurun_out/issuer
prof() { # name skip count workload
  timeout 110 ncu --set full --clock-control none -k regex:tc_gemm -s $2 -c $3 -f -o /tmp/prof_$1 \
      python bench.py --workload $4 --steps 1 --warmup 3 --no-cpu-baseline --no-extra-legs > $O/ncu_replay_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > /tmp/prof_$1.raw.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/prof_$1.raw.csv > $O/ncu_replay_$1_summary.txt
  tail -2 $O/ncu_replay_$1_summary.txt; grep -c "timed out" $O/ncu_replay_$1.log
}
prof mobilenet_v2 35 16 mobilenet_v2
prof resnet50 49 14 resnet50
timeout 150 python bench.py --layers > $O/bench_resnet50_final.json 2> $O/bench_resnet50_final.layers; tail -c 200 $O/bench_resnet50_final.json; tail -1 $O/bench_resnet50_final.layers
grep -c "timed out" $O/* | grep -v ":0$"
