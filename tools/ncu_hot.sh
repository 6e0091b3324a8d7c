#!/bin/bash
# ncu --set full with source sampling for a few launches of one kernel family; dumps raw + per-SASS-line sample counts.
# Usage: tools/ncu_hot.sh <name> <kernel-regex> <launch-skip> <launch-count> <workload> [extra bench args]
name=$1; regex=$2; skip=$3; count=$4; wl=$5; shift 5
mkdir -p gpurun_out
timeout 900 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:$regex -s $skip -c $count -f -o /tmp/hot_$name \
    python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/hot_$name.log 2>&1
ncu -i /tmp/hot_$name.ncu-rep --page raw --csv > gpurun_out/hot_$name.raw.csv 2>/dev/null
ncu -i /tmp/hot_$name.ncu-rep --page source --csv --print-source sass > gpurun_out/hot_$name.source.csv 2>gpurun_out/hot_$name.source.err
ls -la gpurun_out/hot_$name.*
