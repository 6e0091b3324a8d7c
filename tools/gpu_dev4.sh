#!/bin/bash
mkdir -p gpurun_out
prof() { # name only-regex pair
  NCNN_B200_TC_PAIR=$3 timeout 600 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:tc_gemm -s 4 -c 1 -f -o /tmp/p_$1 \
    python tools/conv_layers.py --only "$2" --iters 3 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/p_$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i /tmp/p_$1.ncu-rep --page source --csv --print-source sass > gpurun_out/$1.source.csv 2>/dev/null
}
prof s2_3x3 "s2 3x3" 0
prof s3_3x3_p0 "s3 3x3" 0
prof s3_3x3_p1 "s3 3x3" 1
prof stem "conv1" 0
prof s4_res_p1 "s4 1x1 256" 1
ls -la gpurun_out | tail -12
