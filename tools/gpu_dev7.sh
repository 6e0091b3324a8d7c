#!/bin/bash
mkdir -p gpurun_out
L="s2 1x1 64->256 @56|s2 1x1 64->64|conv1"
for pr in 0 1; do echo "== pair $pr"; NCNN_B200_TC_PAIR=$pr timeout 300 python tools/conv_layers.py --only "$L" 2>&1 | grep -E "^s[0-9]|^conv"; done
echo "== BN128 pair0"; NCNN_B200_TC_BN=128 NCNN_B200_TC_PAIR=0 timeout 300 python tools/conv_layers.py --only "$L" 2>&1 | grep -E "^s[0-9]|^conv"
echo "== BN128 pair1"; NCNN_B200_TC_BN=128 NCNN_B200_TC_PAIR=1 timeout 300 python tools/conv_layers.py --only "$L" 2>&1 | grep -E "^s[0-9]|^conv"
echo "== BN64 pair0"; NCNN_B200_TC_BN=64 NCNN_B200_TC_PAIR=0 timeout 300 python tools/conv_layers.py --only "$L" 2>&1 | grep -E "^s[0-9]|^conv"
prof() { # name only-regex pair
  NCNN_B200_TC_PAIR=$3 timeout 600 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:tc_gemm -s 4 -c 1 -f -o /tmp/p_$1 \
    python tools/conv_layers.py --only "$2" --iters 3 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/p_$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i /tmp/p_$1.ncu-rep --page source --csv --print-source sass > gpurun_out/$1.source.csv 2>/dev/null
}
prof k64_p0 "s2 1x1 64->256 @56" 0
ls -la gpurun_out | tail -4
