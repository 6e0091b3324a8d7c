#!/bin/bash
mkdir -p gpurun_out
L="s4 3x3|s3 3x3|s4 1x1 256"
for st in 2 3 4 6; do for pr in 0 1; do
  echo "== stages $st pair $pr"; NCNN_B200_TC_STAGES=$st NCNN_B200_TC_PAIR=$pr timeout 300 python tools/conv_layers.py --only "$L" 2>&1 | grep -E "^s[0-9]"
done; done
for g in 74 148; do for pr in 0 1; do
  echo "== grid $g pair $pr"; NCNN_B200_TC_GRID=$g NCNN_B200_TC_PAIR=$pr timeout 300 python tools/conv_layers.py --only "$L" 2>&1 | grep -E "^s[0-9]"
done; done
for pr in 0 1; do
NCNN_B200_TC_PAIR=$pr timeout 600 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:tc_gemm -s 4 -c 1 -f -o /tmp/s4_pair$pr \
    python tools/conv_layers.py --only "s4 3x3" --iters 3 > gpurun_out/ncu_s4_pair$pr.log 2>&1
ncu -i /tmp/s4_pair$pr.ncu-rep --page raw --csv > gpurun_out/s4_pair$pr.raw.csv 2>/dev/null
ncu -i /tmp/s4_pair$pr.ncu-rep --page source --csv --print-source sass > gpurun_out/s4_pair$pr.source.csv 2>/dev/null
done
ls -la gpurun_out | tail -8
