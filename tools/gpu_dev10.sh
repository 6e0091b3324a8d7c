#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dw_layers.py --check | tee gpurun_out/dw_cold.txt
timeout 300 python tools/dw_layers.py --warm | tee gpurun_out/dw_warm.txt
for L in block15 block8; do
timeout 600 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:dwconv -s 4 -c 1 -f -o /tmp/p_$L \
    python tools/dw_layers.py --only "$L\$" --iters 3 > gpurun_out/ncu_dw_$L.log 2>&1
ncu -i /tmp/p_$L.ncu-rep --page raw --csv > gpurun_out/dw_$L.raw.csv 2>/dev/null
ncu -i /tmp/p_$L.ncu-rep --page source --csv --print-source sass > gpurun_out/dw_$L.source.csv 2>/dev/null
done
ls -la gpurun_out/dw_*
