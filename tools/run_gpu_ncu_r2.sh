#!/bin/bash
# Round 2, closing counters: ncu --set full over the convolution-family launches of one ResNet-50 step and the depthwise launches of one
# MobileNetV2 step (reduced to CSV summaries + DRAM traffic), the per-layer conv / depthwise tables, the CUDA-graph tests, and the default
# bench line once more with roofline.traffic read from the fresh capture.
mkdir -p gpurun_out
prof() { # name regex skip count workload timeout
  timeout $6 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/prof_$1 \
      python bench.py --workload $5 --steps 1 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_full_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1.raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/prof_$1.raw.csv > gpurun_out/ncu_$1_summary.txt
  python tools/ncu_traffic.py gpurun_out/prof_$1.raw.csv > gpurun_out/traffic_$1.json
  tail -2 gpurun_out/ncu_$1_summary.txt
}
# one step of ResNet-50 = 48 tc_gemm convolutions + fc + the fused stem (rows_pack + stem_pool): 51 launches of these families
prof resnet50_conv "tc_gemm|stem_pool|rows_pack" 153 51 resnet50 420
prof mobilenet_v2_dw dwconv 51 17 mobilenet_v2 240
python tools/merge_traffic.py gpurun_out/traffic_resnet50_conv.json gpurun_out/traffic_mobilenet_v2_dw.json > gpurun_out/traffic.json && cp gpurun_out/traffic.json profiles/r2/traffic.json
rm -f gpurun_out/prof_*.raw.csv.tmp
timeout 200 python tools/conv_layers.py --json gpurun_out/conv_layers_resnet50.json > gpurun_out/conv_layers_resnet50.txt 2>&1; tail -1 gpurun_out/conv_layers_resnet50.txt
timeout 120 python tools/dw_layers.py > gpurun_out/dw_layers_cold.txt 2>&1; tail -1 gpurun_out/dw_layers_cold.txt
timeout 200 python -m pytest tests/test_graph_gpu.py -q -m gpu --timeout 150 --tb=short > gpurun_out/pytest_graph_gpu.log 2>&1; tail -3 gpurun_out/pytest_graph_gpu.log
timeout 300 python bench.py --layers > gpurun_out/bench_resnet50.json 2> gpurun_out/bench_resnet50.layers; tail -c 300 gpurun_out/bench_resnet50.json; tail -1 gpurun_out/bench_resnet50.layers
grep -c "mbarrier wait timed out" gpurun_out/*.json gpurun_out/*.log gpurun_out/*.txt | grep -v ":0$"
du -sh gpurun_out
