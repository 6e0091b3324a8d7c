#!/usr/bin/env python3
"""Assemble profiles/rN/traffic.json (read by bench.py for roofline.traffic) from the per-capture sums of tools/ncu_traffic.py.
    python tools/merge_traffic.py traffic_resnet50_conv.json traffic_mobilenet_v2_dw.json > traffic.json"""
import json, sys
conv = json.load(open(sys.argv[1]))
dw = json.load(open(sys.argv[2]))
out = {
    "resnet50": {"kernel": "tc_gemm_kernel + stem_pool_kernel + rows_pack_kernel (every Convolution / InnerProduct launch of one step)", "storage": "fp16",
                 "launches": conv["launches"], "dram_read_bytes": conv["dram_read_bytes"], "dram_write_bytes": conv["dram_write_bytes"], "time_us": conv["time_us"],
                 "per_kernel": conv["kernels"]},
    "mobilenet_v2": {"kernel": "dwconv3x3_tma_kernel", "storage": "fp16", "launches": dw["launches"], "dram_read_bytes": dw["dram_read_bytes"],
                     "dram_write_bytes": dw["dram_write_bytes"], "time_us": dw["time_us"]},
    "how": "ncu --set full --clock-control none over the launches of ONE step of bench.py (the convolution-family launches of resnet50 bs256 fp16, the 17 depthwise "
           "launches of mobilenet_v2 bs128 fp16); sums over those launches; tools/run_gpu_ncu_r2.sh, tools/ncu_traffic.py.  Algorithmic bytes of the resnet50 "
           "convolutions after the stem and projection-shortcut folds: 9.0 GB (r1: 10.67 GB).",
}
json.dump(out, sys.stdout, indent=1)
print()
