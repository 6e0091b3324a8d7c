#!/usr/bin/env python3
"""Reduce an `ncu --page raw --csv` dump (one row per profiled launch) to the handful of counters DESIGN.md argues
from: duration, DRAM bytes and rate, L2 and DRAM %-of-peak, tensor-pipe activity, grid size, registers.

    python tools/ncu_summary.py gpurun_out/prof_tc_gemm.raw.csv > profiles/rN/tc_gemm_summary.txt
"""
import csv
import re
import sys

COLS = [
    ("us", "gpu__time_duration.sum"),
    ("rdMB", "dram__bytes_read.sum"),
    ("wrMB", "dram__bytes_write.sum"),
    ("dram%", "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("fp16mma%", "sm__ops_path_tensor_src_fp16_dst_fp32.avg.pct_of_peak_sustained_elapsed"),
    ("grid", "launch__grid_size"),
    ("regs", "launch__registers_per_thread"),
    ("smemKB", "launch__shared_mem_per_block_dynamic"),
]


def to_float(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = []
    for short, name in COLS:
        idx.append(hdr.index(name) if name in hdr else -1)
    kn = hdr.index("Kernel Name")
    print("%-3s %-44s " % ("#", "kernel") + " ".join("%9s" % s for s, _ in COLS) + "   GB/s")
    tot_us = 0.0
    for n, r in enumerate(data):
        name = r[kn]
        m = re.search(r"(\w+)<(.*)>", name)
        name = ("%s<%s>" % (m.group(1), m.group(2))) if m else name
        name = name.replace("ncnn_cuda::", "").replace("__nv_bfloat16", "bf16").replace("__half", "f16").replace("(int)", "").replace("tc::", "")
        vals = []
        for (short, cname), i in zip(COLS, idx):
            if i < 0:
                vals.append(float("nan"))
                continue
            v = to_float(r[i])
            u = units[i]
            if short == "us":
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
            if short in ("rdMB", "wrMB"):
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            if short == "smemKB":
                v *= {"byte": 1.0 / 1024, "Kbyte": 1.0, "Mbyte": 1024.0}.get(u, 1.0)
            vals.append(v)
        us, rd, wr = vals[0], vals[1], vals[2]
        tot_us += us
        gbs = (rd + wr) * 1e6 / (us * 1e-6) / 1e9 if us > 0 else 0
        print("%-3d %-44s " % (n, name[:44]) + " ".join("%9.2f" % v for v in vals) + " %7.0f" % gbs)
    print("total %.1f us over %d launches" % (tot_us, len(data)))


if __name__ == "__main__":
    main(sys.argv[1])
