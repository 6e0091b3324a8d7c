#!/usr/bin/env python3
"""Sum DRAM bytes and durations over the launches of an `ncu --page raw --csv` dump -> one JSON object (profiles/rN/traffic.json entry).
    python tools/ncu_traffic.py gpurun_out/prof_tc_gemm.raw.csv [kernel-name-regex]"""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
rx = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
def val(r, name):
    i = col.get(name)
    if i is None:
        return 0.0
    v = float(r[i].replace(",", "") or 0)
    u = units[i]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return v * scale
out = {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "time_us": 0.0, "kernels": {}}
for r in data:
    name = r[col["Kernel Name"]]
    if rx and not rx.search(name):
        continue
    out["launches"] += 1
    out["dram_read_bytes"] += val(r, "dram__bytes_read.sum")
    out["dram_write_bytes"] += val(r, "dram__bytes_write.sum")
    t = val(r, "gpu__time_duration.sum")
    out["time_us"] += t
    k = re.sub(r"\(.*", "", name).replace("ncnn_cuda::", "").replace("void ", "")
    e = out["kernels"].setdefault(k, {"launches": 0, "time_us": 0.0})
    e["launches"] += 1
    e["time_us"] += t
print(json.dumps(out, indent=1))
