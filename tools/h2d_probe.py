#!/usr/bin/env python3
"""tools/h2d_probe.py -- pinned host -> device copy bandwidth of this box (what bounds bench.py's e2e leg when the input is fp32)."""
import json, sys, torch
out = {}
for mb in (16, 154, 616):
    n = mb * 1000 * 1000
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    out["h2d_%dMB_gbs" % mb] = n * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    e0.record()
    for _ in range(10):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    out["d2h_%dMB_gbs" % mb] = n * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9
print(json.dumps(out))
