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

# ---- the same copy while the SMs are busy (a bf16 matmul loop on another stream), with 1 and 2 concurrent copy streams
import time
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
src = [torch.empty(154 * 1000 * 1000, dtype=torch.uint8).pin_memory() for _ in range(2)]
dst = [torch.empty(154 * 1000 * 1000, dtype=torch.uint8, device="cuda") for _ in range(2)]
big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
big2 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
res = {}
for load in ("idle", "matmul", "memcpy"):
    for ncopy in (1, 2):
        cs = [torch.cuda.Stream() for _ in range(ncopy)]
        ws = torch.cuda.Stream()
        torch.cuda.synchronize()
        stop = time.perf_counter() + 0.25
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(ncopy)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(ncopy)]
        reps = 8
        for i in range(ncopy):
            e0[i].record(cs[i])
        for r in range(reps):
            if load == "matmul":
                with torch.cuda.stream(ws):
                    for _ in range(6):
                        torch.matmul(a, b)
            elif load == "memcpy":
                with torch.cuda.stream(ws):
                    for _ in range(3):
                        big2.copy_(big)
            for i in range(ncopy):
                with torch.cuda.stream(cs[i]):
                    dst[i].copy_(src[i], non_blocking=True)
        for i in range(ncopy):
            e1[i].record(cs[i])
        torch.cuda.synchronize()
        t = max(e0[i].elapsed_time(e1[i]) for i in range(ncopy)) * 1e-3
        res["h2d_%s_%dstream_gbs" % (load, ncopy)] = 154e6 * reps * ncopy / t / 1e9
print(json.dumps(res))
