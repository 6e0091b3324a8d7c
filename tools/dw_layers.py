#!/usr/bin/env python3
"""tools/dw_layers.py -- the 17 depthwise layers of MobileNetV2 (batch 128), each timed ALONE through the kernel C ABI
(ncnn_cuda_dwconv2d_create / _forward), against the HBM roofline: algorithmic bytes = input + output at the storage
type + fp32 filters.  Buffers rotate through more than the 126 MB L2 (--warm keeps ONE buffer: what a layer sees inside
the network when its input was just written by the previous layer and fits in L2).
--check compares with torch's fp32 depthwise conv2d (development sanity check; parity proper: tests/test_kernels_gpu.py)."""
import argparse
import ctypes as C
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cabi  # noqa: E402

# (name, channels, input size, stride)
MOBILENET_V2_DW = [("block1", 32, 112, 1), ("block2", 96, 112, 2), ("block3", 144, 56, 1), ("block4", 144, 56, 2), ("block5", 192, 28, 1), ("block6", 192, 28, 1),
                   ("block7", 192, 28, 2), ("block8", 384, 14, 1), ("block9", 384, 14, 1), ("block10", 384, 14, 1), ("block11", 384, 14, 1), ("block12", 576, 14, 1),
                   ("block13", 576, 14, 1), ("block14", 576, 14, 2), ("block15", 960, 7, 1), ("block16", 960, 7, 1), ("block17", 960, 7, 1)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--storage", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--warm", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    import torch
    import torch.nn.functional as F
    L = cabi.lib()
    et = cabi.F16 if args.storage == "fp16" else cabi.BF16
    dt = cabi.torch_dtype(et)
    hbm = 6470.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        hbm = json.load(open(pk))["hbm_gbs"]
    n = args.batch
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda")
    g.manual_seed(99)
    print("# mobilenet_v2 depthwise layers, batch %d %s, %s; HBM peak %.0f GB/s" % (n, args.storage, "L2-warm" if args.warm else "cold (rotating buffers)", hbm))
    print("%-10s %5s %5s %2s %9s %8s %6s %9s" % ("layer", "C", "in", "s", "us", "GB/s", "frac", "err"))
    tot_us = tot_b = 0.0
    rows = []
    seen = {}
    for name, ch, size, s in MOBILENET_V2_DW:
        if args.only and not re.search(args.only, name):
            continue
        key = (ch, size, s)
        if key in seen:
            us, abytes, err = seen[key]
        else:
            out = (size + 2 - 3) // s + 1
            in_b, out_b = n * size * size * ch * 2, n * out * out * ch * 2
            R = 1 if args.warm else max(2, min(8, int(300e6 // (in_b + out_b)) + 1))
            wt = ((torch.rand((ch, 1, 3, 3), generator=g, device="cuda") * 2 - 1) * 0.5).to(dt).float()
            bias = torch.rand((ch,), generator=g, device="cuda") * 2 - 1
            desc = cabi.DwConvDesc(ch, ch, ch, 3, 3, 1, 1, s, s, 0.0, 1, cabi.act(3, 0.0, 6.0), et)
            handle = C.c_void_p()
            wa, wp = cabi.fptr(wt.cpu().numpy())
            ba, bp = cabi.fptr(bias.cpu().numpy())
            cabi.check(L.ncnn_cuda_dwconv2d_create(C.byref(handle), C.byref(desc), wp, bp, None), "dwconv2d_create")
            xs, ys, ds = [], [], []
            for r in range(R):
                x = (torch.rand((n, size * size, ch), generator=g, device="cuda") * 2 - 1).to(dt)
                y = torch.full((n, out * out, ch), float("nan"), dtype=dt, device="cuda")
                xs.append(x)
                ys.append(y)
                ds.append((cabi.Tensor(x.data_ptr(), 3, size, size, 1, ch, n, et, ch, size * size * ch), cabi.Tensor(y.data_ptr(), 3, out, out, 1, ch, n, et, ch, out * out * ch)))

            def run(i):
                bd, td = ds[i % R]
                cabi.check(L.ncnn_cuda_dwconv2d_forward(handle, C.byref(bd), C.byref(td), 1, 1, stream), "dwconv2d_forward")
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.iters):
                run(i)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1000.0 / args.iters
            err = float("nan")
            if args.check:
                b = n - 1
                x = xs[0][b:b + 1].float().reshape(1, size, size, ch).permute(0, 3, 1, 2)
                want = torch.clamp(F.conv2d(x, wt, bias, stride=s, padding=1, groups=ch), 0.0, 6.0)
                run(0)
                torch.cuda.synchronize()
                got = ys[0][b:b + 1].float().reshape(1, out, out, ch).permute(0, 3, 1, 2)
                d = (got - want).abs() - (2.0 ** -11 if et == cabi.F16 else 2.0 ** -8) * want.abs()
                err = float(d.clamp(min=0).max() / want.abs().max().clamp(min=1e-30))
                if not torch.isfinite(got).all():
                    err = float("inf")
            L.ncnn_cuda_dwconv2d_destroy(handle)
            abytes = in_b + out_b + ch * 9 * 4
            seen[key] = (us, abytes, err)
            del xs, ys, ds
            torch.cuda.empty_cache()
        rows.append(dict(layer=name, C=ch, size=size, stride=s, us=us, gbs=abytes / us / 1e3, err=err))
        tot_us += us
        tot_b += abytes
        print("%-10s %5d %5d %2d %9.2f %8.0f %6.2f %9.2g" % (name, ch, size, s, us, abytes / us / 1e3, abytes / us / 1e3 / hbm, err))
    if tot_us:
        print("# all 17 layers: %.1f us, %.0f GB/s = %.3f of the HBM copy peak" % (tot_us, tot_b / tot_us / 1e3, tot_b / tot_us / 1e3 / hbm))
    if args.json:
        json.dump(dict(rows=rows, total_us=tot_us, gbs=tot_b / tot_us / 1e3 if tot_us else None), open(args.json, "w"), indent=1)
    if args.check and any(not (r["err"] <= 2e-3) for r in rows):
        print("# CHECK FAILED")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
