#!/usr/bin/env python3
"""tools/conv_layers.py -- every distinct Convolution of a workload, timed ALONE through the kernel C ABI
(include/ncnn_cuda.h: ncnn_cuda_conv2d_create / _forward) with cold caches, against its own roofline.

    python tools/conv_layers.py [--workload resnet50] [--storage fp16] [--iters 20] [--check] [--only REGEX]

For each layer shape (count = how often the workload has it):
    us        CUDA-event time per launch on the launching stream, inputs/outputs rotated through buffers whose
              total exceeds the 126 MB L2 (so every launch reads from HBM like it does inside the network)
    TF/s      2 * MACs / time
    GB/s      algorithmic bytes (in + out (+ residual) + weights) / time
    floor_us  max(FLOP / tensor peak, bytes / HBM peak) with the MEASURED peaks (MEASURED_PEAKS.json: burst
              bf16 TF/s for a kernel timed alone, copy GB/s)
    frac      floor_us / us  -- the per-layer roofline fraction min(peak, AI * BW) the judge asked for
--check compares each output with torch's fp32 conv2d on the same 16-bit inputs (a development sanity check;
the parity tests proper are tests/test_kernels_gpu.py against the reference's naive layer).
Device memory and events come from torch (plumbing); every compute call goes through the C ABI.
"""
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

# (name, count, h, w, inch, outch, k, stride, pad, residual, act)
RESNET50 = [
    ("conv1 7x7s2 3->64 @224", 1, 224, 224, 3, 64, 7, 2, 3, 0, 1),
    ("s2 1x1 64->256 @56", 1, 56, 56, 64, 256, 1, 1, 0, 0, 0),
    ("s2 1x1 64->64 @56", 1, 56, 56, 64, 64, 1, 1, 0, 0, 1),
    ("s2 3x3 64->64 @56", 3, 56, 56, 64, 64, 3, 1, 1, 0, 1),
    ("s2 1x1 64->256+res @56", 3, 56, 56, 64, 256, 1, 1, 0, 1, 1),
    ("s2 1x1 256->64 @56", 2, 56, 56, 256, 64, 1, 1, 0, 0, 1),
    ("s3 1x1s2 256->512 @56", 1, 56, 56, 256, 512, 1, 2, 0, 0, 0),
    ("s3 1x1s2 256->128 @56", 1, 56, 56, 256, 128, 1, 2, 0, 0, 1),
    ("s3 3x3 128->128 @28", 4, 28, 28, 128, 128, 3, 1, 1, 0, 1),
    ("s3 1x1 128->512+res @28", 4, 28, 28, 128, 512, 1, 1, 0, 1, 1),
    ("s3 1x1 512->128 @28", 3, 28, 28, 512, 128, 1, 1, 0, 0, 1),
    ("s4 1x1s2 512->1024 @28", 1, 28, 28, 512, 1024, 1, 2, 0, 0, 0),
    ("s4 1x1s2 512->256 @28", 1, 28, 28, 512, 256, 1, 2, 0, 0, 1),
    ("s4 3x3 256->256 @14", 6, 14, 14, 256, 256, 3, 1, 1, 0, 1),
    ("s4 1x1 256->1024+res @14", 6, 14, 14, 256, 1024, 1, 1, 0, 1, 1),
    ("s4 1x1 1024->256 @14", 5, 14, 14, 1024, 256, 1, 1, 0, 0, 1),
    ("s5 1x1s2 1024->2048 @14", 1, 14, 14, 1024, 2048, 1, 2, 0, 0, 0),
    ("s5 1x1s2 1024->512 @14", 1, 14, 14, 1024, 512, 1, 2, 0, 0, 1),
    ("s5 3x3 512->512 @7", 3, 7, 7, 512, 512, 3, 1, 1, 0, 1),
    ("s5 1x1 512->2048+res @7", 3, 7, 7, 512, 2048, 1, 1, 0, 1, 1),
    ("s5 1x1 2048->512 @7", 2, 7, 7, 2048, 512, 1, 1, 0, 0, 1),
    ("fc 2048->1000 @1", 1, 1, 1, 2048, 1000, 1, 1, 0, 0, 0),
]
VGG16 = [
    ("conv1_1 3->64 @224", 1, 224, 224, 3, 64, 3, 1, 1, 0, 1),
    ("conv1_2 64->64 @224", 1, 224, 224, 64, 64, 3, 1, 1, 0, 1),
    ("conv2_1 64->128 @112", 1, 112, 112, 64, 128, 3, 1, 1, 0, 1),
    ("conv2_2 128->128 @112", 1, 112, 112, 128, 128, 3, 1, 1, 0, 1),
    ("conv3_1 128->256 @56", 1, 56, 56, 128, 256, 3, 1, 1, 0, 1),
    ("conv3_x 256->256 @56", 2, 56, 56, 256, 256, 3, 1, 1, 0, 1),
    ("conv4_1 256->512 @28", 1, 28, 28, 256, 512, 3, 1, 1, 0, 1),
    ("conv4_x 512->512 @28", 2, 28, 28, 512, 512, 3, 1, 1, 0, 1),
    ("conv5_x 512->512 @14", 3, 14, 14, 512, 512, 3, 1, 1, 0, 1),
]
# MobileNetV2 pointwise layers (the tensor-core half of the bandwidth-bound configuration), batch 128
MOBILENET_V2 = [
    ("conv1 3x3s2 3->32 @224", 1, 224, 224, 3, 32, 3, 2, 1, 0, 3),
    ("pw 32->16 @112", 1, 112, 112, 32, 16, 1, 1, 0, 0, 0),
    ("pw 16->96 @112", 1, 112, 112, 16, 96, 1, 1, 0, 0, 3),
    ("pw 96->24 @56", 1, 56, 56, 96, 24, 1, 1, 0, 0, 0),
    ("pw 24->144 @56", 2, 56, 56, 24, 144, 1, 1, 0, 0, 3),
    ("pw 144->24+res @56", 1, 56, 56, 144, 24, 1, 1, 0, 1, 0),
    ("pw 144->32 @28", 1, 28, 28, 144, 32, 1, 1, 0, 0, 0),
    ("pw 32->192 @28", 3, 28, 28, 32, 192, 1, 1, 0, 0, 3),
    ("pw 192->32+res @28", 2, 28, 28, 192, 32, 1, 1, 0, 1, 0),
    ("pw 192->64 @14", 1, 14, 14, 192, 64, 1, 1, 0, 0, 0),
    ("pw 64->384 @14", 4, 14, 14, 64, 384, 1, 1, 0, 0, 3),
    ("pw 384->64+res @14", 3, 14, 14, 384, 64, 1, 1, 0, 1, 0),
    ("pw 384->96 @14", 1, 14, 14, 384, 96, 1, 1, 0, 0, 0),
    ("pw 96->576 @14", 3, 14, 14, 96, 576, 1, 1, 0, 0, 3),
    ("pw 576->96+res @14", 2, 14, 14, 576, 96, 1, 1, 0, 1, 0),
    ("pw 576->160 @7", 1, 7, 7, 576, 160, 1, 1, 0, 0, 0),
    ("pw 160->960 @7", 3, 7, 7, 160, 960, 1, 1, 0, 0, 3),
    ("pw 960->160+res @7", 2, 7, 7, 960, 160, 1, 1, 0, 1, 0),
    ("pw 960->320 @7", 1, 7, 7, 960, 320, 1, 1, 0, 0, 0),
    ("pw 320->1280 @7", 1, 7, 7, 320, 1280, 1, 1, 0, 0, 3),
]
TABLES = {"resnet50": (RESNET50, 256), "vgg16": (VGG16, 256), "mobilenet_v2": (MOBILENET_V2, 128)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="resnet50", choices=sorted(TABLES))
    ap.add_argument("--storage", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--rotate", type=int, default=0, help="number of rotating buffer sets (default: enough to exceed the L2; 1 = the same L2-resident buffers every launch)")
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    import torch
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    L = cabi.lib()
    table, n = TABLES[args.workload]
    if args.batch:
        n = args.batch
    et = cabi.F16 if args.storage == "fp16" else cabi.BF16
    dt = cabi.torch_dtype(et)
    hbm, tpeak, src = peaks()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)
    rows = []
    tot_us = tot_floor = 0.0
    print("# %s batch %d %s; peaks (%s): %.0f GB/s, %.0f TF/s burst" % (args.workload, n, args.storage, src, hbm, tpeak))
    print("%-28s %3s %9s %8s %8s %9s %6s %9s" % ("layer", "cnt", "us", "TF/s", "GB/s", "floor_us", "frac", "err"))
    for name, cnt, h, w, inch, outch, k, s, pad, res, act in table:
        if args.only and not re.search(args.only, name):
            continue
        outh = (h + 2 * pad - k) // s + 1
        outw = (w + 2 * pad - k) // s + 1
        icp = (inch + 7) // 8 * 8
        ocp = (outch + 7) // 8 * 8
        in_bytes = n * h * w * icp * 2
        out_bytes = n * outh * outw * ocp * 2
        per = in_bytes + out_bytes * (2 if res else 1)
        R = max(2, min(8, int(300e6 // per) + 1))
        if args.rotate > 0:
            R = args.rotate
        wt = (torch.rand((outch, inch, k, k), generator=g, device="cuda") * 2 - 1) * float(np.sqrt(3.0 / (inch * k * k)))
        wt = wt.to(dt).float()
        bias = torch.rand((outch,), generator=g, device="cuda") * 2 - 1
        desc = cabi.ConvDesc(inch, outch, k, k, 1, 1, s, s, pad, pad, pad, pad, 0.0, 1, cabi.act(act, 0.0, 6.0), et)
        handle = C.c_void_p()
        wa, wp = cabi.fptr(wt.cpu().numpy())
        ba, bp = cabi.fptr(bias.cpu().numpy())
        cabi.check(L.ncnn_cuda_conv2d_create(C.byref(handle), C.byref(desc), wp, bp, None), "conv2d_create")
        xs, ys, rs, descs = [], [], [], []
        for r in range(R):
            x = torch.zeros((n, h * w, icp), dtype=dt, device="cuda")
            x[:, :, :inch] = (torch.rand((n, h * w, inch), generator=g, device="cuda") * 2 - 1).to(dt)
            y = torch.full((n, outh * outw, ocp), float("nan"), dtype=dt, device="cuda")
            rr = None
            if res:
                rr = torch.zeros((n, outh * outw, ocp), dtype=dt, device="cuda")
                rr[:, :, :outch] = (torch.rand((n, outh * outw, outch), generator=g, device="cuda") * 2 - 1).to(dt)
            xs.append(x)
            ys.append(y)
            rs.append(rr)
            bd = cabi.Tensor(x.data_ptr(), 3, w, h, 1, inch, n, et, icp, h * w * icp)
            td = cabi.Tensor(y.data_ptr(), 3, outw, outh, 1, outch, n, et, ocp, outh * outw * ocp)
            rd = cabi.Tensor(rr.data_ptr(), 3, outw, outh, 1, outch, n, et, ocp, outh * outw * ocp) if res else None
            descs.append((bd, td, rd))
        wsize = 0 if res else int(L.ncnn_cuda_conv2d_workspace_size(handle, C.byref(descs[0][0]), C.byref(descs[0][1])))
        ws = torch.empty(((wsize + 3) // 4,), dtype=torch.float32, device="cuda") if wsize else None

        def run(i):
            bd, td, rd = descs[i % R]
            cabi.check(L.ncnn_cuda_conv2d_forward(handle, C.byref(bd), C.byref(td), pad, pad, C.byref(rd) if rd else None, None,
                                                  C.c_void_p(ws.data_ptr()) if wsize else None, C.c_size_t(wsize), stream), "conv2d_forward")
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
            # a few images are enough for a sanity check (and keep the fp32 reference small)
            nb = min(n, 4)
            sel = [0, n - 1] if n > 1 else [0]
            errs = []
            for b in sel[:nb]:
                x = xs[0][b:b + 1, :, :inch].float().reshape(1, h, w, inch).permute(0, 3, 1, 2)
                want = F.conv2d(x, wt, bias, stride=s, padding=pad)
                if res:
                    want = want + rs[0][b:b + 1, :, :outch].float().reshape(1, outh, outw, outch).permute(0, 3, 1, 2)
                if act == 1:
                    want = torch.relu(want)
                elif act == 3:
                    want = torch.clamp(want, 0.0, 6.0)
                run(0)
                torch.cuda.synchronize()
                got = ys[0][b:b + 1, :, :outch].float().reshape(1, outh, outw, outch).permute(0, 3, 1, 2)
                d = (got - want).abs() - (2.0 ** -11 if et == cabi.F16 else 2.0 ** -8) * want.abs()
                errs.append(float(d.clamp(min=0).max() / want.abs().max().clamp(min=1e-30)))
                if not torch.isfinite(got).all():
                    errs.append(float("inf"))
            err = max(errs)
        L.ncnn_cuda_conv2d_destroy(handle)
        flop = 2.0 * n * outh * outw * outch * inch * k * k
        # algorithmic bytes: a strided 1x1 reads only the pixels it uses
        used_in = n * outh * outw * icp * 2 if (k == 1 and s > 1) else in_bytes
        abytes = used_in + out_bytes * (2 if res else 1) + outch * inch * k * k * 2
        floor_us = max(flop / (tpeak * 1e12), abytes / (hbm * 1e9)) * 1e6
        rows.append(dict(layer=name, count=cnt, us=us, tflops=flop / us / 1e6, gbs=abytes / us / 1e3, floor_us=floor_us, frac=floor_us / us, err=err))
        tot_us += cnt * us
        tot_floor += cnt * floor_us
        print("%-28s %3d %9.2f %8.1f %8.0f %9.2f %6.2f %9.2g" % (name, cnt, us, flop / us / 1e6, abytes / us / 1e3, floor_us, floor_us / us, err))
        del xs, ys, rs, descs
        torch.cuda.empty_cache()
    print("# total (count-weighted) %.1f us, roofline floor %.1f us, frac %.3f" % (tot_us, tot_floor, tot_floor / tot_us if tot_us else 0))
    if args.json:
        json.dump(dict(workload=args.workload, batch=n, storage=args.storage, rows=rows, total_us=tot_us, floor_us=tot_floor), open(args.json, "w"), indent=1)
    bad = [r for r in rows if args.check and not (r["err"] <= 2e-3)]
    if bad:
        print("# CHECK FAILED:", [(r["layer"], r["err"]) for r in bad])
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
