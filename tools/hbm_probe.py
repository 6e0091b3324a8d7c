#!/usr/bin/env python3
"""tools/hbm_probe.py -- what this B200 sustains for pure reads, pure writes and read:write mixes (plain torch kernels,
CUDA events, buffers >> L2): the per-layer floors of write-dominated layers (1x1 convs that expand channels, the stem)
are below the 1:1 copy figure of MEASURED_PEAKS.json."""
import json
import sys

import torch


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3


def main():
    n = 1 << 29  # 512 Mi fp16 elements = 1 GiB
    a = torch.empty(n, dtype=torch.float16, device="cuda").normal_()
    b = torch.empty(n, dtype=torch.float16, device="cuda")
    out = {}
    out["write_only_gbs"] = n * 2 / timed(lambda: b.fill_(1.0)) / 1e9
    out["read_only_gbs"] = n * 2 / timed(lambda: a.view(torch.int32).max()) / 1e9
    out["copy_1to1_gbs"] = n * 4 / timed(lambda: b.copy_(a)) / 1e9
    # 1 read : 4 writes (a 64 -> 256 channel 1x1 conv): broadcast a quarter-size source over the destination
    q = a[: n // 4]
    out["read1_write4_gbs"] = (n // 4 * 2 + n * 2) / timed(lambda: b.view(4, n // 4).copy_(q.unsqueeze(0).expand(4, n // 4))) / 1e9
    # 4 reads : 1 write (256 -> 64): sum four quarters into one
    c = torch.empty(n // 4, dtype=torch.float16, device="cuda")
    out["read4_write1_gbs"] = (n * 2 + n // 4 * 2) / timed(lambda: torch.sum(a.view(4, n // 4), dim=0, out=c)) / 1e9
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
