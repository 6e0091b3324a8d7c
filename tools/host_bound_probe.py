#!/usr/bin/env python3
"""tools/host_bound_probe.py -- is a workload's device-resident step bound by the GPU or by the host enqueueing it?
Times N forward walks two ways: wall clock of the enqueue loop alone (no sync inside; the host's cost per step) and CUDA events
around the same loop (the device's).  host >= device means the GPU waits for launches."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import modelzoo
import bench
from ncnn_b200 import runner

for wl in sys.argv[1:] or ["resnet50", "mobilenet_v2", "yolov8s", "squeezenet_v1_1"]:
    model, batch, size = bench.WORKLOADS[wl]
    text = bench.with_input_size(modelzoo.param_text(model), size)
    weights = modelzoo.random_model_bytes(text, seed=bench.WEIGHT_SEED)
    s = runner.Session(text, weights, storage="fp16", device=0)
    lib = s.L.lib
    x = np.random.default_rng(1).uniform(-1, 1, (batch, 3, size, size)).astype(np.float32)
    d = s.upload(s.pinned_input(x))
    for _ in range(5):
        lib.ncnn_cuda_mat_destroy(s.enqueue_device(d))
    s.sync()
    n = 30
    e0, e1 = s.event(), s.event()
    l0 = s.launch_count()
    s.record(e0)
    t0 = time.perf_counter()
    for _ in range(n):
        lib.ncnn_cuda_mat_destroy(s.enqueue_device(d))
    t1 = time.perf_counter()
    s.record(e1)
    s.sync()
    t2 = time.perf_counter()
    print("%-16s batch %3d: host enqueue %.3f ms/step, device %.3f ms/step, wall incl. sync %.3f ms/step, %d launches/step"
          % (wl, batch, (t1 - t0) / n * 1e3, s.elapsed_ms(e0, e1) / n, (t2 - t0) / n * 1e3, (s.launch_count() - l0) // n))
    s.close()
