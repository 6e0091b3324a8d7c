#!/usr/bin/env python3
"""tools/e2e_probe.py -- what bounds the end-to-end leg?  The same host-Mat-in / host-Mat-out calls as bench.py's e2e leg, on
(a) the real network and (b) a network with (almost) no compute behind the same 154 MB input (Input -> global average Pooling):
(b) is the rate of the upload pipeline alone."""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import modelzoo, bench
from ncnn_b200 import runner

def run(sess, x, nthreads, steps):
    lib = sess.L.lib
    inputs = [sess.pinned_input(x) for _ in range(nthreads)]
    per = [steps // nthreads + (1 if i < steps % nthreads else 0) for i in range(nthreads)]
    def worker(i):
        lib.ncnn_cuda_set_device(0)
        for _ in range(per[i]):
            lib.ncnn_mat_destroy(sess.extract_host(inputs[i]))
    for rep in range(2):
        th = [threading.Thread(target=worker, args=(i,)) for i in range(nthreads)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
    return dt / steps * 1e3

model, batch, size = bench.WORKLOADS["resnet50"]
x = np.random.default_rng(1).uniform(-1, 1, (batch, 3, size, size)).astype(np.float32)
text = bench.with_input_size(modelzoo.param_text(model), size)
weights = modelzoo.random_model_bytes(text, seed=bench.WEIGHT_SEED)
thin = "7767517\n2 2\nInput data 0 1 data 0=%d 1=%d 2=3\nPooling output 1 1 data output 0=1 4=1\n" % (size, size)
for name, t, w in (("upload only (Input -> global Pooling)", thin, b""), ("resnet50", text, weights)):
    s = runner.Session(t, w, storage="fp16", device=0)
    for nt in (1, 2, 3, 4):
        ms = run(s, x, nt, 24)
        print("%-40s threads %d: %.3f ms/step  %.1f GB/s H2D  %.0f images/s" % (name, nt, ms, x.nbytes / ms / 1e6, batch / ms * 1e3))
    s.close()
