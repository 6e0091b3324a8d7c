#!/usr/bin/env python3
"""bench.py -- throughput of the hot path (batched CNN inference behind ncnn's Net/Extractor API) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload resnet50] [--storage fp16] [--impl reference]

One "step" = one forward pass of the workload's graph over one batch of synthetic images (seeded U(-1,1) input,
seeded random-init weights; BASELINE.json: ResNet-50 224x224 batch 256).  Prints ONE JSON line (rank 0):

  value      images/s, whole job, input blob already resident in HBM, CUDA events on the runtime's own stream
  e2e        the same metric through the reference-facing call (Extractor.input(host Mat) + extract(host Mat)) with
             pinned host buffers: H2D of the fp32 batch and D2H of the result inside the timed region
  roofline   the dominant kernel family (tcgen05 implicit-GEMM conv): algorithmic FLOP of the conv layers of one step /
             their summed CUDA-event time, against the measured dense bf16/fp16 tensor peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, built from /root/reference) on this box's host
             cores, bounded sample of the same workload

Multi-GPU (N > 1, launched by torchrun): one replica of the Net per GPU, each rank runs its own batch; no collective on
the data path (inference shards by batch only).  torch.distributed is used for the barrier and the max-over-ranks time.
`--impl reference` times only the reference CPU path (rank 0) and prints the same line shape with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import modelzoo  # noqa: E402

WORKLOADS = {
    # name: (model, batch per GPU, input size)
    "resnet50": ("resnet50", 256, 224),
    "mobilenet_v2": ("mobilenet_v2", 128, 224),
    "vgg16": ("vgg16", 256, 224),
    "squeezenet_v1_1": ("squeezenet_v1_1", 1, 227),
    "yolov8s": ("yolov8s", 64, 640),
}
WEIGHT_SEED = 7767517


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback")


def with_input_size(text, size):
    lines = text.splitlines()
    for i, l in enumerate(lines):
        if l.startswith("Input"):
            tok = l.split()
            tok = [("0=%d" % size) if t.startswith("0=") else (("1=%d" % size) if t.startswith("1=") else t) for t in tok]
            lines[i] = " ".join(tok)
    return "\n".join(lines) + "\n"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md's clocks line)"""

    def __init__(self, gpu_index):
        threading.Thread.__init__(self, daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append((time.time(), line.strip()))
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self, t0, t1):
        sm, mx, reasons = [], 0, set()
        for ts, line in self.samples:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": mx or None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(model, size, batch, repeats, threads=None):
    """the reference's own CPU path (oracle/_ref) on this box: benchncnn's options (winograd/sgemm/packing/fp16 defaults,
    benchmark/benchncnn.cpp:346-365), a batched Mat so the reference takes its own per-sample loop (src/net.cpp:654-705)"""
    from oracle import ref as oref
    R = oref.reference()
    threads = threads or R.cpu_count()
    text = with_input_size(modelzoo.param_text(model), size)
    weights = modelzoo.random_model_bytes(text, seed=WEIGHT_SEED)
    opt = R.make_option(threads, use_vulkan_compute=0)
    net = oref.Net(R, text, weights, opt)
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (batch, 3, size, size)).astype(np.float32)
    name = net.input_names[0]
    net.run({name: x[:1]}, batched=True)  # warm-up (weight repack caches, thread pool)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        net.run({name: x}, batched=True)
        times.append(time.perf_counter() - t0)
    net.close()
    return dict(images_per_s=batch / float(np.mean(times)), seconds=times, threads=threads, batch=batch, kind="reference", lib=os.path.basename(R.path))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="resnet50", choices=sorted(WORKLOADS))
    ap.add_argument("--storage", default="bf16", choices=["fp16", "bf16", "fp32"],
                    help="element type of device blobs; bf16 is the north-star dtype (fp32 in, bf16 tensor cores, fp32 accumulate)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-fusion", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", action="store_true", help="also print the per-layer table to stderr")
    ap.add_argument("--e2e-threads", type=int, default=3,
                    help="host threads feeding the end-to-end path, one Extractor (own stream, own device pool) per call: with 2 the H2D "
                         "copy of one step overlaps the kernels of the previous ones (the reference's one-extractor-per-thread rule); "
                         "measured on ResNet-50 bs256: 1 -> 36 k, 2 -> 51 k, 3 -> 60 k, 4 -> 61 k images/s")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    model, batch, size = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    config = {"workload": "%s %dx%d batch %d per GPU, %s storage" % (model, size, size, batch, args.storage), "global_batch": batch * world,
              "batch_per_gpu": batch, "replicas": world, "parallelism": "replicas x%d (batch split, no collective)" % world,
              "l2": "input batch and every activation blob exceed the 126 MB L2 (no flush needed)"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        cpu_batch = 8 if model != "squeezenet_v1_1" else 16
        reps = max(1, min(args.steps, 3))
        r = cpu_reference_run(model, size, cpu_batch, reps)
        line = {"impl": "reference", "metric": "images/sec", "value": r["images_per_s"], "unit": "images/s", "n_gpus": args.gpus, "gpus_used": 0, "steps": reps, "warmup": 1,
                "ms_per_step": 1000.0 * float(np.mean(r["seconds"])), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": dict(config, workload="%s %dx%d, reference CPU path, bounded sample of %d images per step" % (model, size, size, cpu_batch)),
                "cpu_baseline": {"value": r["images_per_s"], "unit": "images/s", "cores": r["threads"], "kind": "reference",
                                 "sample": "%d steps of a %d-image batch through %s, benchncnn options" % (reps, cpu_batch, r["lib"])},
                "e2e": {"value": r["images_per_s"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    from ncnn_b200 import replicas  # torch.distributed plumbing only: rendezvous, barrier, max-over-ranks
    group = replicas.Group(backend="nccl" if world > 1 else None)

    from ncnn_b200 import runner
    text = with_input_size(modelzoo.param_text(model), size)
    weights = modelzoo.random_model_bytes(text, seed=WEIGHT_SEED)
    sess = runner.Session(text, weights, storage=args.storage, device=local_rank, fusion=not args.no_fusion)
    del weights
    rng = np.random.default_rng(1 + rank)
    x = rng.uniform(-1, 1, (batch, 3, size, size)).astype(np.float32)
    host_in = sess.pinned_input(x)
    dev_in = sess.upload(host_in)
    lib = sess.L.lib

    def barrier():
        sess.sync()
        group.barrier()

    max_over_ranks = group.max

    # ---- device-resident throughput
    for _ in range(max(args.warmup, 3)):
        lib.ncnn_cuda_mat_destroy(sess.enqueue_device(dev_in))
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = sess.event(), sess.event()
    launches0 = sess.launch_count()
    t_wall0 = time.time()
    barrier()
    sess.record(e0)
    for _ in range(args.steps):
        lib.ncnn_cuda_mat_destroy(sess.enqueue_device(dev_in))
    sess.record(e1)
    ms = sess.elapsed_ms(e0, e1)
    barrier()
    t_wall1 = time.time()
    launches = sess.launch_count() - launches0
    ms = max_over_ranks(ms)
    value = replicas.throughput(batch, args.steps, world, ms)

    # ---- end to end through the reference-facing call (host Mat in, host Mat out)
    # every step = ncnn_extractor_input(pinned host Mat) + ncnn_extractor_extract(host Mat): H2D of the fp32 batch, the
    # kernels, D2H of the result, all inside the timed region.  Steps are dealt to --e2e-threads host threads, each with
    # its own input Mat and its own Extractor per step (ctypes releases the GIL), so consecutive steps overlap on the GPU's
    # copy and compute engines; with 1 thread the steps run strictly one after the other.
    def e2e_run(n_threads, steps):
        inputs = [host_in] + [sess.pinned_input(x) for _ in range(n_threads - 1)]
        per = [steps // n_threads + (1 if i < steps % n_threads else 0) for i in range(n_threads)]
        errs = []

        def worker(i):
            try:
                lib.ncnn_cuda_set_device(local_rank)
                for _ in range(per[i]):
                    lib.ncnn_mat_destroy(sess.extract_host(inputs[i]))
            except Exception as e:  # surfaced below: a failed step must fail the bench
                errs.append(e)

        threads = [threading.Thread(target=worker, args=(i,)) for i in range(n_threads)]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        dt = time.perf_counter() - t0
        for m in inputs[1:]:
            lib.ncnn_mat_destroy(m)
        if errs:
            raise errs[0]
        return dt

    for _ in range(3):
        lib.ncnn_mat_destroy(sess.extract_host(host_in))
    barrier()
    e2e_serial_s = max_over_ranks(e2e_run(1, args.steps))
    barrier()
    nthreads = max(1, args.e2e_threads)
    if nthreads > 1:
        e2e_run(nthreads, 3 * nthreads)  # untimed: every thread's stream + device pool is warm before the timed steps
        barrier()
        e2e_s = max_over_ranks(e2e_run(nthreads, args.steps))
        if e2e_s > e2e_serial_s:  # launch-bound small networks gain nothing from a second feeder thread: report the serial run
            e2e_s, nthreads = e2e_serial_s, 1
    else:
        e2e_s = e2e_serial_s
    e2e_value = replicas.throughput(batch, args.steps, world, e2e_s * 1000.0)
    e2e_serial_value = replicas.throughput(batch, args.steps, world, e2e_serial_s * 1000.0)
    h2d, d2h = int(sess.last_h2d), int(sess.last_d2h)

    # ---- the same with device pre-processing: the step's input is the batch of 8-bit BGR images a deployment
    # actually holds; from_pixels + substract_mean_normalize run on the device (ncnn_extractor_input_pixels), so 1/4 of the bytes
    # cross PCIe.  Reported beside, not instead of, the fp32-Mat figure.
    px = (np.random.default_rng(2 + rank).integers(0, 256, (batch, size, size, 3), dtype=np.uint8))
    mean_vals = np.asarray([104.0, 117.0, 123.0], np.float32)
    norm_vals = np.asarray([0.017, 0.0175, 0.0171], np.float32)

    def e2e_pixels_run(n_threads, steps, decode=None):
        bufs = [sess.pinned_pixels(px) for _ in range(n_threads)]
        per = [steps // n_threads + (1 if i < steps % n_threads else 0) for i in range(n_threads)]
        errs = []

        def worker(i):
            try:
                lib.ncnn_cuda_set_device(local_rank)
                for _ in range(per[i]):
                    lib.ncnn_mat_destroy(sess.extract_host_pixels(bufs[i][1], bufs[i][2], 2, mean_vals, norm_vals, yolov8_decode=decode))
            except Exception as e:
                errs.append(e)

        threads = [threading.Thread(target=worker, args=(i,)) for i in range(n_threads)]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        dt = time.perf_counter() - t0
        for b in bufs:
            lib.ncnn_mat_destroy(b[0])
        if errs:
            raise errs[0]
        return dt

    e2e_pixels_value = None
    h2d_pixels = None
    try:
        e2e_pixels_run(nthreads, 2 * nthreads)
        barrier()
        pix_s = max_over_ranks(e2e_pixels_run(nthreads, args.steps))
        e2e_pixels_value = replicas.throughput(batch, args.steps, world, pix_s * 1000.0)
        h2d_pixels = int(sess.last_h2d_pixels)
    except Exception as e:  # the fp32-Mat e2e above is the contract figure; this one is extra
        sys.stderr.write("e2e with pixel input failed: %s\n" % e)
    # ---- detection heads: pixel input AND the decode of the prediction blob on the device (examples/yolov8.cpp generate_proposals,
    # ncnn_extractor_extract_yolov8_proposals): 6 floats per anchor come back instead of 64 + classes
    e2e_decoded_value = d2h_decoded = None
    if model == "yolov8s":
        try:
            dec = ([8, 16, 32], 0.25)
            e2e_pixels_run(nthreads, 2 * nthreads, decode=dec)
            barrier()
            dec_s = max_over_ranks(e2e_pixels_run(nthreads, args.steps, decode=dec))
            e2e_decoded_value = replicas.throughput(batch, args.steps, world, dec_s * 1000.0)
            d2h_decoded = int(sess.last_d2h_pixels)
        except Exception as e:
            sys.stderr.write("e2e with device decode failed: %s\n" % e)
    clocks = sampler.summary(t_wall0, t_wall1) if rank == 0 else None
    if rank == 0:
        sampler.stop()

    # ---- per-layer device times (CUDA events around every layer on the same stream) -> roofline of the dominant kernel
    peaks = load_peaks()
    prof = sess.profile(dev_in, repeats=max(3, min(args.steps, 10)))
    work = runner.layer_work(text, prof)
    esize = 4 if args.storage == "fp32" else 2
    conv_ms = conv_flop = dw_ms = dw_bytes = fc_ms = fc_flop = 0.0
    total_ms = sum(p[3] for p in prof)
    rows = []
    for li, t, name, lms, shape in prof:
        w = work.get(li)
        row = {"layer": name, "type": t, "ms": lms}
        if w and t == "Convolution":
            conv_ms += lms
            conv_flop += 2.0 * w["macs"]
            row["tflops"] = 2.0 * w["macs"] / (lms * 1e-3) / 1e12 if lms > 0 else None
        elif w and t == "ConvolutionDepthWise":
            s = w.get("s", 1)
            b = (w["out_elems"] * s * s + w["out_elems"]) * esize + w["weights"] * 4
            dw_ms += lms
            dw_bytes += b
            row["gbs"] = b / (lms * 1e-3) / 1e9 if lms > 0 else None
        elif w and t == "InnerProduct":
            fc_ms += lms
            fc_flop += 2.0 * w["macs"]
        rows.append(row)
    if args.layers and rank == 0:
        for r in rows:
            sys.stderr.write("%-28s %-22s %8.4f ms %s\n" % (r["layer"][:28], r["type"], r["ms"],
                                                          ("%7.1f TFLOP/s" % r["tflops"]) if r.get("tflops") else (("%7.1f GB/s" % r["gbs"]) if r.get("gbs") else "")))
        sys.stderr.write("layers total %.3f ms (conv %.3f, dw %.3f, fc %.3f); step %.3f ms\n" % (total_ms, conv_ms, dw_ms, fc_ms, ms / args.steps))

    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1c", "traffic.json")
    if os.path.exists(tpath) and batch == WORKLOADS[args.workload][1]:
        t = json.load(open(tpath)).get(args.workload)
        if t and t.get("storage") == args.storage:
            # DRAM bytes of the dominant kernel family over one step, from the committed ncu --set full capture of this command
            traffic = {"dram_bytes_per_step": t["dram_read_bytes"] + t["dram_write_bytes"], "launches": t["launches"], "source": "profiles/r1c/traffic.json (ncu)"}
    # MobileNetV2 is the depthwise (bandwidth) configuration of BASELINE.json: its roofline line is the depthwise family
    if dw_bytes > 0 and (dw_ms > conv_ms or args.workload == "mobilenet_v2"):
        achieved = dw_bytes / (dw_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "dwconv3x3_tma_kernel (ConvolutionDepthWise, TMA halo tiles)", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm"], "traffic": traffic, "peak_source": peaks["source"] + " copy bandwidth",
                    "algorithmic_bytes_per_step": dw_bytes, "dw_ms_per_step": dw_ms,
                    "share_of_step": dw_ms / total_ms if total_ms else None}
    else:
        achieved = conv_flop / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": "tc_gemm_kernel (Convolution, tcgen05 implicit GEMM)" if args.storage != "fp32" else "conv_simt_kernel (fp32 CUDA cores)",
                    "achieved": achieved, "peak": peaks["tensor_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tensor_sustained"], "traffic": traffic,
                    "peak_source": peaks["source"] + " sustained dense bf16 (kernel timed inside a long step)",
                    "share_of_step": conv_ms / total_ms if total_ms else None,
                    "algorithmic_flop_per_step": conv_flop, "conv_ms_per_step": conv_ms}
    if dw_bytes > 0:
        roofline["depthwise"] = {"achieved_gbs": dw_bytes / (dw_ms * 1e-3) / 1e9, "frac_of_hbm": dw_bytes / (dw_ms * 1e-3) / 1e9 / peaks["hbm"], "ms_per_step": dw_ms}

    line = {"metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.storage], "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "host_threads": nthreads,
                    "serial_value": e2e_serial_value,
                    "pixels_value": e2e_pixels_value, "pixels_h2d_bytes_per_step": h2d_pixels,
                    "pixels_decoded_value": e2e_decoded_value, "pixels_decoded_d2h_bytes_per_step": d2h_decoded,
                    "mode": "each step = extractor.input(pinned host Mat) + extract(host Mat); steps dealt to %d host thread(s), one Extractor/stream per step" % nthreads},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "fused_layers": sess.fused_layers}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_batch = 8 if model != "squeezenet_v1_1" else 16
            r = cpu_reference_run(model, size, cpu_batch, 2)
            line["cpu_baseline"] = {"value": r["images_per_s"], "unit": "images/s", "cores": r["threads"], "kind": "reference",
                                    "sample": "2 steps of a %d-image batch of the same workload through %s (benchncnn options)" % (cpu_batch, r["lib"])}
        except Exception as e:  # the oracle library is test infrastructure: report, do not fail the product bench
            line["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    if rank == 0:
        print(json.dumps(line))
    sess.close()
    group.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
