#!/usr/bin/env python3
"""bench.py -- throughput of the hot path (batched CNN inference behind ncnn's Net/Extractor API) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload resnet50] [--storage fp16|bf16|fp32] [--impl reference]

One "step" = one forward pass of the workload's graph over one batch of synthetic images (seeded U(-1,1) input,
seeded random-init weights; BASELINE.json: ResNet-50 224x224 batch 256).  Prints ONE JSON line (rank 0):

  value      images/s, whole job, input blob already resident in HBM, CUDA events on the runtime's own stream
  e2e        the same metric through the reference-facing call (Extractor.input(host Mat) + extract(host Mat)) with
             pinned host buffers: H2D of the fp32 batch and D2H of the result inside the timed region
  roofline   the dominant kernel family (tcgen05 implicit-GEMM conv): algorithmic FLOP of the conv layers of one step /
             the conv layers' SHARE of the un-profiled step time (per-layer CUDA events give the shares; they break the
             programmatic-dependent-launch overlap, so their sum is scaled to ms_per_step and never exceeds it), against
             the measured dense 16-bit tensor peaks of MEASURED_PEAKS.json: frac_burst (the timed region is a short burst)
             and frac_sustained next to `sustained` = the same step repeated for >= 2 s with its own clock record.
             roofline.depthwise: the MobileNetV2 batch-128 leg (the bandwidth configuration of BASELINE.json) run in the
             same invocation: depthwise layers' algorithmic bytes / their share of that step, as a fraction of HBM copy peak
  launch_bound / graph  BASELINE.json configs[0] (SqueezeNet v1.1 batch 1): the walk eager and replayed as ONE CUDA graph per step
             (runner.Session.capture), resident and through the whole extract
  parity     max|ours - reference| / max|reference| on samples picked from the benched batch, against oracle/_ref (the
             reference's own CPU fp32 path), for the dtype the line reports
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
# the last linear blob of each graph (logits / detection head): where the parity bound is asserted (tests/test_nets_gpu.py)
LOGITS_BLOB = {"squeezenet_v1_1": "pool10", "mobilenet_v2": "fc", "resnet50": "fc1000", "vgg16": "fc8", "yolov8s": None}


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


def numa_bind(local_rank):
    """Bind this process (and the feeder threads and pinned allocations it makes from here on) to the NUMA node of its GPU:
    the end-to-end path is a PCIe H2D stream per rank, and a rank whose pinned pages or feeder threads sit on the other socket
    pulls them across the inter-socket link (round 1: 36 GB/s per GPU alone, 21.6 GB/s with 8 ranks all on node 0)."""
    info = {"node": None, "cpus": None}
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout
        bus = None
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) == 2 and int(f[0]) == local_rank:
                bus = f[1].lower()
        if not bus:
            return info
        if bus.startswith("00000000:"):
            bus = bus[4:]  # sysfs uses a 4-digit PCI domain
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return info
        cpulist = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info = {"node": node, "cpus": len(cpus)}
    except Exception as e:  # binding is an optimisation; report and carry on
        info["error"] = str(e)[:80]
    return info


def layer_shares(sess, text, dev_in, storage, repeats):
    """per-layer CUDA-event times (the recorder's profiling mode) -> rows + the sums of the families the roofline names"""
    from ncnn_b200 import runner
    prof = sess.profile(dev_in, repeats=repeats)
    work = runner.layer_work(text, prof)
    esize = 4 if storage == "fp32" else 2
    acc = dict(conv_ms=0.0, conv_flop=0.0, dw_ms=0.0, dw_bytes=0.0, fc_ms=0.0, fc_flop=0.0, total_ms=sum(p[3] for p in prof))
    rows = []
    for li, t, name, lms, shape in prof:
        w = work.get(li)
        row = {"layer": name, "type": t, "ms": lms}
        if w and t == "Convolution":
            acc["conv_ms"] += lms
            acc["conv_flop"] += 2.0 * w["macs"]
            row["tflops"] = 2.0 * w["macs"] / (lms * 1e-3) / 1e12 if lms > 0 else None
        elif w and t == "ConvolutionDepthWise":
            st = w.get("s", 1)
            b = (w["out_elems"] * st * st + w["out_elems"]) * esize + w["weights"] * 4
            acc["dw_ms"] += lms
            acc["dw_bytes"] += b
            row["gbs"] = b / (lms * 1e-3) / 1e9 if lms > 0 else None
        elif w and t == "InnerProduct":
            acc["fc_ms"] += lms
            acc["fc_flop"] += 2.0 * w["macs"]
        rows.append(row)
    return rows, acc


def algorithmic_conv_flop(text, weights, storage, device, size, batch):
    """2 x MACs of every Convolution of the graph AS WRITTEN (load-time fusion off, one image, scaled by the batch): the roofline's
    numerator must not shrink when a projection shortcut is folded into its neighbour or a stem swallows its pooling layer -- the
    folded layers' arithmetic is still executed"""
    from ncnn_b200 import runner
    s = runner.Session(text, weights, storage=storage, device=device, fusion=False)
    try:
        h = s.pinned_input(np.zeros((1, 3, size, size), np.float32))
        d = s.upload(h)
        prof = s.profile(d, repeats=1)
        work = runner.layer_work(text, prof)
        flop = sum(2.0 * work[li]["macs"] for li, t, name, lms, shape in prof if t == "Convolution" and li in work)
        s.L.lib.ncnn_cuda_mat_destroy(d)
        s.L.lib.ncnn_mat_destroy(h)
    finally:
        s.close()
    return flop * batch


def timed_steps(sess, lib, dev_in, steps, barrier=None):
    """`steps` forward walks with the input resident in HBM, CUDA events on the recorder's own stream -> ms"""
    e0, e1 = sess.event(), sess.event()
    if barrier:
        barrier()
    sess.record(e0)
    for _ in range(steps):
        lib.ncnn_cuda_mat_destroy(sess.enqueue_device(dev_in))
    sess.record(e1)
    return sess.elapsed_ms(e0, e1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="resnet50", choices=sorted(WORKLOADS))
    ap.add_argument("--storage", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="element type of device blobs.  fp16 (the reference's own default, opt.use_fp16_storage) is the contract dtype: fp32 in, "
                         "16-bit tensor cores (tcgen05 kind::f16), fp32 accumulate, and it meets the north-star 2e-3 at network level; bf16 blobs run "
                         "at the same speed but their 8-bit mantissa measures 4e-3..1e-2 through 50 stored layers")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-fusion", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the sustained, depthwise (MobileNetV2), strong-scaling and parity legs")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
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
        # The reference has no batched forward: a batched Mat takes its per-sample loop (src/net.cpp:654-705), so its images/s
        # does not depend on the batch size; the bounded sample is 16 images per step (the full 256 would take ~6 s a step).
        cpu_batch = 16
        reps = max(3, min(args.steps, 5))
        r = cpu_reference_run(model, size, cpu_batch, reps)
        line = {"impl": "reference", "metric": "images/sec", "value": r["images_per_s"], "unit": "images/s", "n_gpus": args.gpus, "gpus_used": 0, "steps": reps, "warmup": 1,
                "ms_per_step": 1000.0 * float(np.mean(r["seconds"])), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": dict(config, workload="%s %dx%d, reference CPU path, bounded sample of %d images per step" % (model, size, size, cpu_batch),
                               sample_note="the reference loops over the samples of a batched Mat (src/net.cpp:654-705): images/s is independent of the batch size, "
                                           "so %d images per step x %d timed steps measure the same rate as the full batch of %d" % (cpu_batch, reps, batch)),
                "cpu_baseline": {"value": r["images_per_s"], "unit": "images/s", "cores": r["threads"], "kind": "reference",
                                 "sample": "%d steps of a %d-image batch through %s, benchncnn options" % (reps, cpu_batch, r["lib"])},
                "e2e": {"value": r["images_per_s"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    numa = numa_bind(local_rank)  # before any pinned allocation or feeder thread exists
    from ncnn_b200 import replicas  # torch.distributed plumbing only: rendezvous, barrier, max-over-ranks
    group = replicas.Group(backend="nccl" if world > 1 else None)

    from ncnn_b200 import runner
    text = with_input_size(modelzoo.param_text(model), size)
    weights = modelzoo.random_model_bytes(text, seed=WEIGHT_SEED)
    sess = runner.Session(text, weights, storage=args.storage, device=local_rank, fusion=not args.no_fusion)
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
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        lib.ncnn_cuda_mat_destroy(sess.enqueue_device(dev_in))
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = sess.launch_count()
    t_wall0 = time.time()
    ms = timed_steps(sess, lib, dev_in, args.steps, barrier)
    barrier()
    t_wall1 = time.time()
    launches = sess.launch_count() - launches0
    ms = max_over_ranks(ms)
    value = replicas.throughput(batch, args.steps, world, ms)
    ms_per_step = ms / args.steps

    # ---- the same step repeated for >= 2 s: the figure a serving loop sees once the 1 kW power cap has pulled the clocks down
    sustained = None
    if not args.no_extra_legs and args.sustained_seconds > 0:
        n_sus = max(args.steps, int(args.sustained_seconds * 1000.0 / max(ms_per_step, 1e-3)) + 1)
        t_s0 = time.time()
        ms_sus = max_over_ranks(timed_steps(sess, lib, dev_in, n_sus, barrier))
        barrier()
        t_s1 = time.time()
        sustained = {"value": replicas.throughput(batch, n_sus, world, ms_sus), "unit": "images/s", "steps": n_sus, "ms_per_step": ms_sus / n_sus,
                     "seconds": ms_sus / 1000.0, "clocks": sampler.summary(t_s0, t_s1) if rank == 0 else None}

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

    # ---- strong scaling: the SAME global batch (the workload's, 256 for ResNet-50) split over the replicas with Mat::batch_range
    # views (src/mat.h:241-242); efficiency = t(global batch on one GPU) / (N * max over ranks of t(global batch / N))
    strong = None
    if not args.no_extra_legs and world > 1 and batch % world == 0:
        sb = batch // world
        view = replicas.batch_view(sess.L, host_in, rank * sb, sb)
        dev_part = sess.upload(view)
        for _ in range(3):
            lib.ncnn_cuda_mat_destroy(sess.enqueue_device(dev_part))
        barrier()
        ms_part = max_over_ranks(timed_steps(sess, lib, dev_part, args.steps, barrier))
        barrier()
        lib.ncnn_cuda_mat_destroy(dev_part)
        lib.ncnn_mat_destroy(view)
        strong = {"global_batch": batch, "batch_per_gpu": sb, "value": batch * args.steps / (ms_part * 1e-3), "unit": "images/s",
                  "ms_per_step": ms_part / args.steps, "efficiency": ms / (world * ms_part),
                  "how": "global batch %d split %d ways through ncnn_mat_batch_range views; efficiency = t(batch %d, one GPU) / (%d x max-over-ranks t(batch %d))"
                         % (batch, world, batch, world, sb)}

    # ---- per-layer device times (CUDA events around every layer on the same stream) -> roofline of the dominant kernel family.
    # The per-layer events serialise the layers (no programmatic-dependent-launch overlap), so their sum exceeds the step: the
    # family's SHARE of that sum is applied to the un-profiled ms_per_step.
    peaks = load_peaks()
    rows, acc = layer_shares(sess, text, dev_in, args.storage, repeats=max(3, min(args.steps, 10)))
    if sess.fused_layers and acc["conv_flop"] > 0:
        try:
            acc["conv_flop"] = algorithmic_conv_flop(text, weights, args.storage, local_rank, size, batch)
        except Exception as e:
            sys.stderr.write("algorithmic FLOP count from the unfused graph failed (%s): the fused layers' own count is used\n" % e)
    if args.layers and rank == 0:
        for r in rows:
            sys.stderr.write("%-28s %-22s %8.4f ms %s\n" % (r["layer"][:28], r["type"], r["ms"],
                                                          ("%7.1f TFLOP/s" % r["tflops"]) if r.get("tflops") else (("%7.1f GB/s" % r["gbs"]) if r.get("gbs") else "")))
        sys.stderr.write("layers total %.3f ms (conv %.3f, dw %.3f, fc %.3f); step %.3f ms\n" % (acc["total_ms"], acc["conv_ms"], acc["dw_ms"], acc["fc_ms"], ms_per_step))
    scale = min(1.0, ms_per_step / acc["total_ms"]) if acc["total_ms"] > 0 else 1.0

    traffic = None
    for tdir in ("r2", "r1c"):
        tpath = os.path.join(ROOT, "profiles", tdir, "traffic.json")
        if traffic is None and os.path.exists(tpath) and batch == WORKLOADS[args.workload][1]:
            t = json.load(open(tpath)).get(args.workload)
            if t and t.get("storage") == args.storage and all(isinstance(t.get(k), (int, float)) and t[k] == t[k] for k in ("dram_read_bytes", "dram_write_bytes")):
                # DRAM bytes of the dominant kernel family over one step, from the committed ncu --set full capture of this command
                traffic = {"dram_bytes_per_step": t["dram_read_bytes"] + t["dram_write_bytes"], "launches": t["launches"], "source": "profiles/%s/traffic.json (ncu)" % tdir}

    def depthwise_block(a, step_ms, sc):
        dw_ms = a["dw_ms"] * sc
        gbs = a["dw_bytes"] / (dw_ms * 1e-3) / 1e9
        return {"achieved_gbs": gbs, "frac_of_hbm": gbs / peaks["hbm"], "ms_per_step": dw_ms, "algorithmic_bytes_per_step": a["dw_bytes"],
                "share_of_step": dw_ms / step_ms, "layers": 17 if model == "mobilenet_v2" else None}

    # MobileNetV2 is the depthwise (bandwidth) configuration of BASELINE.json: its roofline line is the depthwise family
    if acc["dw_bytes"] > 0 and (acc["dw_ms"] > acc["conv_ms"] or args.workload == "mobilenet_v2"):
        d = depthwise_block(acc, ms_per_step, scale)
        roofline = {"bound": "hbm", "kernel": "dwconv3x3 kernels (ConvolutionDepthWise, TMA halo tiles)", "achieved": d["achieved_gbs"], "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": d["frac_of_hbm"], "traffic": traffic, "peak_source": peaks["source"] + " copy bandwidth",
                    "algorithmic_bytes_per_step": acc["dw_bytes"], "dw_ms_per_step": d["ms_per_step"], "share_of_step": d["share_of_step"], "depthwise": d}
    else:
        conv_ms = acc["conv_ms"] * scale
        achieved = acc["conv_flop"] / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": "tc_gemm_kernel (Convolution, tcgen05 implicit GEMM)" if args.storage != "fp32" else "conv_simt_kernel (fp32 CUDA cores)",
                    "achieved": achieved, "peak": peaks["tensor_burst"], "unit": "TFLOP/s", "frac": achieved / peaks["tensor_burst"],
                    "frac_burst": achieved / peaks["tensor_burst"], "frac_sustained": achieved / peaks["tensor_sustained"], "traffic": traffic,
                    "peak_source": peaks["source"] + " burst dense bf16 (the %d-step timed region lasts %.0f ms and does not reach the power cap); frac_sustained uses the "
                                                     "seconds-long figure" % (args.steps, ms),
                    "share_of_step": conv_ms / ms_per_step if ms_per_step else None,
                    "algorithmic_flop_per_step": acc["conv_flop"], "conv_ms_per_step": conv_ms,
                    "profiled_layer_sum_ms": acc["total_ms"], "scaled_by": scale}
        if sustained:
            sus_conv_ms = conv_ms / ms_per_step * sustained["ms_per_step"]
            roofline["sustained_achieved"] = acc["conv_flop"] / (sus_conv_ms * 1e-3) / 1e12
            roofline["sustained_frac_of_sustained_peak"] = roofline["sustained_achieved"] / peaks["tensor_sustained"]
        if acc["dw_bytes"] > 0:
            roofline["depthwise"] = depthwise_block(acc, ms_per_step, scale)

    line = {"metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.storage], "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "host_threads": nthreads,
                    "serial_value": e2e_serial_value,
                    "h2d_gbs_per_gpu": h2d * args.steps / e2e_s / 1e9, "numa": numa,
                    "pixels_value": e2e_pixels_value, "pixels_h2d_bytes_per_step": h2d_pixels,
                    "pixels_decoded_value": e2e_decoded_value, "pixels_decoded_d2h_bytes_per_step": d2h_decoded,
                    "mode": "each step = extractor.input(pinned host Mat) + extract(host Mat); steps dealt to %d host thread(s), one Extractor/stream per step; %s"
                            % (nthreads, "process, feeder threads and pinned buffers bound to the GPU's NUMA node" if numa.get("node") is not None
                               else "no NUMA binding (this box reports numa_node -1 for the GPU)")},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "fused_layers": sess.fused_layers}
    if sustained:
        line["sustained_value"] = sustained["value"]
        line["sustained"] = sustained
    if strong:
        line["strong"] = strong

    # ---- parity of the benched dtype on samples of the benched batch against the reference's CPU fp32 path (oracle/_ref):
    # max|ours - ref| / max|ref| on the network's last linear blob and identical top-1
    if rank == 0 and not args.no_extra_legs:
        try:
            line["parity"] = parity_check(sess, model, text, weights, x, args.storage)
        except Exception as e:  # test infrastructure missing on this box: say so, do not fail the product bench
            line["parity"] = {"dtype": line["dtype"], "max_norm_err": None, "samples": 0, "oracle": "oracle/_ref", "error": str(e)[:120]}
    del weights

    # ---- the bandwidth configuration of BASELINE.json in the same invocation: MobileNetV2 224x224 batch 128, depthwise layers vs HBM
    if not args.no_extra_legs and args.workload == "resnet50" and args.storage != "fp32":
        try:
            d = depthwise_leg(local_rank, args.storage, group, max_over_ranks, args.steps, peaks)
            line["roofline"]["depthwise"] = d
        except Exception as e:
            line["roofline"]["depthwise"] = {"error": str(e)[:120]}
    # ---- launch-bound networks: one CUDA graph per walk.  On the batch-1 workload itself, and -- in the default invocation -- as a
    # secondary leg on BASELINE.json configs[0] (SqueezeNet v1.1 227x227 batch 1)
    if not args.no_extra_legs and rank == 0 and args.storage != "fp32":
        try:
            if batch == 1:
                line["graph"] = graph_leg(sess, host_in, dev_in, batch, max(args.steps, 200))
            elif args.workload == "resnet50":
                line["launch_bound"] = launch_bound_leg(local_rank, args.storage)
        except Exception as e:
            line["graph" if batch == 1 else "launch_bound"] = {"error": str(e)[:160]}
    if rank == 0:
        sampler.stop()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_batch = 16
            r = cpu_reference_run(model, size, cpu_batch, 2)
            line["cpu_baseline"] = {"value": r["images_per_s"], "unit": "images/s", "cores": r["threads"], "kind": "reference",
                                    "sample": "2 steps of a %d-image batch of the same workload through %s (benchncnn options; the reference loops per sample, "
                                              "so the rate does not depend on the batch)" % (cpu_batch, r["lib"])}
        except Exception as e:  # the oracle library is test infrastructure: report, do not fail the product bench
            line["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    if rank == 0:
        print(json.dumps(_finite(line)))
    sess.close()
    group.close()
    return 0


def _finite(o):
    """NaN / inf are not JSON: a strict parser on the driver side must be able to read the line"""
    if isinstance(o, float):
        return o if o == o and o not in (float("inf"), float("-inf")) else None
    if isinstance(o, dict):
        return {k: _finite(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_finite(v) for v in o]
    if isinstance(o, np.floating):
        return _finite(float(o))
    if isinstance(o, np.integer):
        return int(o)
    return o


def parity_check(sess, model, text, weights, x, storage):
    """the benched Session on 4 samples picked across the benched batch vs the reference CPU fp32 path on the same samples"""
    from oracle import ref as oref
    R = oref.reference()
    n = x.shape[0]
    picked = sorted(set([0, n // 3, (2 * n) // 3, n - 1]))
    key = LOGITS_BLOB.get(model)
    ours = sess.run_host(x, blob=key)[picked]  # the whole benched batch through the product, rows picked afterwards
    opt = R.strict_fp32_option(num_threads=R.cpu_count(), packing=True)
    net = oref.Net(R, text, weights, opt)
    try:
        in_name = net.input_names[0]
        want = net.run({in_name: x[picked]}, outputs=[key] if key else None, batched=True)
        want = want[key] if key else list(want.values())[-1]
    finally:
        net.close()
    err = float(np.abs(ours.astype(np.float64) - want).max() / max(np.abs(want).max(), 1e-30))
    out = {"dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[storage], "max_norm_err": err, "samples": len(picked), "picked": picked, "blob": key or "output",
           "metric": "max|ours - ref| / max|ref|", "bound": 1e-5 if storage == "fp32" else 2e-3, "oracle": "oracle/_ref (the reference's CPU fp32 path, %s)" % os.path.basename(R.path)}
    if want.ndim == 2 and want.shape[1] > 1:
        out["top1_identical"] = bool((np.argmax(ours, axis=1) == np.argmax(want, axis=1)).all())
    out["within_bound"] = bool(err <= out["bound"])
    return out


def graph_leg(sess, host_in, dev_in, batch, steps):
    """the same walk eager (one launch per kernel through the recorder) and as ONE cudaGraphLaunch per step (runner.Session.capture:
    ncnn_cuda_graph_begin_capture / _end_capture around the ordinary recorder + Extractor calls), input resident in HBM, CUDA events
    on the recorder's stream; and the whole reference-facing extract (pinned host Mat in -> pinned host Mat out, one stream sync per
    step) eager vs as one graph, wall clock"""
    lib = sess.L.lib
    for _ in range(5):
        lib.ncnn_cuda_mat_destroy(sess.enqueue_device(dev_in))
    sess.sync()
    eager_ms = timed_steps(sess, lib, dev_in, steps) / steps
    g = sess.capture(dev_in=dev_in)
    for _ in range(5):
        g.replay()
    sess.sync()
    e0, e1 = sess.event(), sess.event()
    sess.record(e0)
    for _ in range(steps):
        g.replay()
    sess.record(e1)
    graph_ms = sess.elapsed_ms(e0, e1) / steps
    kernels = g.kernels
    g.close()
    # whole extract, strictly serial (what a batch-1 caller sees): eager Extractor vs one graph + one sync
    for _ in range(3):
        lib.ncnn_mat_destroy(sess.extract_host(host_in))
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.ncnn_mat_destroy(sess.extract_host(host_in))
    eager_e2e_ms = (time.perf_counter() - t0) * 1000.0 / steps
    gh = sess.capture(host_mat=host_in)
    for _ in range(3):
        gh.replay()
        sess.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        gh.replay()
        sess.sync()
    graph_e2e_ms = (time.perf_counter() - t0) * 1000.0 / steps
    gh.close()
    return {"steps": steps, "kernels_per_walk": int(kernels), "eager_ms_per_step": eager_ms, "graph_ms_per_step": graph_ms,
            "eager_images_per_s": batch * 1000.0 / eager_ms, "graph_images_per_s": batch * 1000.0 / graph_ms,
            "e2e_eager_ms_per_step": eager_e2e_ms, "e2e_graph_ms_per_step": graph_e2e_ms,
            "e2e_eager_images_per_s": batch * 1000.0 / eager_e2e_ms, "e2e_graph_images_per_s": batch * 1000.0 / graph_e2e_ms,
            "how": "resident: CUDA events around `steps` walks on the recorder's stream; e2e: wall clock, pinned host Mat in -> pinned host Mat out, "
                   "one stream sync per step, strictly serial"}


def launch_bound_leg(local_rank, storage):
    """BASELINE.json configs[0]: SqueezeNet v1.1 227x227 batch 1 (the reference's own CPU-runnable case), eager and graph replay"""
    from ncnn_b200 import runner
    model, batch, size = WORKLOADS["squeezenet_v1_1"]
    text = with_input_size(modelzoo.param_text(model), size)
    weights = modelzoo.random_model_bytes(text, seed=WEIGHT_SEED)
    s3 = runner.Session(text, weights, storage=storage, device=local_rank)
    del weights
    try:
        x = np.random.default_rng(13).uniform(-1, 1, (batch, 3, size, size)).astype(np.float32)
        hin = s3.pinned_input(x)
        din = s3.upload(hin)
        d = graph_leg(s3, hin, din, batch, 300)
        d["workload"] = "squeezenet_v1_1 %dx%d batch %d, %s storage" % (size, size, batch, storage)
        s3.L.lib.ncnn_cuda_mat_destroy(din)
        s3.L.lib.ncnn_mat_destroy(hin)
        return d
    finally:
        s3.close()


def depthwise_leg(local_rank, storage, group, max_over_ranks, steps, peaks):
    """MobileNetV2 224x224 batch 128 (BASELINE.json configs[1]): step time with the input resident in HBM, the depthwise layers'
    share of it from the per-layer events, their algorithmic bytes (in + out + filters) against the measured HBM copy peak"""
    from ncnn_b200 import runner
    model, batch, size = WORKLOADS["mobilenet_v2"]
    text = with_input_size(modelzoo.param_text(model), size)
    weights = modelzoo.random_model_bytes(text, seed=WEIGHT_SEED)
    s2 = runner.Session(text, weights, storage=storage, device=local_rank)
    del weights
    try:
        lib = s2.L.lib
        x = np.random.default_rng(11).uniform(-1, 1, (batch, 3, size, size)).astype(np.float32)
        hin = s2.pinned_input(x)
        din = s2.upload(hin)
        for _ in range(5):
            lib.ncnn_cuda_mat_destroy(s2.enqueue_device(din))

        def bar():
            s2.sync()
            group.barrier()
        n = max(steps, 20)
        ms = max_over_ranks(timed_steps(s2, lib, din, n, bar))
        bar()
        step_ms = ms / n
        rows, acc = layer_shares(s2, text, din, storage, repeats=5)
        sc = min(1.0, step_ms / acc["total_ms"]) if acc["total_ms"] > 0 else 1.0
        dw_ms = acc["dw_ms"] * sc
        gbs = acc["dw_bytes"] / (dw_ms * 1e-3) / 1e9
        worst = min((r["gbs"] for r in rows if r.get("gbs")), default=None)
        lib.ncnn_cuda_mat_destroy(din)
        lib.ncnn_mat_destroy(hin)
        return {"workload": "mobilenet_v2 224x224 batch 128 per GPU, %s storage" % storage, "images_per_s": batch * n * group.world / (ms * 1e-3), "ms_per_step": step_ms,
                "achieved_gbs": gbs, "frac_of_hbm": gbs / peaks["hbm"], "peak_gbs": peaks["hbm"], "dw_ms_per_step": dw_ms, "dw_layers": sum(1 for r in rows if r.get("gbs")),
                "algorithmic_bytes_per_step": acc["dw_bytes"], "share_of_step": dw_ms / step_ms, "worst_layer_gbs_profiled": worst,
                "how": "depthwise layers' algorithmic bytes (in + out at 2 bytes + fp32 filters) / their share of the un-profiled step (per-layer CUDA events, scaled by %.3f)" % sc}
    finally:
        s2.close()


if __name__ == "__main__":
    sys.exit(main())
