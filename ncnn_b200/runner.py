"""ncnn_b200/runner.py -- small Python conveniences over the C API binding (ncnn_b200/capi.py) used by bench.py,
__graft_entry__.smoke() and the tests: load a Net, run it with host Mats (the reference-facing path) or with
device-resident blobs, time it with CUDA events on the recorder's own stream, read the per-layer profile.
All compute goes through libncnn_b200.so; nothing here falls back to the CPU."""
import ctypes as C

import numpy as np

from . import capi


def _bind_extras(L):
    lib = L.lib
    if getattr(lib, "_b200_extras_bound", False):
        return
    vp, ci = C.c_void_p, C.c_int
    lib.ncnn_cuda_compute_create.restype = vp
    lib.ncnn_cuda_compute_create.argtypes = [ci]
    lib.ncnn_cuda_compute_destroy.argtypes = [vp]
    lib.ncnn_cuda_compute_get_stream.restype = vp
    lib.ncnn_cuda_compute_get_stream.argtypes = [vp]
    lib.ncnn_cuda_compute_record_upload.argtypes = [vp, vp, C.POINTER(vp), vp]
    lib.ncnn_cuda_compute_record_download.argtypes = [vp, vp, C.POINTER(vp), vp]
    lib.ncnn_cuda_compute_submit_and_wait.argtypes = [vp]
    lib.ncnn_cuda_compute_set_profiling.argtypes = [vp, ci]
    lib.ncnn_cuda_compute_get_profile_count.argtypes = [vp]
    lib.ncnn_cuda_compute_get_profile.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(C.c_float), C.POINTER(ci)]
    lib.ncnn_cuda_compute_clear_profile.argtypes = [vp]
    lib.ncnn_cuda_mat_destroy.argtypes = [vp]
    lib.ncnn_extractor_input_cuda.argtypes = [vp, C.c_char_p, vp]
    lib.ncnn_extractor_extract_cuda.argtypes = [vp, C.c_char_p, C.POINTER(vp), vp]
    lib.ncnn_extractor_get_last_h2d_bytes.restype = C.c_size_t
    lib.ncnn_extractor_get_last_h2d_bytes.argtypes = [vp]
    lib.ncnn_extractor_get_last_d2h_bytes.restype = C.c_size_t
    lib.ncnn_extractor_get_last_d2h_bytes.argtypes = [vp]
    lib.ncnn_allocator_create_cuda_staging_allocator.restype = vp
    lib.ncnn_allocator_destroy.argtypes = [vp]
    lib.ncnn_option_set_blob_allocator.argtypes = [vp, vp]
    lib.ncnn_option_set_use_cuda_graph_fusion.argtypes = [vp, ci]
    lib.ncnn_option_set_lightmode.argtypes = [vp, ci]
    lib.ncnn_net_set_cuda_device.argtypes = [vp, ci]
    lib.ncnn_net_get_layer_count.argtypes = [vp]
    lib.ncnn_net_get_layer_type.restype = C.c_char_p
    lib.ncnn_net_get_layer_type.argtypes = [vp, ci]
    lib.ncnn_net_get_layer_name.restype = C.c_char_p
    lib.ncnn_net_get_layer_name.argtypes = [vp, ci]
    lib.ncnn_net_get_fused_layer_count.argtypes = [vp]
    lib.ncnn_get_cuda_device_count.restype = ci
    lib.ncnn_cuda_set_device.argtypes = [ci]
    lib.ncnn_cuda_launch_count.restype = C.c_ulonglong
    lib.ncnn_cuda_last_error.restype = C.c_char_p
    for f in ("ncnn_cuda_event_create",):
        getattr(lib, f).argtypes = [C.POINTER(vp)]
    lib.ncnn_cuda_event_destroy.argtypes = [vp]
    lib.ncnn_cuda_event_record.argtypes = [vp, vp]
    lib.ncnn_cuda_event_sync.argtypes = [vp]
    lib.ncnn_cuda_event_elapsed_ms.argtypes = [vp, vp, C.POINTER(C.c_float)]
    lib.ncnn_cuda_device_sync.argtypes = []
    lib.ncnn_cuda_graph_begin_capture.argtypes = [vp]
    lib.ncnn_cuda_graph_end_capture.argtypes = [vp, C.POINTER(vp)]
    lib.ncnn_cuda_graph_launch.argtypes = [vp, vp]
    lib.ncnn_cuda_graph_destroy.argtypes = [vp]
    lib._b200_extras_bound = True


STORAGE = {
    "fp32": dict(use_fp16_storage=0, use_fp16_packed=0, use_fp16_arithmetic=0, use_bf16_storage=0),
    "fp16": dict(use_fp16_storage=1, use_bf16_storage=0),
    "bf16": dict(use_fp16_storage=0, use_bf16_storage=1),
}


class Session(object):
    def __init__(self, param_text, model_bytes, storage="fp16", device=0, fusion=True):
        self.L = capi.library()
        _bind_extras(self.L)
        lib = self.L.lib
        if lib.ncnn_get_cuda_device_count() <= 0:
            raise RuntimeError("no CUDA device: %s" % lib.ncnn_cuda_last_error().decode())
        lib.ncnn_cuda_set_device(device)
        self.device = device
        self.opt = self.L.make_option(1, **STORAGE[storage])
        lib.ncnn_option_set_use_cuda_graph_fusion(self.opt, 1 if fusion else 0)
        self.staging = lib.ncnn_allocator_create_cuda_staging_allocator()
        lib.ncnn_option_set_blob_allocator(self.opt, self.staging)  # extracted Mats land in pinned memory
        self.net = lib.ncnn_net_create()
        lib.ncnn_net_set_option(self.net, self.opt)
        lib.ncnn_net_set_cuda_device(self.net, device)
        if lib.ncnn_net_load_param_memory(self.net, param_text.encode()) != 0:
            raise RuntimeError("load_param failed")
        rd = capi.MemoryReader(self.L, model_bytes)
        try:
            if lib.ncnn_net_load_model_datareader(self.net, rd.dr) != 0:
                raise RuntimeError("load_model failed: %s" % lib.ncnn_cuda_last_error().decode())
        finally:
            rd.close()
        self.input_name = lib.ncnn_net_get_input_name(self.net, 0)
        self.output_name = lib.ncnn_net_get_output_name(self.net, 0)
        self.cmd = lib.ncnn_cuda_compute_create(device)
        self.stream = lib.ncnn_cuda_compute_get_stream(self.cmd)
        self.layer_types = [lib.ncnn_net_get_layer_type(self.net, i).decode() for i in range(lib.ncnn_net_get_layer_count(self.net))]
        self.layer_names = [lib.ncnn_net_get_layer_name(self.net, i).decode() for i in range(lib.ncnn_net_get_layer_count(self.net))]
        self.fused_layers = lib.ncnn_net_get_fused_layer_count(self.net)

    # ---------------------------------------------------------------- host path (what a user of the reference calls)
    def pinned_input(self, x):
        """x: (n, c, h, w) float32 -> a batched ncnn Mat in page-locked memory"""
        lib = self.L.lib
        n, c, h, w = x.shape
        lib.ncnn_mat_create_3d_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        m = lib.ncnn_mat_create_3d_batch(w, h, c, n, self.staging)
        self.L._view(m, force_batch=True)[...] = x
        return m

    def extract_host(self, mat, blob=None):
        """one reference-style call: Extractor.input(host Mat) + extract(host Mat); H2D and D2H inside.
        blob: name of the blob to extract (default: the graph's first output)"""
        lib = self.L.lib
        ex = lib.ncnn_extractor_create(self.net)
        out = C.c_void_p()
        try:
            if lib.ncnn_extractor_input(ex, self.input_name, mat) != 0:
                raise RuntimeError("input failed")
            r = lib.ncnn_extractor_extract(ex, blob.encode() if isinstance(blob, str) else (blob or self.output_name), C.byref(out))
            if r != 0:
                raise RuntimeError("extract returned %d: %s" % (r, lib.ncnn_cuda_last_error().decode()))
            self.last_h2d = lib.ncnn_extractor_get_last_h2d_bytes(ex)
            self.last_d2h = lib.ncnn_extractor_get_last_d2h_bytes(ex)
        finally:
            lib.ncnn_extractor_destroy(ex)
        return out

    def pinned_pixels(self, pixels):
        """pixels: (n, h, w, ch) uint8 -> (pinned Mat that owns the bytes, ctypes pointer, shape)"""
        lib = self.L.lib
        pixels = np.ascontiguousarray(pixels, np.uint8)
        nbytes = pixels.size
        lib.ncnn_mat_create_1d.restype = C.c_void_p
        lib.ncnn_mat_create_1d.argtypes = [C.c_int, C.c_void_p]
        lib.ncnn_mat_get_data.restype = C.c_void_p
        lib.ncnn_mat_get_data.argtypes = [C.c_void_p]
        m = lib.ncnn_mat_create_1d((nbytes + 3) // 4, self.staging)
        ptr = lib.ncnn_mat_get_data(m)
        C.memmove(ptr, pixels.ctypes.data, nbytes)
        return m, ptr, pixels.shape

    def extract_host_pixels(self, ptr, shape, pixel_type, mean_vals, norm_vals, yolov8_decode=None):
        """one call with device pre-processing: Extractor.input_pixels(pinned 8-bit images) + extract(host Mat).
        yolov8_decode = (strides, prob_threshold): the result is decoded on the device as well
        (ncnn_extractor_extract_yolov8_proposals) and only 6 floats per anchor come back"""
        lib = self.L.lib
        lib.ncnn_extractor_extract_yolov8_proposals.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
        lib.ncnn_extractor_get_last_d2h_bytes.restype = C.c_size_t
        lib.ncnn_extractor_input_pixels.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
        n, h, w, ch = shape
        ex = lib.ncnn_extractor_create(self.net)
        out = C.c_void_p()
        try:
            r = lib.ncnn_extractor_input_pixels(ex, self.input_name, C.c_void_p(ptr), pixel_type, w, h, w * ch, n, h * w * ch,
                                                mean_vals.ctypes.data_as(C.c_void_p) if mean_vals is not None else None,
                                                norm_vals.ctypes.data_as(C.c_void_p) if norm_vals is not None else None)
            if r != 0:
                raise RuntimeError("input_pixels returned %d" % r)
            if yolov8_decode is not None:
                strides, thr = yolov8_decode
                st = (C.c_int * len(strides))(*strides)
                r = lib.ncnn_extractor_extract_yolov8_proposals(ex, self.output_name, st, len(strides), w, h, thr, C.byref(out))
            else:
                r = lib.ncnn_extractor_extract(ex, self.output_name, C.byref(out))
            if r != 0:
                raise RuntimeError("extract returned %d: %s" % (r, lib.ncnn_cuda_last_error().decode()))
            self.last_h2d_pixels = lib.ncnn_extractor_get_last_h2d_bytes(ex)
            self.last_d2h_pixels = lib.ncnn_extractor_get_last_d2h_bytes(ex)
        finally:
            lib.ncnn_extractor_destroy(ex)
        return out

    def run_host(self, x, blob=None):
        m = self.pinned_input(x)
        out = self.extract_host(m, blob)
        res = self.L.mat_to_numpy(out, force_batch=True)
        self.L.lib.ncnn_mat_destroy(out)
        self.L.lib.ncnn_mat_destroy(m)
        return res

    # ---------------------------------------------------------------- device-resident path
    def upload(self, mat):
        lib = self.L.lib
        dm = C.c_void_p()
        r = lib.ncnn_cuda_compute_record_upload(self.cmd, mat, C.byref(dm), self.opt)
        if r != 0:
            raise RuntimeError("record_upload returned %d" % r)
        lib.ncnn_cuda_compute_submit_and_wait(self.cmd)
        return dm

    def enqueue_device(self, dmat):
        """record one forward walk on the session's stream (no sync); returns the device output handle"""
        lib = self.L.lib
        ex = lib.ncnn_extractor_create(self.net)
        out = C.c_void_p()
        try:
            lib.ncnn_extractor_input_cuda(ex, self.input_name, dmat)
            r = lib.ncnn_extractor_extract_cuda(ex, self.output_name, C.byref(out), self.cmd)
            if r != 0:
                raise RuntimeError("extract_cuda returned %d: %s" % (r, lib.ncnn_cuda_last_error().decode()))
        finally:
            lib.ncnn_extractor_destroy(ex)
        return out

    def download(self, dmat):
        lib = self.L.lib
        m = C.c_void_p()
        lib.ncnn_cuda_compute_record_download(self.cmd, dmat, C.byref(m), self.opt)
        lib.ncnn_cuda_compute_submit_and_wait(self.cmd)
        res = self.L.mat_to_numpy(m, force_batch=True)
        lib.ncnn_mat_destroy(m)
        return res

    def sync(self):
        return self.L.lib.ncnn_cuda_compute_submit_and_wait(self.cmd)

    # ---------------------------------------------------------------- CUDA-graph replay of a recorded walk
    def capture(self, host_mat=None, dev_in=None):
        """Record ONE forward walk on the session's stream into a CUDA graph (ncnn_cuda_graph_begin_capture / _end_capture of
        include/ncnn_cuda.h around the ordinary recorder + Extractor calls -- the analogue of re-submitting a recorded VkCompute
        command buffer) and return a Graph whose replay() is a single cudaGraphLaunch.
        host_mat: a PINNED host Mat -> the graph holds H2D copy + layout kernel + walk + layout kernel + D2H copy into a pinned
                  output Mat of fixed address (Graph.host_out): the whole reference-facing extract, replayable after the caller
                  has overwritten host_mat's bytes in place.
        dev_in:   a device blob -> the graph holds the walk only (Graph.dev_out stays resident).
        The walk is run once un-captured first: plans, weight packs and the device pool are built there, so the captured walk
        allocates nothing new (the pool hands back the same blocks; they stay reserved for as long as the Session lives)."""
        lib = self.L.lib
        assert (host_mat is None) != (dev_in is None)
        # warm-up: the same sequence of recorder calls on the same stream / device pool, un-captured
        wd = self.upload(host_mat) if host_mat is not None else None
        out = self.enqueue_device(dev_in if dev_in is not None else wd)
        if host_mat is not None:
            self.download(out)
        self.sync()
        lib.ncnn_cuda_mat_destroy(out)
        if wd is not None:
            lib.ncnn_cuda_mat_destroy(wd)
        g = Graph(self)
        n0 = self.launch_count()
        if lib.ncnn_cuda_graph_begin_capture(self.stream) != 0:
            raise RuntimeError("begin_capture failed: %s" % lib.ncnn_cuda_last_error().decode())
        err = None
        try:
            if host_mat is not None:
                dm = C.c_void_p()
                if lib.ncnn_cuda_compute_record_upload(self.cmd, host_mat, C.byref(dm), self.opt) != 0:
                    raise RuntimeError("record_upload failed under capture")
                g.dev_in = dm
                g.dev_out = self.enqueue_device(dm)
                m = C.c_void_p()
                if lib.ncnn_cuda_compute_record_download(self.cmd, g.dev_out, C.byref(m), self.opt) != 0:
                    raise RuntimeError("record_download failed under capture")
                g.host_out = m
            else:
                g.dev_out = self.enqueue_device(dev_in)
        except Exception as e:  # the stream must leave capture mode whatever happened
            err = e
        ge = C.c_void_p()
        r = lib.ncnn_cuda_graph_end_capture(self.stream, C.byref(ge))
        g.kernels = self.launch_count() - n0
        self.sync()  # nothing was enqueued (capture only records); hands the recorder's scratch blobs back to the pool
        if err is not None or r != 0 or not ge.value:
            g.close()
            raise RuntimeError("graph capture failed: %s" % (err if err is not None else lib.ncnn_cuda_last_error().decode()))
        g.exec = ge
        return g

    # ---------------------------------------------------------------- timing helpers
    def event(self):
        e = C.c_void_p()
        self.L.lib.ncnn_cuda_event_create(C.byref(e))
        return e

    def record(self, e):
        self.L.lib.ncnn_cuda_event_record(e, self.stream)

    def elapsed_ms(self, e0, e1):
        ms = C.c_float()
        self.L.lib.ncnn_cuda_event_sync(e1)
        self.L.lib.ncnn_cuda_event_elapsed_ms(e0, e1, C.byref(ms))
        return ms.value

    def launch_count(self):
        return int(self.L.lib.ncnn_cuda_launch_count())

    def profile(self, dmat, repeats=1):
        """per-layer device time: [(layer index, type, name, ms, (dims, w, h, d, c, n))] averaged over `repeats` walks"""
        lib = self.L.lib
        lib.ncnn_cuda_compute_clear_profile(self.cmd)
        lib.ncnn_cuda_compute_set_profiling(self.cmd, 1)
        for _ in range(repeats):
            out = self.enqueue_device(dmat)
            lib.ncnn_cuda_mat_destroy(out)
        lib.ncnn_cuda_compute_submit_and_wait(self.cmd)
        lib.ncnn_cuda_compute_set_profiling(self.cmd, 0)
        n = lib.ncnn_cuda_compute_get_profile_count(self.cmd)
        acc = {}
        order = []
        for i in range(n):
            li, ms, shape = C.c_int(), C.c_float(), (C.c_int * 6)()
            lib.ncnn_cuda_compute_get_profile(self.cmd, i, C.byref(li), C.byref(ms), shape)
            if li.value not in acc:
                acc[li.value] = [0.0, tuple(shape)]
                order.append(li.value)
            acc[li.value][0] += ms.value
        lib.ncnn_cuda_compute_clear_profile(self.cmd)
        return [(li, self.layer_types[li], self.layer_names[li], acc[li][0] / repeats, acc[li][1]) for li in order]

    def close(self):
        lib = self.L.lib
        if self.cmd:
            lib.ncnn_cuda_compute_destroy(self.cmd)
            self.cmd = None
        if self.net:
            lib.ncnn_net_destroy(self.net)
            self.net = None


class Graph(object):
    """an instantiated CUDA graph of one recorded walk (Session.capture); replay() enqueues it on the session's stream"""

    def __init__(self, sess):
        self.sess = sess
        self.exec = None
        self.dev_in = None    # owned only in the host_mat form
        self.dev_out = None
        self.host_out = None  # pinned output Mat (host_mat form): valid after replay() + Session.sync()
        self.kernels = 0      # kernels recorded into the graph

    def replay(self):
        if self.sess.L.lib.ncnn_cuda_graph_launch(self.exec, self.sess.stream) != 0:
            raise RuntimeError("graph launch failed: %s" % self.sess.L.lib.ncnn_cuda_last_error().decode())

    def result(self):
        """host_mat form: wait for the stream and read the pinned output Mat"""
        self.sess.sync()
        return self.sess.L.mat_to_numpy(self.host_out, force_batch=True).copy()

    def close(self):
        lib = self.sess.L.lib
        if self.exec:
            lib.ncnn_cuda_graph_destroy(self.exec)
            self.exec = None
        if self.host_out:
            lib.ncnn_mat_destroy(self.host_out)
            self.host_out = None
        if self.dev_out:
            lib.ncnn_cuda_mat_destroy(self.dev_out)
            self.dev_out = None
        if self.dev_in:
            lib.ncnn_cuda_mat_destroy(self.dev_in)
            self.dev_in = None


def layer_work(param_text, profile):
    """algorithmic work of each profiled layer from the graph's parameters and the measured top shapes:
    -> {layer index: dict(macs=..., bytes=...)} for Convolution / ConvolutionDepthWise / InnerProduct / Pooling"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import modelzoo
    layers = {l[1]: l for l in modelzoo.parse_param(param_text)}
    out = {}
    for li, t, name, ms, shape in profile:
        if name not in layers:
            continue
        p = layers[name][4]
        dims, w, h, d, c, n = shape
        n = max(n, 1)
        if t == "Convolution" or t == "ConvolutionDepthWise":
            per_out = p[6] // p[0]  # weight_data_size / num_output = inch_per_group * kw * kh
            out[li] = dict(macs=per_out * w * h * c * n, weights=p[6], out_elems=w * h * c * n, kind=t, k=p.get(1, 1), s=p.get(3, 1))
        elif t == "InnerProduct":
            out[li] = dict(macs=p[2] * n, weights=p[2], out_elems=p[0] * n, kind=t)
    return out
