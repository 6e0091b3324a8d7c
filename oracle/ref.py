"""oracle/ref.py -- ctypes driver for the UNMODIFIED reference (Tencent/ncnn CPU path) built into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, as the checker and the CPU baseline; never by the product (ncnn_b200/).

The library is the reference's own sources compiled by oracle/build_ref.py; this module only calls the
reference's public C API (src/c_api.h) plus the few accessors in oracle/ref_driver.cpp.

numpy conventions (ncnn Mat <-> ndarray, fp32):
    dims 1: (w,)   dims 2: (h, w)   dims 3: (c, h, w)   dims 4: (c, d, h, w)
    with a batch (Mat::n, src/mat.h:373-381) a leading n axis is added and `batched=True` is passed.
"""
import ctypes as C
import os

import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

HERE = os.path.dirname(os.path.abspath(__file__))


def _cpu_flags():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def lib_path():
    flags = _cpu_flags()
    order = []
    if {"avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512cd"} <= flags:
        order.append("avx512")
    order.append("avx2")
    for isa in order:
        p = os.path.join(HERE, "_ref", "libncnn_ref_%s.so" % isa)
        if os.path.exists(p):
            return p
    return None


def available():
    return lib_path() is not None


from ncnn_b200.capi import NcnnCApi, MemoryReader, Net, _LP  # noqa: F401  (generic binding of src/c_api.h)


class Reference(NcnnCApi):
    """The oracle: the reference's own CPU implementation."""

    def __init__(self, path=None):
        path = path or lib_path()
        if path is None:
            raise RuntimeError("oracle/_ref/libncnn_ref_*.so not built: run `python oracle/build_ref.py` where /root/reference exists")
        NcnnCApi.__init__(self, path)
        L = self.lib
        for name in ("ref_layer_create_naive", "ref_layer_create_cpu"):
            f = getattr(L, name)
            f.restype = _LP
            f.argtypes = [C.c_char_p]
        L.ref_layer_load_model_from_mats.restype = C.c_int
        L.ref_layer_load_model_from_mats.argtypes = [_LP, C.POINTER(C.c_void_p), C.c_int]
        L.ref_option_set_lightmode.argtypes = [C.c_void_p, C.c_int]
        L.ref_option_set_flush_denormals.argtypes = [C.c_void_p, C.c_int]
        L.ref_cpu_count.restype = C.c_int
        L.ref_physical_big_cpu_count.restype = C.c_int
        L.ref_set_omp_num_threads.argtypes = [C.c_int]

    def cpu_count(self):
        return self.lib.ref_cpu_count()

    def layer_forward(self, type_name, params, weights, bottoms, batched=False, naive=True, opt=None):
        """Run one reference layer the way tests/testutil.cpp:1301-1339 runs its ground truth:
        create_layer_naive, load_param from a ParamDict, load_model from a ModelBinFromMatArray, fp32 Option.
        With `batched`, bottoms carry a leading n axis and the layer is run per sample (the reference's own
        batch semantics, src/net.cpp:654-705), results stacked."""
        L = self.lib
        own_opt = opt is None
        if own_opt:
            opt = self.strict_fp32_option()
        creator = L.ref_layer_create_naive if naive else L.ref_layer_create_cpu
        layer = creator(type_name.encode())
        if not layer:
            raise RuntimeError("reference has no layer " + type_name)
        pd = self.make_paramdict(params)
        wmats = [self.mat_from_numpy(np.asarray(w, np.float32).reshape(-1)) for w in weights]
        arr = (C.c_void_p * max(len(wmats), 1))(*wmats)
        lay = layer.contents
        try:
            if lay.load_param(layer, pd) != 0:
                raise RuntimeError("reference load_param failed")
            if L.ref_layer_load_model_from_mats(layer, arr, len(wmats)) != 0:
                raise RuntimeError("reference load_model failed")
            if lay.create_pipeline(layer, opt) != 0:
                raise RuntimeError("reference create_pipeline failed")
            one_blob = L.ncnn_layer_get_one_blob_only(layer) != 0
            inplace = L.ncnn_layer_get_support_inplace(layer) != 0
            nb = bottoms[0].shape[0] if batched else 1
            per_sample = []
            for b in range(nb):
                ins = [self.mat_from_numpy(x[b] if batched else x) for x in bottoms]
                outs = []
                if one_blob and inplace:
                    r = lay.forward_inplace_1(layer, ins[0], opt)
                    outs = [self.mat_to_numpy(ins[0])]
                elif one_blob:
                    top = C.c_void_p()
                    r = lay.forward_1(layer, ins[0], C.byref(top), opt)
                    if r == 0:
                        outs = [self.mat_to_numpy(top)]
                    L.ncnn_mat_destroy(top)
                elif inplace:
                    a = (C.c_void_p * len(ins))(*ins)
                    r = lay.forward_inplace_n(layer, a, len(ins), opt)
                    outs = [self.mat_to_numpy(m) for m in ins]
                else:
                    ntop = params.get("_ntop", 1)
                    a = (C.c_void_p * len(ins))(*ins)
                    t = (C.c_void_p * ntop)()
                    r = lay.forward_n(layer, a, len(ins), t, ntop, opt)
                    if r == 0:
                        outs = [self.mat_to_numpy(t[i]) for i in range(ntop)]
                    for i in range(ntop):
                        if t[i]:
                            L.ncnn_mat_destroy(t[i])
                for m in ins:
                    L.ncnn_mat_destroy(m)
                if r != 0:
                    raise RuntimeError("reference forward returned %d" % r)
                per_sample.append(outs)
            lay.destroy_pipeline(layer, opt)
        finally:
            L.ncnn_layer_destroy(layer)
            L.ncnn_paramdict_destroy(pd)
            for m in wmats:
                L.ncnn_mat_destroy(m)
            if own_opt:
                L.ncnn_option_destroy(opt)
        if batched:
            return [np.stack([s[i] for s in per_sample]) for i in range(len(per_sample[0]))]
        return per_sample[0]


_ref = None


def reference():
    global _ref
    if _ref is None:
        _ref = Reference()
    return _ref
