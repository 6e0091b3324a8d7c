"""ncnn_b200/capi.py -- Python (ctypes) binding of the C API in include/c_api.h.

The entry points have the reference's names and signatures (src/c_api.h), so the same binding drives
libncnn_b200.so (the product) and, in the tests, the oracle build of the reference itself.  This is the
ctypes stub INTEGRATION.md shows for Python users; host compute stays in the C++ runtime behind it.

numpy conventions (ncnn Mat <-> ndarray, fp32):
    dims 1: (w,)   dims 2: (h, w)   dims 3: (c, h, w)   dims 4: (c, d, h, w)
    with a batch (Mat::n, src/mat.h:373-381) a leading n axis is added and `batched=True` is passed.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libncnn_b200.so")


class _DataReader(C.Structure):
    pass


_SCAN_FN = C.CFUNCTYPE(C.c_int, C.POINTER(_DataReader), C.c_char_p, C.c_void_p)
_READ_FN = C.CFUNCTYPE(C.c_size_t, C.POINTER(_DataReader), C.c_void_p, C.c_size_t)
_DataReader._fields_ = [("pthis", C.c_void_p), ("scan", _SCAN_FN), ("read", _READ_FN)]


class _Layer(C.Structure):
    pass


_LP = C.POINTER(_Layer)
_Layer._fields_ = [
    ("pthis", C.c_void_p),
    ("load_param", C.CFUNCTYPE(C.c_int, _LP, C.c_void_p)),
    ("load_model", C.CFUNCTYPE(C.c_int, _LP, C.c_void_p)),
    ("create_pipeline", C.CFUNCTYPE(C.c_int, _LP, C.c_void_p)),
    ("destroy_pipeline", C.CFUNCTYPE(C.c_int, _LP, C.c_void_p)),
    ("forward_1", C.CFUNCTYPE(C.c_int, _LP, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p)),
    ("forward_n", C.CFUNCTYPE(C.c_int, _LP, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p)),
    ("forward_inplace_1", C.CFUNCTYPE(C.c_int, _LP, C.c_void_p, C.c_void_p)),
    ("forward_inplace_n", C.CFUNCTYPE(C.c_int, _LP, C.POINTER(C.c_void_p), C.c_int, C.c_void_p)),
]


class NcnnCApi(object):
    """Typed handle on a shared library that exports the reference's C API (src/c_api.h).

    Used for the oracle (oracle/_ref/libncnn_ref_*.so); tests/ re-use it for the product library, which
    exports the same entry points for the hot path (include/ncnn_b200_c_api.h)."""

    def __init__(self, path):
        self.path = path
        self.lib = C.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2))
        L = self.lib
        vp, ci, sz = C.c_void_p, C.c_int, C.c_size_t

        def sig(name, res, *args):
            f = getattr(L, name)
            f.restype = res
            f.argtypes = list(args)
            return f

        sig("ncnn_version", C.c_char_p)
        sig("ncnn_option_create", vp)
        sig("ncnn_option_destroy", None, vp)
        sig("ncnn_option_set_num_threads", None, vp, ci)
        for k in ("use_vulkan_compute", "use_local_pool_allocator", "use_winograd_convolution", "use_sgemm_convolution", "use_packing_layout",
                  "use_fp16_packed", "use_fp16_storage", "use_fp16_arithmetic", "use_int8_packed", "use_int8_storage", "use_int8_arithmetic",
                  "use_bf16_packed", "use_bf16_storage"):
            sig("ncnn_option_set_" + k, None, vp, ci)
        sig("ncnn_mat_create", vp)
        sig("ncnn_mat_create_1d", vp, ci, vp)
        sig("ncnn_mat_create_2d", vp, ci, ci, vp)
        sig("ncnn_mat_create_3d", vp, ci, ci, ci, vp)
        sig("ncnn_mat_create_4d", vp, ci, ci, ci, ci, vp)
        sig("ncnn_mat_create_1d_batch", vp, ci, ci, vp)
        sig("ncnn_mat_create_2d_batch", vp, ci, ci, ci, vp)
        sig("ncnn_mat_create_3d_batch", vp, ci, ci, ci, ci, vp)
        sig("ncnn_mat_create_4d_batch", vp, ci, ci, ci, ci, ci, vp)
        sig("ncnn_mat_destroy", None, vp)
        for k in ("dims", "w", "h", "d", "c", "n", "elempack"):
            sig("ncnn_mat_get_" + k, ci, vp)
        for k in ("elemsize", "cstep", "nstep"):
            sig("ncnn_mat_get_" + k, sz, vp)
        sig("ncnn_mat_get_data", vp, vp)
        sig("ncnn_paramdict_create", vp)
        sig("ncnn_paramdict_destroy", None, vp)
        sig("ncnn_paramdict_set_int", None, vp, ci, ci)
        sig("ncnn_paramdict_set_float", None, vp, ci, C.c_float)
        sig("ncnn_paramdict_set_array", None, vp, ci, vp)
        sig("ncnn_modelbin_create_from_mat_array", vp, C.POINTER(vp), ci)
        sig("ncnn_modelbin_destroy", None, vp)
        sig("ncnn_layer_create_by_type", _LP, C.c_char_p)
        sig("ncnn_layer_destroy", None, _LP)
        sig("ncnn_layer_get_one_blob_only", ci, _LP)
        sig("ncnn_layer_get_support_inplace", ci, _LP)
        sig("ncnn_net_create", vp)
        sig("ncnn_net_destroy", None, vp)
        sig("ncnn_net_set_option", None, vp, vp)
        sig("ncnn_net_load_param_memory", ci, vp, C.c_char_p)
        sig("ncnn_net_load_model_datareader", ci, vp, C.POINTER(_DataReader))
        sig("ncnn_net_load_model", ci, vp, C.c_char_p)
        sig("ncnn_net_load_param", ci, vp, C.c_char_p)
        sig("ncnn_net_get_input_count", ci, vp)
        sig("ncnn_net_get_output_count", ci, vp)
        sig("ncnn_net_get_input_name", C.c_char_p, vp, ci)
        sig("ncnn_net_get_output_name", C.c_char_p, vp, ci)
        sig("ncnn_datareader_create", C.POINTER(_DataReader))
        sig("ncnn_datareader_destroy", None, C.POINTER(_DataReader))
        sig("ncnn_extractor_create", vp, vp)
        sig("ncnn_extractor_destroy", None, vp)
        sig("ncnn_extractor_input", ci, vp, C.c_char_p, vp)
        sig("ncnn_extractor_extract", ci, vp, C.c_char_p, C.POINTER(vp))

    # ------------------------------------------------------------------ Mat <-> numpy
    def mat_from_numpy(self, a, batched=False):
        L = self.lib
        a = np.ascontiguousarray(a, dtype=np.float32)
        shp = a.shape
        n = 1
        if batched:
            n, shp = shp[0], shp[1:]
        dims = len(shp)
        if dims == 1:
            m = L.ncnn_mat_create_1d_batch(shp[0], n, None) if batched else L.ncnn_mat_create_1d(shp[0], None)
        elif dims == 2:
            m = L.ncnn_mat_create_2d_batch(shp[1], shp[0], n, None) if batched else L.ncnn_mat_create_2d(shp[1], shp[0], None)
        elif dims == 3:
            m = L.ncnn_mat_create_3d_batch(shp[2], shp[1], shp[0], n, None) if batched else L.ncnn_mat_create_3d(shp[2], shp[1], shp[0], None)
        elif dims == 4:
            m = (L.ncnn_mat_create_4d_batch(shp[3], shp[2], shp[1], shp[0], n, None) if batched
                 else L.ncnn_mat_create_4d(shp[3], shp[2], shp[1], shp[0], None))
        else:
            raise ValueError("unsupported rank")
        view = self._view(m, force_batch=batched)
        view[...] = a
        return m

    def _view(self, m, force_batch=None):
        """strided float32 ndarray over the Mat's own memory"""
        L = self.lib
        dims = L.ncnn_mat_get_dims(m)
        w, h, d, c, n = (L.ncnn_mat_get_w(m), L.ncnn_mat_get_h(m), L.ncnn_mat_get_d(m), L.ncnn_mat_get_c(m), L.ncnn_mat_get_n(m))
        cstep, nstep = L.ncnn_mat_get_cstep(m), L.ncnn_mat_get_nstep(m)
        if L.ncnn_mat_get_elemsize(m) != 4 or L.ncnn_mat_get_elempack(m) != 1:
            raise ValueError("only fp32 elempack=1 Mats are viewed")
        n = max(n, 1)
        total = (n - 1) * nstep + (cstep * c if dims >= 3 else w * max(h, 1))
        ptr = L.ncnn_mat_get_data(m)
        if not ptr or dims == 0:
            return np.zeros((0,), np.float32)
        buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(total,))
        if dims == 1:
            shape, strides = (n, w), (nstep, 1)
        elif dims == 2:
            shape, strides = (n, h, w), (nstep, w, 1)
        elif dims == 3:
            shape, strides = (n, c, h, w), (nstep, cstep, w, 1)
        else:
            shape, strides = (n, c, d, h, w), (nstep, cstep, h * w, w, 1)
        v = np.lib.stride_tricks.as_strided(buf, shape=shape, strides=tuple(4 * s for s in strides))
        batched = (L.ncnn_mat_get_n(m) > 1) if force_batch is None else force_batch
        return v if batched else v[0]

    def mat_to_numpy(self, m, force_batch=None):
        return np.array(self._view(m, force_batch=force_batch), dtype=np.float32, copy=True)

    def mat_n(self, m):
        return self.lib.ncnn_mat_get_n(m)

    # ------------------------------------------------------------------ Option
    def make_option(self, num_threads=1, **flags):
        opt = self.lib.ncnn_option_create()
        self.lib.ncnn_option_set_num_threads(opt, num_threads)
        for k, v in flags.items():
            getattr(self.lib, "ncnn_option_set_" + k)(opt, int(v))
        return opt

    def strict_fp32_option(self, num_threads=1, packing=False, winograd=False, sgemm=True):
        """the Option tests/testutil.cpp:1311-1320 gives the ground-truth layer: fp32 everywhere"""
        return self.make_option(num_threads, use_packing_layout=packing, use_fp16_packed=0, use_fp16_storage=0, use_fp16_arithmetic=0, use_bf16_storage=0,
                                use_bf16_packed=0, use_int8_packed=0, use_int8_storage=0, use_int8_arithmetic=0, use_winograd_convolution=winograd,
                                use_sgemm_convolution=sgemm, use_vulkan_compute=0)

    # ------------------------------------------------------------------ ParamDict
    def make_paramdict(self, params):
        L = self.lib
        pd = L.ncnn_paramdict_create()
        keep = []
        for k, v in params.items():
            if isinstance(k, str):
                continue  # test-harness hints such as "_ntop"
            if isinstance(v, (list, tuple, np.ndarray)):
                arr = np.asarray(v)
                if arr.dtype.kind == "f":
                    m = self.mat_from_numpy(arr.astype(np.float32))
                else:
                    m = self.mat_from_numpy(arr.astype(np.int32).view(np.float32))
                L.ncnn_paramdict_set_array(pd, int(k), m)
                keep.append(m)
            elif isinstance(v, float):
                L.ncnn_paramdict_set_float(pd, int(k), v)
            else:
                L.ncnn_paramdict_set_int(pd, int(k), int(v))
        for m in keep:
            L.ncnn_mat_destroy(m)
        return pd


class MemoryReader(object):
    """A ncnn_datareader_t (src/c_api.h:219-235) serving a bytes-like object; works with any library exporting that API."""

    def __init__(self, api, data):
        self.api = api
        self.data = memoryview(data).cast("B")
        self.pos = 0
        self.dr = api.lib.ncnn_datareader_create()

        def _read(dr, buf, size):
            n = min(size, len(self.data) - self.pos)
            if n > 0:
                C.memmove(buf, (C.c_char * n).from_buffer(self.data, self.pos) if not self.data.readonly
                          else bytes(self.data[self.pos:self.pos + n]), n)
            self.pos += n
            return n

        def _scan(dr, fmt, p):
            return 0

        self._read_cb = _READ_FN(_read)
        self._scan_cb = _SCAN_FN(_scan)
        self.dr.contents.read = self._read_cb
        self.dr.contents.scan = self._scan_cb

    def close(self):
        if self.dr:
            self.api.lib.ncnn_datareader_destroy(self.dr)
            self.dr = None


class Net(object):
    """Net + Extractor through the C API (works for the oracle and for the product library alike)."""

    def __init__(self, api, param_text, model_bytes, opt):
        self.api = api
        L = api.lib
        self.net = L.ncnn_net_create()
        L.ncnn_net_set_option(self.net, opt)
        if L.ncnn_net_load_param_memory(self.net, param_text.encode() if isinstance(param_text, str) else param_text) != 0:
            raise RuntimeError("load_param failed")
        rd = MemoryReader(api, model_bytes)
        try:
            if L.ncnn_net_load_model_datareader(self.net, rd.dr) != 0:
                raise RuntimeError("load_model failed")
        finally:
            rd.close()
        self.input_names = [L.ncnn_net_get_input_name(self.net, i).decode() for i in range(L.ncnn_net_get_input_count(self.net))]
        self.output_names = [L.ncnn_net_get_output_name(self.net, i).decode() for i in range(L.ncnn_net_get_output_count(self.net))]

    def run(self, inputs, outputs=None, batched=False):
        """inputs: {blob name: ndarray}; returns {blob name: ndarray} for `outputs` (default: all net outputs)"""
        L = self.api.lib
        ex = L.ncnn_extractor_create(self.net)
        mats = []
        res = {}
        try:
            for k, v in inputs.items():
                m = self.api.mat_from_numpy(v, batched=batched)
                mats.append(m)
                if L.ncnn_extractor_input(ex, k.encode(), m) != 0:
                    raise RuntimeError("input %s failed" % k)
            for name in (outputs or self.output_names):
                out = C.c_void_p()
                r = L.ncnn_extractor_extract(ex, name.encode(), C.byref(out))
                if r != 0:
                    raise RuntimeError("extract %s returned %d" % (name, r))
                res[name] = self.api.mat_to_numpy(out, force_batch=batched)
                L.ncnn_mat_destroy(out)
        finally:
            L.ncnn_extractor_destroy(ex)
            for m in mats:
                L.ncnn_mat_destroy(m)
        return res

    def run_pixels(self, name, pixels, pixel_type, mean_vals=None, norm_vals=None, outputs=None, resize=None):
        """pixels: (n, h, w, channels) uint8, interleaved; pre-processing (from_pixels + substract_mean_normalize) runs on the
        device (ncnn_extractor_input_pixels, product library only); returns {blob name: ndarray (n, ...)}"""
        import numpy as np
        L = self.api.lib
        L.ncnn_extractor_input_pixels.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
        pixels = np.ascontiguousarray(pixels, np.uint8)
        n, h, w, ch = pixels.shape
        mean = np.asarray(mean_vals, np.float32) if mean_vals is not None else None
        norm = np.asarray(norm_vals, np.float32) if norm_vals is not None else None
        ex = L.ncnn_extractor_create(self.net)
        res = {}
        try:
            mp = mean.ctypes.data_as(C.c_void_p) if mean is not None else None
            npp = norm.ctypes.data_as(C.c_void_p) if norm is not None else None
            if resize is not None:  # (target_w, target_h): the reference's from_pixels_resize, on the device
                L.ncnn_extractor_input_pixels_resize.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                                                 C.c_void_p, C.c_void_p]
                r = L.ncnn_extractor_input_pixels_resize(ex, name.encode(), pixels.ctypes.data_as(C.c_void_p), pixel_type, w, h, w * ch, n, h * w * ch, resize[0], resize[1], mp, npp)
            else:
                r = L.ncnn_extractor_input_pixels(ex, name.encode(), pixels.ctypes.data_as(C.c_void_p), pixel_type, w, h, w * ch, n, h * w * ch, mp, npp)
            if r != 0:
                raise RuntimeError("input_pixels %s returned %d" % (name, r))
            for oname in (outputs or self.output_names):
                out = C.c_void_p()
                r = L.ncnn_extractor_extract(ex, oname.encode(), C.byref(out))
                if r != 0:
                    raise RuntimeError("extract %s returned %d" % (oname, r))
                res[oname] = self.api.mat_to_numpy(out, force_batch=True)
                L.ncnn_mat_destroy(out)
        finally:
            L.ncnn_extractor_destroy(ex)
        return res

    def close(self):
        if self.net:
            self.api.lib.ncnn_net_destroy(self.net)
            self.net = None


_lib = None


def library():
    """the product library; fails loudly when it has not been built (there is no fallback)"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python -m ncnn_b200.build` (or __graft_entry__.build())" % LIB_PATH)
        _lib = NcnnCApi(LIB_PATH)
    return _lib
