"""ctypes view of include/ncnn_cuda.h (the kernel C ABI of ncnn_b200/libncnn_b200.so) for the parity tests.

Device memory comes from torch (plumbing only); every compute call goes through the C ABI.
A device blob is channel-innermost: [n][P][cpitch] (see the header)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "ncnn_b200", "libncnn_b200.so")

F32, BF16, F16 = 0, 1, 2


class Tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dims", C.c_int), ("w", C.c_int), ("h", C.c_int), ("d", C.c_int), ("c", C.c_int), ("n", C.c_int),
                ("elemtype", C.c_int), ("cpitch", C.c_int), ("nstep", C.c_longlong)]


class HostMat(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dims", C.c_int), ("w", C.c_int), ("h", C.c_int), ("d", C.c_int), ("c", C.c_int), ("n", C.c_int),
                ("cstep", C.c_longlong), ("nstep", C.c_longlong)]


class Activation(C.Structure):
    _fields_ = [("type", C.c_int), ("p0", C.c_float), ("p1", C.c_float)]


class ConvDesc(C.Structure):
    _fields_ = [("inch", C.c_int), ("outch", C.c_int), ("kernel_w", C.c_int), ("kernel_h", C.c_int), ("dilation_w", C.c_int), ("dilation_h", C.c_int),
                ("stride_w", C.c_int), ("stride_h", C.c_int), ("pad_left", C.c_int), ("pad_right", C.c_int), ("pad_top", C.c_int), ("pad_bottom", C.c_int),
                ("pad_value", C.c_float), ("bias_term", C.c_int), ("act", Activation), ("elemtype", C.c_int)]


class DwConvDesc(C.Structure):
    _fields_ = [("inch", C.c_int), ("outch", C.c_int), ("group", C.c_int), ("kernel_w", C.c_int), ("kernel_h", C.c_int), ("dilation_w", C.c_int),
                ("dilation_h", C.c_int), ("stride_w", C.c_int), ("stride_h", C.c_int), ("pad_value", C.c_float), ("bias_term", C.c_int), ("act", Activation),
                ("elemtype", C.c_int)]


class DeconvDesc(C.Structure):
    _fields_ = [("inch", C.c_int), ("outch", C.c_int), ("group", C.c_int), ("kernel_w", C.c_int), ("kernel_h", C.c_int), ("dilation_w", C.c_int),
                ("dilation_h", C.c_int), ("stride_w", C.c_int), ("stride_h", C.c_int), ("output_pad_right", C.c_int), ("output_pad_bottom", C.c_int),
                ("bias_term", C.c_int), ("act", Activation), ("elemtype", C.c_int)]


class PoolDesc(C.Structure):
    _fields_ = [("pooling_type", C.c_int), ("kernel_w", C.c_int), ("kernel_h", C.c_int), ("stride_w", C.c_int), ("stride_h", C.c_int), ("pad_left", C.c_int),
                ("pad_top", C.c_int), ("global_pooling", C.c_int), ("avgpool_count_include_pad", C.c_int), ("adaptive_pooling", C.c_int),
                ("area_x0", C.c_int), ("area_x1", C.c_int), ("area_y0", C.c_int), ("area_y1", C.c_int)]


class LinearDesc(C.Structure):
    _fields_ = [("num_input", C.c_int), ("num_output", C.c_int), ("bias_term", C.c_int), ("act", Activation), ("elemtype", C.c_int), ("in_w", C.c_int),
                ("in_h", C.c_int), ("in_c", C.c_int)]


class GemmArgs(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch", C.c_int),
                ("a", C.c_void_p), ("a_rs", C.c_longlong), ("a_cs", C.c_longlong), ("a_bs", C.c_longlong),
                ("b", C.c_void_p), ("b_rs", C.c_longlong), ("b_cs", C.c_longlong), ("b_bs", C.c_longlong),
                ("c", C.c_void_p), ("c_rs", C.c_longlong), ("c_cs", C.c_longlong), ("c_bs", C.c_longlong),
                ("out", C.c_void_p), ("o_rs", C.c_longlong), ("o_cs", C.c_longlong), ("o_bs", C.c_longlong),
                ("alpha", C.c_float), ("beta", C.c_float), ("elemtype", C.c_int), ("c_elemtype", C.c_int)]


# every symbol include/ncnn_cuda.h declares (checked by the CPU-side export test)
KERNEL_ABI_SYMBOLS = [
    "ncnn_cuda_device_count", "ncnn_cuda_set_device", "ncnn_cuda_get_device", "ncnn_cuda_device_info", "ncnn_cuda_last_error",
    "ncnn_cuda_malloc", "ncnn_cuda_free", "ncnn_cuda_malloc_host", "ncnn_cuda_free_host", "ncnn_cuda_host_is_pinned", "ncnn_cuda_memcpy_h2d_async", "ncnn_cuda_memcpy_d2h_async",
    "ncnn_cuda_memcpy_d2d_async", "ncnn_cuda_memset_async", "ncnn_cuda_stream_create", "ncnn_cuda_stream_destroy", "ncnn_cuda_stream_sync",
    "ncnn_cuda_device_sync", "ncnn_cuda_event_create", "ncnn_cuda_event_destroy", "ncnn_cuda_event_record", "ncnn_cuda_event_sync",
    "ncnn_cuda_event_elapsed_ms", "ncnn_cuda_graph_begin_capture", "ncnn_cuda_graph_end_capture", "ncnn_cuda_graph_launch", "ncnn_cuda_graph_destroy",
    "ncnn_cuda_launch_count", "ncnn_cuda_tc_launch_count", "ncnn_cuda_pack_from_planar", "ncnn_cuda_unpack_to_planar", "ncnn_cuda_reshape", "ncnn_cuda_permute",
    "ncnn_cuda_conv2d_create", "ncnn_cuda_conv2d_destroy", "ncnn_cuda_conv2d_forward", "ncnn_cuda_conv2d_workspace_size", "ncnn_cuda_conv2d_algo",
    "ncnn_cuda_dwconv2d_create", "ncnn_cuda_dwconv2d_destroy", "ncnn_cuda_dwconv2d_forward", "ncnn_cuda_pool2d_forward",
    "ncnn_cuda_linear_create", "ncnn_cuda_linear_destroy", "ncnn_cuda_linear_forward", "ncnn_cuda_gemm_strided",
    "ncnn_cuda_unary", "ncnn_cuda_eltwise", "ncnn_cuda_binaryop", "ncnn_cuda_copy_into_axis", "ncnn_cuda_copy_from_axis", "ncnn_cuda_interp",
    "ncnn_cuda_softmax", "ncnn_cuda_padding",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s missing: run `python -m ncnn_b200.build`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.ncnn_cuda_last_error.restype = C.c_char_p
        L.ncnn_cuda_launch_count.restype = C.c_ulonglong
        L.ncnn_cuda_conv2d_workspace_size.restype = C.c_size_t
        L.ncnn_cuda_conv2d_workspace_size.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(ret, what=""):
    if ret != 0:
        raise RuntimeError("%s returned %d: %s" % (what, ret, lib().ncnn_cuda_last_error().decode()))


def torch_dtype(elemtype):
    import torch
    return {F32: torch.float32, BF16: torch.bfloat16, F16: torch.float16}[elemtype]


class Blob(object):
    """A device blob in the backend's channel-innermost layout, backed by a torch tensor [n][P][cpitch]."""

    def __init__(self, shape, n, elemtype, cpitch_align=8, fill=None):
        import torch
        self.shape = tuple(shape)  # ncnn logical dims as numpy order: (w,), (h,w), (c,h,w), (c,d,h,w)
        self.n = n
        self.elemtype = elemtype
        dims = len(shape)
        if dims == 1:
            P, Cc = 1, shape[0]
        elif dims == 2:
            P, Cc = shape[0], shape[1]
        elif dims == 3:
            P, Cc = shape[1] * shape[2], shape[0]
        else:
            P, Cc = shape[1] * shape[2] * shape[3], shape[0]
        self.P, self.C = P, Cc
        self.cpitch = (Cc + cpitch_align - 1) // cpitch_align * cpitch_align
        self.t = torch.empty((n, P, self.cpitch), dtype=torch_dtype(elemtype), device="cuda")
        if fill is not None:
            self.t.fill_(fill)

    @staticmethod
    def from_numpy(a, elemtype, batched=True, cpitch_align=8, pad_fill=float("nan")):
        """a: (n, ...) planar ncnn order -> device blob; padding lanes get `pad_fill` so kernels that read them show up"""
        import torch
        a = np.asarray(a, np.float32)
        if not batched:
            a = a[None]
        n = a.shape[0]
        shp = a.shape[1:]
        b = Blob(shp, n, elemtype, cpitch_align, fill=pad_fill)
        t = torch.from_numpy(a).cuda()
        dims = len(shp)
        if dims == 1:
            v = t.reshape(n, 1, shp[0])
        elif dims == 2:
            v = t
        elif dims == 3:
            v = t.permute(0, 2, 3, 1).reshape(n, b.P, b.C)
        else:
            v = t.permute(0, 2, 3, 4, 1).reshape(n, b.P, b.C)
        b.t[:, :, :b.C] = v.to(b.t.dtype)
        return b

    def numpy(self):
        """-> (n, ...) planar ncnn order, float32"""
        v = self.t[:, :, :self.C].float()
        n, shp = self.n, self.shape
        dims = len(shp)
        if dims == 1:
            o = v.reshape(n, shp[0])
        elif dims == 2:
            o = v
        elif dims == 3:
            o = v.reshape(n, shp[1], shp[2], shp[0]).permute(0, 3, 1, 2)
        else:
            o = v.reshape(n, shp[1], shp[2], shp[3], shp[0]).permute(0, 4, 1, 2, 3)
        return o.contiguous().cpu().numpy()

    def desc(self):
        shp = self.shape
        dims = len(shp)
        w = shp[-1]
        h = shp[-2] if dims >= 2 else 1
        d = shp[1] if dims == 4 else 1
        c = shp[0] if dims >= 3 else 1
        return Tensor(self.t.data_ptr(), dims, w, h, d, c, self.n, self.elemtype, self.cpitch, self.P * self.cpitch)


def act(type_=0, p0=0.0, p1=0.0):
    return Activation(type_, p0, p1)


def fptr(a):
    a = np.ascontiguousarray(a, np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))
