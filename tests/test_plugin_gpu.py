"""The plugin table end to end (include/c_api.h section 8, the reference's src/c_api.h:252-318): a user operator written
against `ncnn_layer_t` must run in the middle of a CUDA graph and must be able to replace a built-in type.

  * test_custom_layer_in_cuda_graph   -- the reference's tests/test_c_api.cpp:183-250 (MyLayer: in-place +100 registered with
    ncnn_net_register_custom_layer_by_type) grown into a real graph: Convolution -> MyLayer -> ReLU -> Convolution.  The SAME
    Python callbacks are registered in the product (CUDA graph: the runtime downloads the blob, runs the host layer per
    sample, uploads the result) and in the reference (CPU); results must agree to the fp32 bound, batched and unbatched.
  * test_override_builtin_convolution -- tests/test_squeezenet.cpp:233-405 (a user class registered under the built-in
    Convolution type): every Convolution of the graph is served by a host implementation written in numpy that reads its
    parameters and weights through ncnn_paramdict_get_* / the ncnn_modelbin_t function table.  The overridden net must agree
    with the product's own CUDA convolution and with the reference.
"""
import ctypes as C

import numpy as np
import pytest

from netutil import nerr

pytestmark = pytest.mark.gpu

_CREATOR = C.CFUNCTYPE(C.c_void_p, C.c_void_p)
_DESTROYER = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)


class _ModelBin(C.Structure):
    pass


_ModelBin._fields_ = [("pthis", C.c_void_p),
                      ("load_1d", C.CFUNCTYPE(C.c_void_p, C.POINTER(_ModelBin), C.c_int, C.c_int)),
                      ("load_2d", C.CFUNCTYPE(C.c_void_p, C.POINTER(_ModelBin), C.c_int, C.c_int, C.c_int)),
                      ("load_3d", C.CFUNCTYPE(C.c_void_p, C.POINTER(_ModelBin), C.c_int, C.c_int, C.c_int, C.c_int))]


def product():
    from ncnn_b200 import capi
    return capi.library()


def bind_layer_api(api):
    from ncnn_b200 import capi
    L = api.lib
    L.ncnn_layer_create.restype = C.POINTER(capi._Layer)
    L.ncnn_layer_create.argtypes = []
    L.ncnn_layer_destroy.argtypes = [C.POINTER(capi._Layer)]
    L.ncnn_layer_set_one_blob_only.argtypes = [C.POINTER(capi._Layer), C.c_int]
    L.ncnn_layer_set_support_inplace.argtypes = [C.POINTER(capi._Layer), C.c_int]
    L.ncnn_net_register_custom_layer_by_type.argtypes = [C.c_void_p, C.c_char_p, _CREATOR, _DESTROYER, C.c_void_p]
    L.ncnn_net_register_custom_layer_by_type.restype = None
    L.ncnn_paramdict_get_int.restype = C.c_int
    L.ncnn_paramdict_get_int.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ncnn_mat_create_3d.restype = C.c_void_p
    L.ncnn_mat_create_3d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    return capi


class Plugin(object):
    """keeps the ctypes callbacks of one registration alive and counts the calls"""

    def __init__(self):
        self.keep = []
        self.calls = 0
        self.layers = {}


def register_mylayer(api, net, plug):
    """tests/test_c_api.cpp:183-216: one_blob_only + support_inplace, forward_inplace_1 adds 100 to every element"""
    capi = bind_layer_api(api)
    L = api.lib

    def fwd_inplace(layer, mat, opt):
        v = api._view(C.c_void_p(mat))
        v += 100.0
        plug.calls += 1
        return 0

    cb = capi._Layer._fields_[7][1](fwd_inplace)  # forward_inplace_1

    def creator(userdata):
        layer = L.ncnn_layer_create()
        L.ncnn_layer_set_one_blob_only(layer, 1)
        L.ncnn_layer_set_support_inplace(layer, 1)
        layer.contents.forward_inplace_1 = cb
        return C.cast(layer, C.c_void_p).value

    def destroyer(layer, userdata):
        L.ncnn_layer_destroy(C.cast(layer, C.POINTER(capi._Layer)))

    c1, d1 = _CREATOR(creator), _DESTROYER(destroyer)
    plug.keep += [cb, c1, d1]
    L.ncnn_net_register_custom_layer_by_type(net, b"MyLayer", c1, d1, None)


def conv2d_numpy(x, w, b, stride, pad):
    """x (c, h, w), w (outch, c, kh, kw) -> (outch, oh, ow), fp32 accumulation in the reference's order is not needed: fp64 here"""
    c, h, wd = x.shape
    oc, _, kh, kw = w.shape
    xp = np.zeros((c, h + 2 * pad, wd + 2 * pad), np.float64)
    xp[:, pad:pad + h, pad:pad + wd] = x
    oh = (h + 2 * pad - kh) // stride + 1
    ow = (wd + 2 * pad - kw) // stride + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, (kh, kw), axis=(1, 2))[:, ::stride, ::stride][:, :oh, :ow]
    y = np.einsum("chwij,ocij->ohw", win, w.astype(np.float64))
    if b is not None:
        y += b.astype(np.float64)[:, None, None]
    return y.astype(np.float32)


def register_numpy_convolution(api, net, plug):
    """a user Convolution (load_param / load_model / forward_1 through the C tables) registered under the BUILT-IN type name"""
    capi = bind_layer_api(api)
    L = api.lib
    F = dict((name, ftype) for name, ftype in capi._Layer._fields_)

    def load_param(layer, pd):
        key = C.cast(layer, C.c_void_p).value
        g = lambda i, d: L.ncnn_paramdict_get_int(pd, i, d)  # noqa: E731
        kw = g(1, 0)
        plug.layers[key] = dict(outch=g(0, 0), kw=kw, kh=g(11, kw), stride=g(3, 1), pad=g(4, 0), bias=g(5, 0), wsize=g(6, 0), act=g(9, 0))
        return 0

    def load_model(layer, mb):
        key = C.cast(layer, C.c_void_p).value
        st = plug.layers[key]
        mbp = C.cast(mb, C.POINTER(_ModelBin))
        m = mbp.contents.load_1d(mbp, st["wsize"], 0)
        st["w"] = api.mat_to_numpy(C.c_void_p(m)).reshape(-1)
        L.ncnn_mat_destroy(C.c_void_p(m))
        st["b"] = None
        if st["bias"]:
            m = mbp.contents.load_1d(mbp, st["outch"], 1)
            st["b"] = api.mat_to_numpy(C.c_void_p(m)).reshape(-1)
            L.ncnn_mat_destroy(C.c_void_p(m))
        return 0

    def forward_1(layer, bottom, top_out, opt):
        key = C.cast(layer, C.c_void_p).value
        st = plug.layers[key]
        x = api.mat_to_numpy(C.c_void_p(bottom))
        inch = x.shape[0]
        w = st["w"].reshape(st["outch"], inch, st["kh"], st["kw"])
        y = conv2d_numpy(x, w, st["b"], st["stride"], st["pad"])
        if st["act"] == 1:
            y = np.maximum(y, 0)
        m = L.ncnn_mat_create_3d(y.shape[2], y.shape[1], y.shape[0], None)
        api._view(C.c_void_p(m))[...] = y
        top_out[0] = m
        plug.calls += 1
        return 0

    cbs = [F["load_param"](load_param), F["load_model"](load_model), F["forward_1"](forward_1)]

    def creator(userdata):
        layer = L.ncnn_layer_create()
        L.ncnn_layer_set_one_blob_only(layer, 1)
        L.ncnn_layer_set_support_inplace(layer, 0)
        layer.contents.load_param = cbs[0]
        layer.contents.load_model = cbs[1]
        layer.contents.forward_1 = cbs[2]
        return C.cast(layer, C.c_void_p).value

    def destroyer(layer, userdata):
        L.ncnn_layer_destroy(C.cast(layer, C.POINTER(capi._Layer)))

    c1, d1 = _CREATOR(creator), _DESTROYER(destroyer)
    plug.keep += cbs + [c1, d1]
    L.ncnn_net_register_custom_layer_by_type(net, b"Convolution", c1, d1, None)


def build_net(api, text, weights, opt, register=None):
    from ncnn_b200 import capi
    L = api.lib
    net = L.ncnn_net_create()
    L.ncnn_net_set_option(net, opt)
    plug = Plugin()
    if register:
        register(api, net, plug)
    assert L.ncnn_net_load_param_memory(net, text.encode()) == 0
    rd = capi.MemoryReader(api, weights)
    try:
        assert L.ncnn_net_load_model_datareader(net, rd.dr) == 0
    finally:
        rd.close()
    return net, plug


def run_net(api, net, x, batched, in_name=b"data", out_name=b"output"):
    L = api.lib
    m = api.mat_from_numpy(x, batched=batched)
    ex = L.ncnn_extractor_create(net)
    out = C.c_void_p()
    try:
        assert L.ncnn_extractor_input(ex, in_name, m) == 0
        assert L.ncnn_extractor_extract(ex, out_name, C.byref(out)) == 0
        return api.mat_to_numpy(out, force_batch=batched)
    finally:
        if out:
            L.ncnn_mat_destroy(out)
        L.ncnn_extractor_destroy(ex)
        L.ncnn_mat_destroy(m)


def model_bytes(rng, specs):
    """.bin stream for Convolution layers: per layer a zero fp32 tag, the weights, then the bias (src/modelbin.cpp:75-151)"""
    out = b""
    for wshape, bias in specs:
        w = (rng.uniform(-1, 1, wshape) * np.sqrt(3.0 / np.prod(wshape[1:]))).astype(np.float32)
        out += np.zeros(1, np.uint32).tobytes() + w.tobytes()
        if bias:
            out += rng.uniform(-1, 1, (wshape[0],)).astype(np.float32).tobytes()
    return out


FP32 = dict(use_fp16_storage=0, use_fp16_packed=0, use_fp16_arithmetic=0, use_bf16_storage=0)


@pytest.mark.parametrize("batched", [False, True])
def test_custom_layer_in_cuda_graph(ref, batched):
    ours = product()
    text = ("7767517\n5 5\n"
            "Input data 0 1 data 0=12 1=10 2=3\n"
            "Convolution conv1 1 1 data c1 0=8 1=3 4=1 5=1 6=216\n"
            "MyLayer mylayer 1 1 c1 c2\n"
            "ReLU relu 1 1 c2 c3\n"
            "Convolution conv2 1 1 c3 output 0=4 1=1 5=1 6=32\n")
    rng = np.random.default_rng(5)
    weights = model_bytes(rng, [((8, 3, 3, 3), True), ((4, 8, 1, 1), True)])
    x = rng.uniform(-1, 1, ((3, 3, 10, 12) if batched else (3, 10, 12))).astype(np.float32)
    res = []
    for api in (ours, ref):
        opt = api.make_option(1, **FP32) if api is ours else api.strict_fp32_option(num_threads=1)
        net, plug = build_net(api, text, weights, opt, register=register_mylayer)
        got = run_net(api, net, x, batched)
        # a host layer knows nothing about batches: once per sample in both runtimes (src/net.cpp:654-705)
        assert plug.calls == (3 if batched else 1), plug.calls
        res.append(got)
        api.lib.ncnn_net_destroy(net)
        api.lib.ncnn_option_destroy(opt)
    assert res[0].shape == res[1].shape
    assert np.abs(res[1]).max() > 50.0  # the +100 went through the second convolution
    assert nerr(res[0], res[1]) <= 1e-5


def test_override_builtin_convolution(ref):
    ours = product()
    text = ("7767517\n5 5\n"
            "Input data 0 1 data 0=16 1=16 2=3\n"
            "Convolution conv1 1 1 data c1 0=8 1=3 3=2 4=1 5=1 6=216 9=1\n"
            "Pooling pool1 1 1 c1 p1 0=0 1=2 2=2\n"
            "Convolution conv2 1 1 p1 c2 0=6 1=1 5=1 6=48\n"
            "Softmax prob 1 1 c2 output 0=0 1=1\n")
    rng = np.random.default_rng(9)
    weights = model_bytes(rng, [((8, 3, 3, 3), True), ((6, 8, 1, 1), True)])
    x = rng.uniform(-1, 1, (2, 3, 16, 16)).astype(np.float32)
    opt = ours.make_option(1, **FP32)
    plain, _ = build_net(ours, text, weights, opt)
    want_cuda = run_net(ours, plain, x, True)
    ours.lib.ncnn_net_destroy(plain)
    over, plug = build_net(ours, text, weights, opt, register=register_numpy_convolution)
    got = run_net(ours, over, x, True)
    assert plug.calls == 2 * 2, "both Convolution layers must be served by the user class, once per sample: %d" % plug.calls
    assert len(plug.layers) == 2
    ours.lib.ncnn_net_destroy(over)
    ours.lib.ncnn_option_destroy(opt)
    ropt = ref.strict_fp32_option(num_threads=1)
    rnet, _ = build_net(ref, text, weights, ropt)
    want_ref = run_net(ref, rnet, x, True)
    ref.lib.ncnn_net_destroy(rnet)
    ref.lib.ncnn_option_destroy(ropt)
    assert nerr(want_cuda, want_ref) <= 1e-5
    assert nerr(got, want_ref) <= 1e-5
    assert nerr(got, want_cuda) <= 1e-5
