"""Host-side logic of the product library checked against the reference's own host code through the SAME C API
(src/c_api.h) -- runs without a GPU.  Both libraries are driven by one ctypes binding (ncnn_b200/capi.py), so every
assertion is "product == reference" on identical calls:

  * Mat allocation layout: cstep 16-byte alignment, nstep 4 KiB alignment, elemsize/elempack, batch views
    (src/mat.h:50-382, src/mat.cpp:299-861, tests/test_mat_batch.cpp)
  * Layer capability flags the executor reads (src/layer.h:46-90) for every hot-path layer type
  * Net::load_param on the five benchmark graphs: input/output blob names and counts (src/net.cpp:1255-1560)
  * ParamDict set/get round trip (src/paramdict.cpp)
  * the .bin byte stream the seeded weight generator writes is consumed to the last byte by the reference's
    load_model -- i.e. tools/modelzoo.py's understanding of ModelBin (src/modelbin.cpp:75-151) matches the reference."""
import ctypes as C

import numpy as np
import pytest

from netutil import modelzoo

HOT_PATH_LAYERS = ["Convolution", "ConvolutionDepthWise", "Pooling", "InnerProduct", "Gemm", "ReLU", "Eltwise", "BinaryOp", "Concat", "Split", "Softmax",
                   "Interp", "Swish", "Sigmoid", "Slice", "Reshape", "Permute", "Flatten", "Dropout", "Input", "Padding", "BatchNorm", "Scale", "ShuffleChannel", "LRN", "Noop", "Crop", "Deconvolution", "DeconvolutionDepthWise", "Reduction", "MemoryData", "LayerNorm", "GELU", "MultiHeadAttention"]


@pytest.fixture(scope="module")
def ours():
    from ncnn_b200 import capi
    return capi.library()


SHAPES = [(1,), (7,), (1000,), (4097,), (5, 3), (224, 224), (227, 227, 3), (13, 13, 1000), (7, 7, 2048), (1, 1, 5), (3, 5, 7, 2), (20, 20, 2, 64)]


@pytest.mark.parametrize("n", [0, 1, 2, 5])
def test_mat_layout_matches_reference(ours, ref, n):
    for shp in SHAPES:
        mats = []
        for api in (ours, ref):
            L = api.lib
            if n == 0:
                f = getattr(L, "ncnn_mat_create_%dd" % len(shp))
                m = f(*shp, None)
            else:
                f = getattr(L, "ncnn_mat_create_%dd_batch" % len(shp))
                m = f(*shp, n, None)
            mats.append((api, m))
        desc = []
        for api, m in mats:
            L = api.lib
            desc.append(tuple(int(getattr(L, "ncnn_mat_get_" + k)(m)) for k in ("dims", "w", "h", "d", "c", "n", "elemsize", "elempack", "cstep", "nstep")))
            assert L.ncnn_mat_get_data(m) % 16 == 0
            L.ncnn_mat_destroy(m)
        assert desc[0] == desc[1], (shp, n, desc)
        cstep, nstep = desc[0][8], desc[0][9]
        if len(shp) >= 3:
            assert (cstep * 4) % 16 == 0
        if n > 1:
            assert (nstep * 4) % 4096 == 0  # src/mat.cpp batch stride alignment


def test_mat_batch_roundtrip(ours):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 4, 5, 6)).astype(np.float32)
    m = ours.mat_from_numpy(a, batched=True)
    assert ours.lib.ncnn_mat_get_n(m) == 3 and ours.lib.ncnn_mat_get_c(m) == 4
    assert np.array_equal(ours.mat_to_numpy(m, force_batch=True), a)
    ours.lib.ncnn_mat_destroy(m)


def test_layer_flags_match_reference(ours, ref):
    for t in HOT_PATH_LAYERS:
        flags = []
        for api in (ours, ref):
            ly = api.lib.ncnn_layer_create_by_type(t.encode())
            assert ly, (api.path, t)
            flags.append((api.lib.ncnn_layer_get_one_blob_only(ly), api.lib.ncnn_layer_get_support_inplace(ly)))
            api.lib.ncnn_layer_destroy(ly)
        assert flags[0] == flags[1], (t, flags)


def test_unknown_layer_type_is_null(ours):
    assert not ours.lib.ncnn_layer_create_by_type(b"NoSuchLayer")


@pytest.mark.parametrize("name", ["squeezenet_v1_1", "mobilenet_v2", "resnet50", "vgg16", "yolov8s"])
def test_load_param_matches_reference(ours, ref, name):
    text = modelzoo.param_text(name)
    seen = []
    for api in (ours, ref):
        L = api.lib
        net = L.ncnn_net_create()
        assert L.ncnn_net_load_param_memory(net, text.encode()) == 0
        ins = [L.ncnn_net_get_input_name(net, i) for i in range(L.ncnn_net_get_input_count(net))]
        outs = [L.ncnn_net_get_output_name(net, i) for i in range(L.ncnn_net_get_output_count(net))]
        seen.append((ins, outs))
        L.ncnn_net_destroy(net)
    assert seen[0] == seen[1] and seen[0][0] and seen[0][1]


def _extra_graphs():
    import glob
    import os
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "benchmark_graphs")
    return sorted(glob.glob(os.path.join(d, "*.param")))


@pytest.mark.parametrize("path", _extra_graphs(), ids=lambda p: p.split("/")[-1][:-6])
def test_load_param_benchmark_set_matches_reference(ours, ref, path):
    """every fp32 graph of the reference's benchmark set kept under tests/golden/benchmark_graphs/ parses in the product runtime to the same input /
    output blob names as in the reference; graphs that end in host-side detection post-processing are refused by the product
    (no creator for PriorBox / DetectionOutput / Yolo*DetectionOutput: a graph never falls back to the CPU silently)"""
    text = open(path).read()
    types = set(l.split()[0] for l in text.splitlines()[2:] if l.strip())
    post = types & {"PriorBox", "DetectionOutput", "YoloDetectionOutput", "Yolov3DetectionOutput"}
    seen = []
    for api in (ours, ref):
        L = api.lib
        net = L.ncnn_net_create()
        r = L.ncnn_net_load_param_memory(net, text.encode())
        if api is ours and post:
            assert r != 0, "a graph with %s must be refused" % sorted(post)
            L.ncnn_net_destroy(net)
            continue
        assert r == 0
        ins = [L.ncnn_net_get_input_name(net, i) for i in range(L.ncnn_net_get_input_count(net))]
        outs = [L.ncnn_net_get_output_name(net, i) for i in range(L.ncnn_net_get_output_count(net))]
        seen.append((ins, outs))
        L.ncnn_net_destroy(net)
    if not post:
        assert seen[0] == seen[1] and seen[0][0] and seen[0][1]


def _graph_signature(L, net):
    ins = [L.ncnn_net_get_input_name(net, i) for i in range(L.ncnn_net_get_input_count(net))]
    outs = [L.ncnn_net_get_output_name(net, i) for i in range(L.ncnn_net_get_output_count(net))]
    return ins, outs


@pytest.mark.parametrize("name", ["squeezenet_v1_1", "mobilenet_v2", "resnet50", "vgg16", "yolov8s"])
def test_load_param_through_datareader_scan(ours, ref, name):
    """text params come through DataReader::scan (src/net.cpp:1305 SCAN_VALUE, src/paramdict.cpp:263-480): (1) the stock memory
    reader, which has no length to drain with read(); (2) a custom C-API reader that implements scan() ONLY (read returns 0).
    Both must give the graph ncnn_net_load_param_memory gives, in the product as in the reference."""
    from ncnn_b200 import capi
    text = modelzoo.param_text(name).encode() + b"\0"
    libc = C.CDLL(None)
    for api in (ours, ref):
        L = api.lib
        net = L.ncnn_net_create()
        assert L.ncnn_net_load_param_memory(net, text) == 0
        want = _graph_signature(L, net)
        L.ncnn_net_destroy(net)

        # (1) ncnn_datareader_create_from_memory
        buf = C.create_string_buffer(text, len(text))
        cursor = C.c_void_p(C.addressof(buf))
        L.ncnn_datareader_create_from_memory.restype = C.POINTER(capi._DataReader)
        L.ncnn_datareader_create_from_memory.argtypes = [C.POINTER(C.c_void_p)]
        dr = L.ncnn_datareader_create_from_memory(C.byref(cursor))
        net = L.ncnn_net_create()
        assert L.ncnn_net_load_param_datareader(net, dr) == 0
        assert _graph_signature(L, net) == want
        assert 0 < cursor.value - C.addressof(buf) <= len(text)  # the reader consumed the text and stopped inside it
        L.ncnn_net_destroy(net)
        L.ncnn_datareader_destroy(dr)

        # (2) scan-only custom reader (sscanf with %n over a Python-held buffer, as DataReaderFromMemory::scan does)
        state = {"pos": 0}
        base = C.addressof(buf)

        def _scan(dr_, fmt, out):
            consumed = C.c_int(0)
            n = libc.sscanf(C.c_void_p(base + state["pos"]), fmt + b"%n", C.c_void_p(out), C.byref(consumed))
            state["pos"] += consumed.value
            return n if consumed.value > 0 else 0

        def _read(dr_, b, size):
            return 0

        dr = L.ncnn_datareader_create()
        scan_cb, read_cb = capi._SCAN_FN(_scan), capi._READ_FN(_read)
        dr.contents.scan = scan_cb
        dr.contents.read = read_cb
        net = L.ncnn_net_create()
        assert L.ncnn_net_load_param_datareader(net, dr) == 0
        assert _graph_signature(L, net) == want
        L.ncnn_net_destroy(net)
        L.ncnn_datareader_destroy(dr)


def test_load_param_rejects_garbage(ours):
    L = ours.lib
    for bad in (b"", b"1234\n1 1\n", b"7767517\n2 2\nInput data 0 1 data\n"):
        net = L.ncnn_net_create()
        assert L.ncnn_net_load_param_memory(net, bad) != 0
        L.ncnn_net_destroy(net)


def test_paramdict_roundtrip(ours, ref):
    for api in (ours, ref):
        L = api.lib
        L.ncnn_paramdict_get_int.restype = C.c_int
        L.ncnn_paramdict_get_int.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ncnn_paramdict_get_float.restype = C.c_float
        L.ncnn_paramdict_get_float.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.ncnn_paramdict_get_type.restype = C.c_int
        L.ncnn_paramdict_get_type.argtypes = [C.c_void_p, C.c_int]
        pd = api.make_paramdict({0: 64, 1: 3, 9: 2, 18: 0.25, 10: [0.1, 0.2]})
        assert L.ncnn_paramdict_get_int(pd, 0, -1) == 64
        assert L.ncnn_paramdict_get_int(pd, 5, -7) == -7  # default for an absent id
        assert abs(L.ncnn_paramdict_get_float(pd, 18, 0.0) - 0.25) < 1e-7
        assert L.ncnn_paramdict_get_type(pd, 31) == 0
        L.ncnn_paramdict_destroy(pd)


@pytest.mark.parametrize("name", ["squeezenet_v1_1", "mobilenet_v2", "yolov8s"])
def test_seeded_weight_stream_is_consumed_exactly(ref, name):
    """the reference reads the generated .bin to its last byte and not beyond"""
    from ncnn_b200 import capi
    text = modelzoo.param_text(name)
    weights = modelzoo.random_model_bytes(text, seed=7)
    L = ref.lib
    opt = ref.strict_fp32_option()
    net = L.ncnn_net_create()
    L.ncnn_net_set_option(net, opt)
    assert L.ncnn_net_load_param_memory(net, text.encode()) == 0
    rd = capi.MemoryReader(ref, weights)
    assert L.ncnn_net_load_model_datareader(net, rd.dr) == 0
    assert rd.pos == len(weights), (rd.pos, len(weights))
    rd.close()
    L.ncnn_net_destroy(net)
    # a truncated stream must be refused
    net = L.ncnn_net_create()
    L.ncnn_net_set_option(net, opt)
    L.ncnn_net_load_param_memory(net, text.encode())
    rd = capi.MemoryReader(ref, weights[:len(weights) // 2])
    assert L.ncnn_net_load_model_datareader(net, rd.dr) != 0
    rd.close()
    L.ncnn_net_destroy(net)
    L.ncnn_option_destroy(opt)


def test_from_pixels_and_normalize_match_reference(ours, ref):
    """host pre-processing: ncnn_mat_from_pixels for every conversion of the reference's table (src/mat_pixel.cpp:2440-2545) and
    ncnn_mat_substract_mean_normalize, bit-exact against the reference's own implementation on random images with a row stride"""
    rng = np.random.default_rng(23)
    RGB, BGR, GRAY, RGBA, BGRA = 1, 2, 3, 4, 5
    chans = {RGB: 3, BGR: 3, GRAY: 1, RGBA: 4, BGRA: 4}
    pairs = [(a, a) for a in chans] + [(RGB, BGR), (BGR, RGB), (RGB, GRAY), (BGR, GRAY), (RGB, RGBA), (BGR, BGRA), (BGR, RGBA), (RGB, BGRA), (GRAY, RGB), (GRAY, BGR),
                                       (GRAY, RGBA), (GRAY, BGRA), (RGBA, RGB), (BGRA, BGR), (RGBA, BGR), (BGRA, RGB), (RGBA, GRAY), (BGRA, GRAY), (RGBA, BGRA), (BGRA, RGBA)]
    for api in (ours, ref):
        api.lib.ncnn_mat_from_pixels.restype = C.c_void_p
        api.lib.ncnn_mat_from_pixels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        api.lib.ncnn_mat_substract_mean_normalize.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    w, h = 37, 11
    for (a, b) in pairs:
        ch = chans[a]
        stride = w * ch + 5
        buf = rng.integers(0, 256, (h, stride), dtype=np.uint8)
        t = a if a == b else (a | (b << 16))
        mean = rng.uniform(0, 200, chans[b]).astype(np.float32)
        norm = rng.uniform(0.005, 0.05, chans[b]).astype(np.float32)
        res = []
        for api in (ours, ref):
            m = api.lib.ncnn_mat_from_pixels(buf.ctypes.data_as(C.c_void_p), t, w, h, stride, None)
            plain = api.mat_to_numpy(C.c_void_p(m)).copy()
            api.lib.ncnn_mat_substract_mean_normalize(C.c_void_p(m), mean.ctypes.data_as(C.c_void_p), norm.ctypes.data_as(C.c_void_p))
            res.append((plain, api.mat_to_numpy(C.c_void_p(m)).copy()))
            api.lib.ncnn_mat_destroy(C.c_void_p(m))
        assert res[0][0].shape == res[1][0].shape == (chans[b], h, w), (a, b, res[0][0].shape, res[1][0].shape)
        assert np.array_equal(res[0][0], res[1][0]), ("from_pixels", a, b)
        assert np.allclose(res[0][1], res[1][1], rtol=0, atol=1e-5), ("substract_mean_normalize", a, b)


def test_to_pixels_matches_reference(ours, ref):
    """host post-processing: ncnn_mat_to_pixels for the conversion families the reference implements (src/mat_pixel.cpp:2710-2753),
    byte-exact against the reference on float images with fractions, negatives and values above 255, with a row stride"""
    rng = np.random.default_rng(29)
    RGB, BGR, GRAY, RGBA, BGRA = 1, 2, 3, 4, 5
    chans = {RGB: 3, BGR: 3, GRAY: 1, RGBA: 4, BGRA: 4}
    pairs = [(a, a) for a in chans] + [(RGB, BGR), (BGR, RGB), (RGB, RGBA), (BGR, BGRA), (BGR, RGBA), (RGB, BGRA), (GRAY, RGBA), (GRAY, BGRA), (RGBA, BGRA), (BGRA, RGBA)]
    for api in (ours, ref):
        api.lib.ncnn_mat_to_pixels.restype = None
        api.lib.ncnn_mat_to_pixels.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    w, h = 29, 13
    for (a, b) in pairs:
        x = rng.uniform(-40.0, 300.0, (chans[a], h, w)).astype(np.float32)
        x[0, 0, :6] = [0.0, 0.999, 254.999, 255.0, 255.5, -0.5]
        stride = w * chans[b] + 7
        t = a if a == b else (a | (b << 16))
        res = []
        for api in (ours, ref):
            m = api.mat_from_numpy(x)
            buf = np.full((h, stride), 171, np.uint8)
            api.lib.ncnn_mat_to_pixels(m, buf.ctypes.data_as(C.c_void_p), t, stride)
            api.lib.ncnn_mat_destroy(m)
            res.append(buf)
        assert np.array_equal(res[0], res[1]), ("to_pixels", a, b)
        assert (res[0][:, w * chans[b]:] == 171).all()  # the stride padding is left alone


def test_resize_tables_reproduce_reference_resize(ref):
    """ncnn_cuda_resize_tables (host code of the product, no device work) + the integer formula the device kernel evaluates per
    output pixel (csrc/cuda/layout.cu pixels_resize_to_blob_kernel, restated here in numpy) reproduce the reference's
    ncnn_mat_from_pixels_resize bit for bit: 1 / 3 / 4 channels, up- and down-scaling, odd sizes, the 2 x 2 minimum, a row stride"""
    import cabi
    L = C.CDLL(cabi.LIB_PATH)
    R = ref.lib
    R.ncnn_mat_from_pixels_resize.restype = C.c_void_p
    R.ncnn_mat_from_pixels_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(5)
    for (ch, typ) in [(3, 1), (1, 3), (4, 4)]:
        for (sw, sh, w, h) in [(37, 23, 16, 16), (16, 16, 37, 23), (64, 48, 224, 224), (301, 199, 224, 224), (2, 2, 5, 7), (50, 50, 49, 51), (640, 360, 320, 192)]:
            stride = sw * ch + 3
            img = rng.integers(0, 256, (sh, stride), dtype=np.uint8)
            m = R.ncnn_mat_from_pixels_resize(img.ctypes.data_as(C.c_void_p), typ, sw, sh, stride, w, h, None)
            want = ref.mat_to_numpy(C.c_void_p(m)).copy()
            R.ncnn_mat_destroy(C.c_void_p(m))
            assert L.ncnn_cuda_resize_tables_count(w, h) == 3 * (w + h)
            tab = (C.c_int * (3 * (w + h)))()
            assert L.ncnn_cuda_resize_tables(sw, sh, w, h, tab) == 0
            t = np.frombuffer(tab, np.int32).astype(np.int64)
            xofs, yofs, al, be = t[:w], t[w:w + h], t[w + h:w + h + 2 * w].reshape(w, 2), t[w + h + 2 * w:].reshape(h, 2)
            src = img[:, :sw * ch].reshape(sh, sw, ch).astype(np.int64)
            got = np.zeros((ch, h, w), np.float32)
            for c in range(ch):
                S = src[:, :, c]
                r = (S[:, xofs] * al[:, 0] + S[:, xofs + 1] * al[:, 1]) >> 4
                q = (((be[:, 0:1] * r[yofs]) >> 16) + ((be[:, 1:2] * r[yofs + 1]) >> 16) + 2) >> 2
                got[c] = np.clip(q, 0, 255)
            assert np.array_equal(got, want), (ch, sw, sh, w, h)
    assert L.ncnn_cuda_resize_tables(1, 5, 4, 4, tab) != 0  # the reference reads column sx + 1: a 1-pixel-wide source has none


def test_pixels_resize_and_roi_match_reference(ours, ref):
    """host pre/post-processing with the reference's 8-bit bilinear resize: ncnn_mat_from_pixels_resize / _roi / _roi_resize and
    ncnn_mat_to_pixels_resize (src/mat_pixel.cpp:2546-2806), bit-exact against the reference: several conversions, up- and
    down-scaling, row strides, and the equal-size shortcut"""
    rng = np.random.default_rng(37)
    RGB, BGR, GRAY, RGBA, BGRA = 1, 2, 3, 4, 5
    chans = {RGB: 3, BGR: 3, GRAY: 1, RGBA: 4, BGRA: 4}
    vp, ci = C.c_void_p, C.c_int
    for api in (ours, ref):
        l = api.lib
        l.ncnn_mat_from_pixels_resize.restype = vp
        l.ncnn_mat_from_pixels_resize.argtypes = [vp, ci, ci, ci, ci, ci, ci, vp]
        l.ncnn_mat_from_pixels_roi.restype = vp
        l.ncnn_mat_from_pixels_roi.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, ci, vp]
        l.ncnn_mat_from_pixels_roi_resize.restype = vp
        l.ncnn_mat_from_pixels_roi_resize.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp]
        l.ncnn_mat_to_pixels_resize.restype = None
        l.ncnn_mat_to_pixels_resize.argtypes = [vp, vp, ci, ci, ci, ci]

    def both(fn):
        out = []
        for api in (ours, ref):
            m = fn(api.lib)
            out.append(api.mat_to_numpy(vp(m)).copy())
            api.lib.ncnn_mat_destroy(vp(m))
        return out

    for (a, b) in [(RGB, RGB), (BGR, RGB), (RGB, GRAY), (GRAY, GRAY), (GRAY, BGR), (RGBA, RGBA), (RGBA, BGR), (BGRA, GRAY), (RGB, BGRA)]:
        ch = chans[a]
        t = a if a == b else (a | (b << 16))
        for (w, h, tw, th) in [(41, 29, 24, 24), (24, 24, 41, 29), (33, 33, 33, 33), (160, 90, 64, 64)]:
            stride = w * ch + 6
            buf = rng.integers(0, 256, (h, stride), dtype=np.uint8)
            p = buf.ctypes.data_as(vp)
            g, r = both(lambda l: l.ncnn_mat_from_pixels_resize(p, t, w, h, stride, tw, th, None))
            assert g.shape == r.shape == (chans[b], th, tw) and np.array_equal(g, r), ("from_pixels_resize", a, b, w, h, tw, th)
            rx, ry, rw, rh = 3, 2, w - 7, h - 5
            g, r = both(lambda l: l.ncnn_mat_from_pixels_roi(p, t, w, h, stride, rx, ry, rw, rh, None))
            assert g.shape == r.shape == (chans[b], rh, rw) and np.array_equal(g, r), ("from_pixels_roi", a, b)
            g, r = both(lambda l: l.ncnn_mat_from_pixels_roi_resize(p, t, w, h, stride, rx, ry, rw, rh, tw, th, None))
            assert g.shape == r.shape == (chans[b], th, tw) and np.array_equal(g, r), ("from_pixels_roi_resize", a, b)
    for (a, b) in [(RGB, RGB), (GRAY, GRAY), (RGBA, RGBA), (RGB, BGR), (BGR, RGBA), (GRAY, BGRA), (RGBA, BGRA)]:
        t = a if a == b else (a | (b << 16))
        for (w, h, tw, th) in [(31, 17, 20, 20), (20, 20, 31, 17), (19, 19, 19, 19)]:
            x = rng.uniform(-30.0, 290.0, (chans[a], h, w)).astype(np.float32)
            stride = tw * chans[b] + (0 if (w, h) == (tw, th) else 5)  # the equal-size shortcut of the reference writes tight rows
            res = []
            for api in (ours, ref):
                m = api.mat_from_numpy(x)
                out = np.full((th, stride), 93, np.uint8)
                api.lib.ncnn_mat_to_pixels_resize(m, out.ctypes.data_as(vp), t, tw, th, stride)
                api.lib.ncnn_mat_destroy(m)
                res.append(out)
            assert np.array_equal(res[0], res[1]), ("to_pixels_resize", a, b, w, h, tw, th)


@pytest.mark.parametrize("which", ["DECONV_PARAM", "NORM_PARAM", "MHA_PARAM", "UNFUSED_PARAM", "efficientnetv2_b0"])
def test_seeded_weight_stream_of_neighbour_layers_is_consumed_exactly(ref, which):
    """same for the graphs that carry the later-added weight-bearing layers (Deconvolution[DepthWise], LayerNorm, MemoryData,
    MultiHeadAttention, BatchNorm, Scale): tools/modelzoo.py writes exactly the bytes the reference's load_model reads"""
    import os
    from ncnn_b200 import capi
    import test_nets_gpu as tn
    if hasattr(tn, which):
        text = getattr(tn, which)
    else:
        text = open(os.path.join(tn.EXTRA, which + ".param")).read()
    weights = modelzoo.random_model_bytes(text, seed=7)
    L = ref.lib
    opt = ref.strict_fp32_option()
    net = L.ncnn_net_create()
    L.ncnn_net_set_option(net, opt)
    assert L.ncnn_net_load_param_memory(net, text.encode()) == 0
    rd = capi.MemoryReader(ref, weights)
    assert L.ncnn_net_load_model_datareader(net, rd.dr) == 0
    assert rd.pos == len(weights), (rd.pos, len(weights))
    rd.close()
    L.ncnn_net_destroy(net)


def test_mat_elem_constructors_match_reference(ours, ref):
    """ncnn_mat_create_*_elem[_batch] / create_external_*_elem (src/c_api.h:124-137): element size and packing other than fp32 x 1,
    same dims / elemsize / elempack / cstep / nstep as the reference"""
    keys = ("dims", "w", "h", "d", "c", "n", "elemsize", "elempack", "cstep", "nstep")
    for (elemsize, elempack) in [(2, 1), (1, 1), (16, 4), (8, 4), (32, 8), (4, 1)]:
        for shp in [(7,), (5, 3), (13, 13, 6), (3, 5, 7, 2)]:
            for n in (0, 3):
                desc = []
                for api in (ours, ref):
                    L = api.lib
                    f = getattr(L, "ncnn_mat_create_%dd_elem%s" % (len(shp), "_batch" if n else ""))
                    f.restype = C.c_void_p
                    f.argtypes = [C.c_int] * len(shp) + [C.c_size_t, C.c_int] + ([C.c_int] if n else []) + [C.c_void_p]
                    m = C.c_void_p(f(*shp, elemsize, elempack, *([n] if n else []), None))
                    desc.append(tuple(int(getattr(L, "ncnn_mat_get_" + k)(m)) for k in keys))
                    L.ncnn_mat_destroy(m)
                assert desc[0] == desc[1], (elemsize, elempack, shp, n, desc)
            buf = np.zeros(4096, np.uint8)
            desc = []
            for api in (ours, ref):
                L = api.lib
                f = getattr(L, "ncnn_mat_create_external_%dd_elem" % len(shp))
                f.restype = C.c_void_p
                f.argtypes = [C.c_int] * len(shp) + [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
                m = C.c_void_p(f(*shp, buf.ctypes.data_as(C.c_void_p), elemsize, elempack, None))
                desc.append(tuple(int(getattr(L, "ncnn_mat_get_" + k)(m)) for k in keys) + (int(L.ncnn_mat_get_data(m)) == buf.ctypes.data,))
                L.ncnn_mat_destroy(m)
            assert desc[0] == desc[1] and desc[0][-1], (elemsize, elempack, shp, desc)


def _ref_type_index(ref, text):
    """typeindex of every operator of a graph in the ORACLE build (its registry only holds the operators it was built with, so
    its numbering differs from the stock one the product uses)"""
    L = ref.lib
    out = {}
    for t in sorted(set(l.split()[0] for l in text.splitlines()[2:] if l.strip())):
        layer = L.ncnn_layer_create_by_type(t.encode())  # typed by ncnn_b200/capi.py
        assert layer, t
        out[t] = int(L.ncnn_layer_get_typeindex(layer))
        L.ncnn_layer_destroy(layer)
    return out


@pytest.mark.parametrize("name", ["squeezenet_v1_1", "mobilenet_v2", "yolov8s"])
def test_param_bin_format(ours, ref, name):
    """the binary graph format.  tools/param2bin.py writes it; (1) the REFERENCE loads that image and computes
    exactly what it computes from the text form (so the writer's understanding of src/net.cpp:1667-1940 is right), (2) the product
    parses the stock-typeindex image -- from a DataReader and from memory -- to the same layers, blob wiring and input / output
    indices as from the text form, and refuses truncated images"""
    import param2bin
    from ncnn_b200 import capi
    size = 64 if name != "squeezenet_v1_1" else 227
    text = modelzoo.param_text(name)
    lines = text.splitlines()
    text = "\n".join(((" ".join(("0=%d" % size) if t.startswith("0=") else (("1=%d" % size) if t.startswith("1=") else t) for t in l.split())) if l.startswith("Input") else l)
                     for l in lines) + "\n"
    weights = modelzoo.random_model_bytes(text, seed=3)
    x = np.random.default_rng(1).uniform(-1, 1, (3, size, size)).astype(np.float32)
    # (1) the reference: text vs binary
    L = ref.lib
    res = []
    for form in ("text", "bin"):
        opt = ref.strict_fp32_option()
        net = L.ncnn_net_create()
        L.ncnn_net_set_option(net, opt)
        if form == "text":
            assert L.ncnn_net_load_param_memory(net, text.encode()) == 0
        else:
            image = param2bin.convert(text, _ref_type_index(ref, text))
            rd = capi.MemoryReader(ref, image)
            assert L.ncnn_net_load_param_bin_datareader(net, rd.dr) == 0
            assert rd.pos == len(image)
            rd.close()
        rd = capi.MemoryReader(ref, weights)
        assert L.ncnn_net_load_model_datareader(net, rd.dr) == 0 and rd.pos == len(weights)
        rd.close()
        i_in, i_out = L.ncnn_net_get_input_index(net, 0), L.ncnn_net_get_output_index(net, 0)
        ex = L.ncnn_extractor_create(net)
        m = ref.mat_from_numpy(x)
        assert L.ncnn_extractor_input_index(ex, i_in, m) == 0
        out = C.c_void_p()
        assert L.ncnn_extractor_extract_index(ex, i_out, C.byref(out)) == 0
        res.append((i_in, i_out, ref.mat_to_numpy(out).copy()))
        L.ncnn_mat_destroy(out)
        L.ncnn_mat_destroy(m)
        L.ncnn_extractor_destroy(ex)
        L.ncnn_net_destroy(net)
        L.ncnn_option_destroy(opt)
    assert res[0][:2] == res[1][:2] and np.array_equal(res[0][2], res[1][2])
    # (2) the product: structure from the binary image == structure from the text
    P = ours.lib
    P.ncnn_net_get_layer_type.restype = C.c_char_p
    P.ncnn_net_load_param_bin_memory.restype = C.c_size_t
    P.ncnn_net_load_param_bin_memory.argtypes = [C.c_void_p, C.c_void_p]
    image = param2bin.convert(text)

    def structure(net):
        return ([P.ncnn_net_get_layer_type(net, i) for i in range(P.ncnn_net_get_layer_count(net))],
                [P.ncnn_net_get_input_index(net, i) for i in range(P.ncnn_net_get_input_count(net))],
                [P.ncnn_net_get_output_index(net, i) for i in range(P.ncnn_net_get_output_count(net))])

    net = P.ncnn_net_create()
    assert P.ncnn_net_load_param_memory(net, text.encode()) == 0
    want = structure(net)
    P.ncnn_net_destroy(net)
    net = P.ncnn_net_create()
    rd = capi.MemoryReader(ours, image)
    assert P.ncnn_net_load_param_bin_datareader(net, rd.dr) == 0 and rd.pos == len(image)
    rd.close()
    assert structure(net) == want
    P.ncnn_net_destroy(net)
    net = P.ncnn_net_create()
    buf = np.frombuffer(image, np.uint8).copy()
    assert P.ncnn_net_load_param_bin_memory(net, buf.ctypes.data_as(C.c_void_p)) == len(image)
    assert structure(net) == want
    assert want[1] == [res[1][0]] and want[2] == [res[1][1]]  # same blob numbering as the reference
    P.ncnn_net_destroy(net)
    net = P.ncnn_net_create()
    rd = capi.MemoryReader(ours, image[:len(image) // 2])
    assert P.ncnn_net_load_param_bin_datareader(net, rd.dr) != 0
    rd.close()
    P.ncnn_net_destroy(net)


def test_modelbin_storage_tags_match_reference(ours, ref):
    """SURVEY 8f row f1, weight side: ModelBinFromDataReader (src/modelbin.cpp:75-151) on one byte stream -- a tag-0 fp32 blob, an
    fp16-tagged blob (0x01306B47, odd length: its 4-byte alignment padding must be skipped), a raw (type 1) blob and a type-0 scalar
    -- decoded to identical Mats by the product and the reference, and consumed to the same position"""
    import struct
    from ncnn_b200 import capi
    rng = np.random.default_rng(41)
    a = rng.uniform(-1, 1, 37).astype(np.float32)
    h = rng.uniform(-4, 4, 21).astype(np.float16)          # 42 bytes -> padded to 44
    r = rng.uniform(-1, 1, 9).astype(np.float32)
    stream = (struct.pack("<I", 0) + a.tobytes() + struct.pack("<I", 0x01306B47) + h.tobytes() + b"\0" * 2 + r.tobytes()
              + struct.pack("<I", 0) + struct.pack("<f", 0.75) + b"tail")

    class ModelBin(C.Structure):
        pass
    ModelBin._fields_ = [("pthis", C.c_void_p), ("load_1d", C.CFUNCTYPE(C.c_void_p, C.POINTER(ModelBin), C.c_int, C.c_int)),
                         ("load_2d", C.c_void_p), ("load_3d", C.c_void_p)]
    got = []
    for api in (ours, ref):
        L = api.lib
        rd = capi.MemoryReader(api, stream)
        f = L.ncnn_modelbin_create_from_datareader
        f.restype = C.POINTER(ModelBin)
        f.argtypes = [C.c_void_p]
        mb = f(rd.dr)
        mats = []
        for (w, t) in [(37, 0), (21, 0), (9, 1), (1, 0)]:
            m = C.c_void_p(mb.contents.load_1d(mb, w, t))
            assert m
            mats.append(api.mat_to_numpy(m).copy())
            L.ncnn_mat_destroy(m)
        got.append((mats, rd.pos))
        L.ncnn_modelbin_destroy.argtypes = [C.c_void_p]
        L.ncnn_modelbin_destroy(C.cast(mb, C.c_void_p))
        rd.close()
    for x, y in zip(got[0][0], got[1][0]):
        assert x.shape == y.shape and np.array_equal(x, y)
    assert np.array_equal(got[0][0][1], h.astype(np.float32)) and got[0][0][3][0] == np.float32(0.75)
    assert got[0][1] == got[1][1] == len(stream) - 4


def test_host_border_and_flatten_helpers_match_reference(ours, ref):
    """ncnn_copy_make_border (constant / replicate / reflect), ncnn_copy_cut_border and ncnn_flatten on host Mats of 1 to 3 dims
    (src/c_api.h:188, :413-415), bit-exact against the reference (which runs its Padding / Crop / Flatten layers on the host)"""
    rng = np.random.default_rng(43)
    vp, ci = C.c_void_p, C.c_int
    for api in (ours, ref):
        l = api.lib
        l.ncnn_copy_make_border.restype = None
        l.ncnn_copy_make_border.argtypes = [vp, vp, ci, ci, ci, ci, ci, C.c_float, vp]
        l.ncnn_copy_cut_border.restype = None
        l.ncnn_copy_cut_border.argtypes = [vp, vp, ci, ci, ci, ci, vp]
        l.ncnn_flatten.restype = None
        l.ncnn_flatten.argtypes = [vp, vp, vp]
    for shape in [(19,), (7, 13), (3, 9, 14), (5, 4, 4)]:
        x = rng.uniform(-1, 1, shape).astype(np.float32)
        for (t, b, lft, r) in [(1, 2, 3, 1), (0, 0, 2, 0), (3, 0, 0, 3)]:
            for typ in (0, 1, 2):
                got = []
                for api in (ours, ref):
                    opt = api.strict_fp32_option() if hasattr(api, "strict_fp32_option") else None
                    src = api.mat_from_numpy(x)
                    dst = vp(api.lib.ncnn_mat_create())
                    api.lib.ncnn_copy_make_border(src, dst, t, b, lft, r, typ, 0.5, opt)
                    padded = api.mat_to_numpy(dst).copy()
                    cut = vp(api.lib.ncnn_mat_create())
                    api.lib.ncnn_copy_cut_border(dst, cut, t if len(shape) > 1 else 0, b if len(shape) > 1 else 0, lft, r, opt)
                    back = api.mat_to_numpy(cut).copy()
                    flat = vp()
                    api.lib.ncnn_flatten(dst, C.byref(flat), opt)
                    got.append((padded, back, api.mat_to_numpy(flat).copy()))
                    for m in (src, dst, cut, flat):
                        api.lib.ncnn_mat_destroy(m)
                    if opt:
                        api.lib.ncnn_option_destroy(opt)
                for a, bb in zip(got[0], got[1]):
                    assert a.shape == bb.shape and np.array_equal(a, bb), (shape, (t, b, lft, r), typ)
                assert np.array_equal(got[0][1], x)  # cutting the border off gives the source back
    # channel padding of 3-D Mats (ncnn_copy_make_border_3d)
    for api in (ours, ref):
        api.lib.ncnn_copy_make_border_3d.restype = None
        api.lib.ncnn_copy_make_border_3d.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, ci, C.c_float, vp]
    x = rng.uniform(-1, 1, (4, 6, 9)).astype(np.float32)
    for (t, b, lft, r, f, bh) in [(1, 0, 2, 1, 2, 1), (0, 0, 0, 0, 1, 3), (2, 2, 0, 0, 0, 2)]:
        for typ in (0, 1, 2):
            got = []
            for api in (ours, ref):
                opt = api.strict_fp32_option()
                src = api.mat_from_numpy(x)
                dst = vp(api.lib.ncnn_mat_create())
                api.lib.ncnn_copy_make_border_3d(src, dst, t, b, lft, r, f, bh, typ, -0.25, opt)
                got.append(api.mat_to_numpy(dst).copy())
                api.lib.ncnn_mat_destroy(src)
                api.lib.ncnn_mat_destroy(dst)
                api.lib.ncnn_option_destroy(opt)
            assert got[0].shape == got[1].shape == (4 + f + bh, 6 + t + b, 9 + lft + r) and np.array_equal(got[0], got[1]), ((t, b, lft, r, f, bh), typ)
