"""Kernel-level parity: every compute entry point of include/ncnn_cuda.h against the reference's own naive layer
(oracle/_ref, create_layer_naive -- the ground truth tests/testutil.cpp:1301 of the reference uses) on seeded inputs.

Shape grids follow the reference's tests/test_convolution.cpp:98-163, test_convolutiondepthwise.cpp:69-79,
test_pooling.cpp, test_innerproduct.cpp (reduced), with a batch axis added (the backend is natively batched; the
reference loops per sample, src/net.cpp:654-705).

Tolerances (BASELINE.json north_star), metric = max|a-b| / max|ref| per tensor:
    fp32 CUDA-core path            <= 1e-5
    bf16 / fp16 tensor-core path   <= 2e-3   (inputs and weights pre-rounded to the storage type on both sides)
A 16-bit top blob additionally carries its own storage rounding (unit roundoff 2^-8 for bf16, 2^-11 for fp16, relative
to each element); `nerr(..., elemtype)` discounts exactly that per-element amount before normalising, so the 2e-3 bound
is on the arithmetic (fp32-accumulated tensor-core sums), not on the storage format.
"""
import ctypes as C

import numpy as np
import pytest

import cabi
import geom
from cabi import BF16, F16, F32

pytestmark = pytest.mark.gpu

TOL = {F32: 1e-5, BF16: 2e-3, F16: 2e-3}
ROUNDOFF = {F32: 0.0, BF16: 2.0 ** -8, F16: 2.0 ** -11}


def nerr(a, b, elemtype=F32):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = max(np.abs(b).max(), 1e-30)
    d = np.abs(a - b) - ROUNDOFF[elemtype] * np.abs(b)
    return max(d.max(), 0.0) / den


def quant(a, elemtype):
    """round to the storage type so that the oracle sees exactly the values the device blob holds"""
    import torch
    if elemtype == F32:
        return np.asarray(a, np.float32)
    return torch.from_numpy(np.asarray(a, np.float32)).to(cabi.torch_dtype(elemtype)).float().numpy()


def rand(rng, shape, lo=-1.0, hi=1.0):
    return rng.uniform(lo, hi, shape).astype(np.float32)


def sync():
    import torch
    torch.cuda.synchronize()


ACT_PARAMS = {0: [], 1: [], 2: [0.1], 3: [-0.5, 0.8], 4: [], 5: [], 6: [0.2, 0.5]}


def act_of(t):
    p = ACT_PARAMS[t] + [0.0, 0.0]
    return cabi.act(t, p[0], p[1])


# ------------------------------------------------------------------------------------------ Convolution
def run_conv(ref, rng, elemtype, n, w, h, inch, outch, k, d, s, pad, bias, act_type, pad_value=0.0, kh=None, residual=False, expect_algo=None, expect_workspace=None):
    L = cabi.lib()
    kw = k
    kh = kh or k
    x = quant(rand(rng, (n, inch, h, w)), elemtype)
    wt = quant(rand(rng, (outch, inch, kh, kw)) * np.float32(np.sqrt(3.0 / (inch * kw * kh))), elemtype)
    b = rand(rng, (outch,)) if bias else None
    params = {0: outch, 1: kw, 11: kh, 2: d, 3: s, 4: pad, 5: int(bias), 6: wt.size, 9: act_type, 18: float(pad_value)}
    if ACT_PARAMS[act_type]:
        params[10] = np.asarray(ACT_PARAMS[act_type], np.float32)
    if residual:
        params[9] = 0
    want = ref.layer_forward("Convolution", params, [wt] + ([b] if bias else []), [x], batched=True)[0]

    pads = geom.conv_pads(w, h, kw, kh, d, d, s, s, pad, pad, pad, pad)
    outw, outh = geom.conv_out(w, h, kw, kh, d, d, s, s, pads)
    assert want.shape == (n, outch, outh, outw)
    res_np = None
    if residual:
        res_np = quant(rand(rng, want.shape), elemtype)
        want = want + res_np
        if act_type == 1:
            want = np.maximum(want, 0)

    desc = cabi.ConvDesc(inch, outch, kw, kh, d, d, s, s, pads[0], pads[1], pads[2], pads[3], pad_value, int(bias), act_of(act_type), elemtype)
    handle = C.c_void_p()
    wa, wp = cabi.fptr(wt)
    if bias:
        ba, bp = cabi.fptr(b)
    else:
        bp = None
    cabi.check(L.ncnn_cuda_conv2d_create(C.byref(handle), C.byref(desc), wp, bp, None), "conv2d_create")
    bottom = cabi.Blob.from_numpy(x, elemtype)
    top = cabi.Blob((outch, outh, outw), n, elemtype, fill=float("nan"))
    bd, td = bottom.desc(), top.desc()
    rd = None
    if residual:
        resb = cabi.Blob.from_numpy(res_np, elemtype)
        rd = resb.desc()
    if expect_algo is not None:
        assert L.ncnn_cuda_conv2d_algo(handle, C.byref(bd)) == expect_algo
    # scratch for the small-channel stem variant (A_ROWS); 0 for every other geometry
    import torch
    wsize = 0 if residual else int(L.ncnn_cuda_conv2d_workspace_size(handle, C.byref(bd), C.byref(td)))
    if expect_workspace is not None:
        assert (wsize > 0) == expect_workspace, "workspace %d" % wsize
    ws = torch.full(((wsize + 3) // 4,), float("nan"), dtype=torch.float32, device="cuda") if wsize else None
    cabi.check(L.ncnn_cuda_conv2d_forward(handle, C.byref(bd), C.byref(td), pads[0], pads[2], C.byref(rd) if rd else None, None,
                                          C.c_void_p(ws.data_ptr()) if wsize else None, C.c_size_t(wsize), None), "conv2d_forward")
    sync()
    got = top.numpy()
    L.ncnn_cuda_conv2d_destroy(handle)
    e = nerr(got, want, elemtype)
    assert np.isfinite(got).all(), "non-finite output"
    assert e <= TOL[elemtype], "conv n%d %dx%dx%d->%d k%dx%d d%d s%d p%d act%d: err %.3g" % (n, w, h, inch, outch, kw, kh, d, s, pad, act_type, e)
    return e


CONV_GRID = [
    # k, d, s, pad   (tests/test_convolution.cpp:100-117 of the reference)
    (1, 1, 1, 0), (1, 1, 2, 0), (2, 1, 1, 1), (2, 1, 2, -233), (3, 1, 1, 1), (3, 1, 2, 1), (3, 2, 1, -234), (4, 1, 1, 2),
    (4, 1, 2, -233), (4, 2, 1, -234), (5, 1, 1, -233), (5, 1, 2, 2), (5, 2, 2, 2), (7, 1, 1, 3), (7, 1, 2, 3), (7, 2, 1, -233),
]


@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_convolution_grid(ref, elemtype):
    rng = np.random.default_rng(7767517)
    chans = [(1, 1), (4, 13), (13, 4), (12, 12), (8, 12), (16, 24), (15, 15), (16, 16), (3, 64), (40, 72)]
    i = 0
    for (k, d, s, pad) in CONV_GRID:
        for (ci, co) in chans[i % 3::3]:
            run_conv(ref, rng, elemtype, n=1 + i % 3, w=9 + i % 5, h=7 + i % 4, inch=ci, outch=co, k=k, d=d, s=s, pad=pad, bias=i % 2 == 0, act_type=i % 7)
            i += 1


@pytest.mark.parametrize("elemtype", [BF16, F16])
def test_convolution_tensor_core_shapes(ref, elemtype):
    """shapes of the named models' layers (reduced spatial size / batch), all on the tcgen05 path"""
    rng = np.random.default_rng(1)
    cases = [
        # w, h, cin, cout, k, s, pad, n
        (56, 56, 64, 256, 1, 1, 0, 2),     # resnet50 1x1 (TILED)
        (28, 28, 256, 64, 1, 1, 0, 3),     # M tail: 3*784 = 2352 = 18.4 tiles
        (30, 30, 64, 64, 3, 1, 1, 2),      # 3x3 s1 p1 (IM2COL)
        (56, 56, 256, 512, 1, 2, 0, 2),    # strided 1x1 projection (IM2COL)
        (64, 64, 3, 64, 7, 2, 3, 2),       # stem
        (33, 31, 3, 32, 3, 2, 1, 3),       # mobilenet / yolo stem, odd sizes
        (14, 14, 512, 512, 3, 1, 1, 2),    # vgg-like, K = 4608
        (13, 13, 512, 1000, 1, 1, 1, 2),   # squeezenet conv10: 1x1 with pad 1
        (20, 20, 320, 96, 1, 1, 0, 2),     # non-multiple-of-64 channels
        (17, 19, 24, 144, 3, 2, -233, 2),
    ]
    for (w, h, ci, co, k, s, pad, n) in cases:
        run_conv(ref, rng, elemtype, n, w, h, ci, co, k, 1, s, pad, True, 1, expect_algo=2)


@pytest.mark.parametrize("elemtype", [BF16, F16])
def test_convolution_shifted_window_shapes(ref, elemtype):
    """stride-1 k x k convolutions with shared-memory-resident weights run the A_SHIFT variant (tc_gemm.cuh): one staged
    pixel buffer per 64-channel slab, taps as row-shifted UMMA descriptors.  Rows narrower and wider than the 128-row
    MMA tile, several rows per tile, column chunks, image/row-group tails, channel counts that are not a multiple of 64,
    asymmetric padding, 'valid' windows, 5x5 and non-square kernels"""
    rng = np.random.default_rng(6)
    cases = [
        # w, h, cin, cout, k, pad, n, act
        (56, 56, 64, 64, 3, 1, 2, 1),      # resnet50 res2x_branch2b: 2 rows of 58 per tile
        (28, 28, 64, 64, 3, 1, 3, 0),      # 4 rows of 30 per tile
        (14, 14, 64, 48, 3, 1, 5, 1),      # 8 rows of 16 per tile, 14 = 8 + 6 (row-group tail)
        (224, 20, 64, 64, 3, 1, 1, 1),     # vgg16 conv1_2: 226 > 128 -> column chunks of 126
        (130, 9, 64, 32, 3, 1, 2, 3),      # 132-wide rows: two chunks, second nearly empty
        (126, 7, 40, 64, 3, 1, 2, 1),      # exactly one 128-wide buffer row; cin = 40 (zero-filled channel tail)
        (61, 33, 96, 32, 3, 1, 2, 2),      # two 64-channel slabs (the second half empty)
        (40, 40, 64, 64, 3, 0, 2, 1),      # valid window: outw = 38
        (37, 29, 48, 24, 5, 2, 2, 7),      # 5x5, swish epilogue
        (45, 31, 64, 64, 3, -233, 2, 1),   # SAME padding
        (50, 18, 64, 16, 2, 1, 2, 0),      # even kernel: outw = 51
    ]
    for (w, h, ci, co, k, pad, n, act) in cases:
        if act == 7:
            continue  # (swish is a graph-level fold, not a layer activation_type: covered by the network tests)
        run_conv(ref, rng, elemtype, n, w, h, ci, co, k, 1, 1, pad, True, act, expect_algo=2)
    run_conv(ref, rng, elemtype, 2, 41, 37, 64, 48, 5, 1, 1, 2, True, 2, kh=3)
    run_conv(ref, rng, elemtype, 30, 56, 56, 64, 64, 3, 1, 1, 1, True, 1)  # more tiles than CTAs x stages


@pytest.mark.parametrize("elemtype", [BF16, F16])
def test_convolution_small_channel_stems(ref, elemtype):
    """cin <= 8 first layers run the A_ROWS variant (zero-padded 4/8-channel copy + overlapping-window TMA): the five
    models' stems, odd sizes, rows wider than one 128-column chunk, and the cases that must NOT take it"""
    rng = np.random.default_rng(5)
    cases = [
        # w, h, cin, cout, k, s, pad, n, takes A_ROWS
        (64, 64, 3, 64, 7, 2, 3, 2, True),       # resnet50 conv1
        (225, 57, 3, 64, 7, 2, 3, 1, True),      # odd width, outw = 113
        (33, 31, 3, 32, 3, 2, 1, 3, True),       # mobilenet_v2 / yolov8 stem, odd sizes
        (640, 24, 3, 32, 3, 2, 1, 1, True),      # yolov8 stem width: outw = 320 -> 3 column chunks
        (67, 35, 3, 64, 3, 2, 0, 2, True),       # squeezenet conv1 (no padding)
        (48, 20, 3, 64, 3, 1, 1, 2, True),       # vgg16 conv1_1: stride 1 -> 8-channel pixels
        (300, 9, 3, 64, 3, 1, 1, 1, True),       # stride 1, outw = 300
        (40, 40, 1, 16, 5, 2, 2, 2, True),
        (40, 40, 4, 24, 3, 2, 1, 2, True),
        (30, 30, 8, 24, 3, 1, 1, 2, True),       # cin = 8
        (30, 30, 3, 24, 3, 2, -233, 2, True),    # SAME padding, resolved by the caller before create (odd left pad -> shift)
        (30, 30, 12, 24, 3, 2, 1, 2, False),
    ]
    for (w, h, ci, co, k, s, pad, n, rows) in cases:
        run_conv(ref, rng, elemtype, n, w, h, ci, co, k, 1, s, pad, True, 1, expect_algo=2, expect_workspace=rows)
    run_conv(ref, rng, elemtype, 2, 41, 37, 3, 48, 7, 1, 2, 3, True, 0, kh=3, expect_workspace=True)
    run_conv(ref, rng, elemtype, 2, 41, 37, 3, 48, 3, 1, 2, 1, True, 2, kh=5, expect_workspace=True)


def test_convolution_pad_value_and_kernel_wh(ref):
    rng = np.random.default_rng(2)
    run_conv(ref, rng, F32, 2, 11, 9, 5, 7, 3, 1, 1, 1, True, 0, pad_value=0.7)
    run_conv(ref, rng, BF16, 2, 11, 9, 16, 16, 3, 1, 1, 1, True, 0, pad_value=0.5)  # non-zero pad value: falls back to the SIMT kernel
    run_conv(ref, rng, F32, 2, 11, 9, 5, 7, 3, 1, 2, 1, True, 1, kh=5)
    run_conv(ref, rng, BF16, 2, 12, 10, 32, 48, 5, 1, 1, 2, True, 1, kh=3)


@pytest.mark.parametrize("elemtype", [F32, BF16])
def test_convolution_fused_residual(ref, elemtype):
    rng = np.random.default_rng(3)
    run_conv(ref, rng, elemtype, 2, 14, 14, 64, 256, 1, 1, 1, 0, True, 1, residual=True)
    run_conv(ref, rng, elemtype, 3, 9, 9, 24, 40, 3, 1, 1, 1, True, 0, residual=True)


# ------------------------------------------------------------------------------------------ ConvolutionDepthWise
def run_dw(ref, rng, elemtype, n, w, h, ch, outch, group, k, d, s, pad, bias, act_type):
    L = cabi.lib()
    x = quant(rand(rng, (n, ch, h, w)), elemtype)
    wt = rand(rng, (outch, ch // group, k, k)) * np.float32(0.5)
    b = rand(rng, (outch,)) if bias else None
    params = {0: outch, 1: k, 2: d, 3: s, 4: pad, 5: int(bias), 6: wt.size, 7: group, 9: act_type}
    if ACT_PARAMS[act_type]:
        params[10] = np.asarray(ACT_PARAMS[act_type], np.float32)
    want = ref.layer_forward("ConvolutionDepthWise", params, [wt] + ([b] if bias else []), [x], batched=True)[0]
    pads = geom.conv_pads(w, h, k, k, d, d, s, s, pad, pad, pad, pad)
    outw, outh = geom.conv_out(w, h, k, k, d, d, s, s, pads)
    desc = cabi.DwConvDesc(ch, outch, group, k, k, d, d, s, s, 0.0, int(bias), act_of(act_type), elemtype)
    handle = C.c_void_p()
    wa, wp = cabi.fptr(wt)
    if bias:
        ba, bp = cabi.fptr(b)
    else:
        bp = None
    cabi.check(L.ncnn_cuda_dwconv2d_create(C.byref(handle), C.byref(desc), wp, bp, None), "dwconv2d_create")
    bottom = cabi.Blob.from_numpy(x, elemtype, pad_fill=0.0)
    top = cabi.Blob((outch, outh, outw), n, elemtype, fill=float("nan"))
    bd, td = bottom.desc(), top.desc()
    cabi.check(L.ncnn_cuda_dwconv2d_forward(handle, C.byref(bd), C.byref(td), pads[0], pads[2], None), "dwconv2d_forward")
    sync()
    got = top.numpy()
    L.ncnn_cuda_dwconv2d_destroy(handle)
    tol = 1e-5 if elemtype == F32 else 1e-4  # 16-bit: fp32 weights and accumulation, only the output store rounds
    e = nerr(got, want, elemtype)
    assert e <= tol, "dw n%d %dx%dx%d g%d k%d d%d s%d p%d: err %.3g" % (n, w, h, ch, group, k, d, s, pad, e)


@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_convolutiondepthwise_grid(ref, elemtype):
    rng = np.random.default_rng(11)
    i = 0
    for (k, d, s, pad) in [(1, 1, 1, 0), (2, 1, 2, -233), (3, 1, 1, 1), (3, 1, 2, 1), (3, 2, 1, -234), (5, 1, 2, 2), (7, 2, 1, -233)]:
        for (ch, outch, group) in [(1, 1, 1), (2, 2, 2), (3, 3, 3), (7, 7, 7), (8, 8, 8), (12, 12, 12), (15, 15, 15), (16, 16, 16), (32, 32, 32), (96, 96, 96),
                                   (4, 2, 2), (6, 6, 2), (12, 24, 4)]:
            if i % 2 == 0 or group != ch:
                run_dw(ref, rng, elemtype, 1 + i % 3, 11 + i % 6, 9 + i % 3, ch, outch, group, k, d, s, pad, i % 2 == 0, i % 7)
            i += 1


@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_convolutiondepthwise_mobilenet_shapes(ref, elemtype):
    """the 3x3 stride-1/2 depthwise shapes of MobileNetV2 (TMA-staged halo-tile kernel, dwconv_tma.cuh): every channel-block
    width (C % 64, % 32, % 16), odd and even map sizes, maps smaller than a tile, ReLU and ReLU6 (Clip) epilogues"""
    rng = np.random.default_rng(12)
    for (w, c, s, act) in [(56, 32, 1, 1), (56, 96, 2, 3), (28, 144, 1, 3), (29, 144, 2, 0), (14, 384, 1, 3), (14, 576, 2, 1), (7, 960, 1, 3), (15, 192, 2, 3),
                           (33, 48, 1, 2), (9, 64, 2, 4)]:
        run_dw(ref, rng, elemtype, 2, w, w + (w % 3), c, c, c, 3, 1, s, 1, True, act)
    # maps of 7k rows (one column x seven rows per thread, wide and possibly partial channel blocks)
    for (w, h, c, act) in [(7, 7, 960, 3), (14, 14, 576, 3), (28, 28, 192, 1), (14, 14, 384, 0), (13, 7, 208, 3), (30, 14, 72, 1), (8, 21, 1040, 3)]:
        run_dw(ref, rng, elemtype, 3, w, h, c, c, c, 3, 1, 1, 1, True, act)


def test_convolutiondepthwise_many_tiles(ref):
    """more tiles than 4 ring stages x 148 persistent CTAs: the stage/phase wrap of the TMA ring and the weight reload
    when a CTA moves to the next channel block"""
    rng = np.random.default_rng(13)
    run_dw(ref, rng, F16, 16, 112, 112, 32, 32, 32, 3, 1, 1, 1, True, 3)
    run_dw(ref, rng, BF16, 24, 56, 56, 96, 96, 96, 3, 1, 2, 1, False, 1)
    run_dw(ref, rng, F16, 40, 14, 14, 384, 384, 384, 3, 1, 1, 1, True, 3)


# ------------------------------------------------------------------------------------------ Deconvolution
def deconv_cut(w, h, k, d, s, pad, opr, opb, output_w, output_h):
    """bordered size and the cut of the reference (src/layer/deconvolution.cpp:156-157 and cut_padding :364-392) ->
    (outw, outh, cut_left, cut_top)"""
    ext = d * (k - 1) + 1
    fw, fh = (w - 1) * s + ext + opr, (h - 1) * s + ext + opb
    if pad > 0:
        return fw - 2 * pad, fh - 2 * pad, pad, pad
    if output_w > 0 and output_h > 0:
        wcut, hcut = fw - output_w, fh - output_h
        if pad == -233:
            return output_w, output_h, wcut // 2, hcut // 2
        if pad == -234:
            return output_w, output_h, wcut - wcut // 2, hcut - hcut // 2
    return fw, fh, 0, 0


def run_deconv(ref, rng, elemtype, n, w, h, ch, outch, group, k, d, s, pad, bias, act_type, opr=0, opb=0, output_w=0, output_h=0):
    L = cabi.lib()
    if output_w > 0 and output_h > 0 and pad not in (-233, -234):
        pad = -233  # as tests/test_deconvolution.cpp:10-13 of the reference
    x = quant(rand(rng, (n, ch, h, w)), elemtype)
    wt = rand(rng, (outch, ch // group, k, k)) * np.float32(0.5)
    b = rand(rng, (outch,)) if bias else None
    params = {0: outch, 1: k, 2: d, 3: s, 4: pad, 5: int(bias), 6: wt.size, 9: act_type, 18: opr, 19: opb, 20: output_w, 21: output_h}
    if group != 1:
        params[7] = group
    if ACT_PARAMS[act_type]:
        params[10] = np.asarray(ACT_PARAMS[act_type], np.float32)
    want = ref.layer_forward("Deconvolution" if group == 1 else "DeconvolutionDepthWise", params, [wt] + ([b] if bias else []), [x], batched=True)[0]
    outw, outh, cut_left, cut_top = deconv_cut(w, h, k, d, s, pad, opr, opb, output_w, output_h)
    assert want.shape == (n, outch, outh, outw), (want.shape, (n, outch, outh, outw))
    desc = cabi.DeconvDesc(ch, outch, group, k, k, d, d, s, s, opr, opb, int(bias), act_of(act_type), elemtype)
    handle = C.c_void_p()
    wa, wp = cabi.fptr(wt)
    if bias:
        ba, bp = cabi.fptr(b)
    else:
        bp = None
    cabi.check(L.ncnn_cuda_deconv2d_create(C.byref(handle), C.byref(desc), wp, bp, None), "deconv2d_create")
    bottom = cabi.Blob.from_numpy(x, elemtype)
    top = cabi.Blob((outch, outh, outw), n, elemtype, fill=float("nan"))
    bd, td = bottom.desc(), top.desc()
    cabi.check(L.ncnn_cuda_deconv2d_forward(handle, C.byref(bd), C.byref(td), cut_left, cut_top, None), "deconv2d_forward")
    sync()
    got = top.numpy()
    L.ncnn_cuda_deconv2d_destroy(handle)
    tol = 1e-5 if elemtype == F32 else 1e-4  # 16-bit: fp32 weights and accumulation, only the output store rounds
    e = nerr(got, want, elemtype)
    assert e <= tol, "deconv n%d %dx%dx%d->%d g%d k%d d%d s%d p%d op%d,%d out%dx%d: err %.3g" % (n, w, h, ch, outch, group, k, d, s, pad, opr, opb, output_w, output_h, e)


DECONV_KDSP = [(1, 1, 1, 0), (1, 1, 2, 0), (2, 1, 1, 1), (2, 1, 2, -233), (3, 1, 1, 1), (3, 1, 2, 1), (3, 2, 1, 1), (4, 1, 1, -233), (4, 1, 2, -234),
               (4, 2, 1, -234), (5, 1, 1, 2), (5, 1, 2, 2), (5, 2, 2, 2), (7, 1, 1, 3), (7, 1, 2, 3), (7, 2, 1, -233)]


@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_deconvolution_grid(ref, elemtype):
    """the kernel / dilation / stride / pad table and the output_pad / output_w,h cases of the reference's
    tests/test_deconvolution.cpp:97-131, with a batch axis, against its naive layer"""
    rng = np.random.default_rng(71)
    i = 0
    for (k, d, s, pad) in DECONV_KDSP:
        for (w, h, c, outch, bias, opr, opb, ow, oh) in [(9, 7, 1, 1, 1, 0, 0, 0, 0), (9, 7, 4, 13, 0, 1, 1, 7, 5), (9, 7, 13, 4, 1, 1, 0, 0, 0),
                                                         (9, 7, 8, 4, 1, 0, 0, 7, 5), (7, 7, 12, 12, 1, 0, 1, 0, 0), (4, 5, 12, 11, 0, 0, 1, 1, 0),
                                                         (9, 7, 8, 13, 0, 2, 2, 0, 0), (9, 7, 16, 16, 0, 0, 2, 7, 5)]:
            if pad > 0 and ((w - 1) * s + d * (k - 1) + 1 + opr - 2 * pad <= 0 or (h - 1) * s + d * (k - 1) + 1 + opb - 2 * pad <= 0):
                continue
            if elemtype == F32 or i % 2 == 0:
                run_deconv(ref, rng, elemtype, 1 + i % 3, w, h, c, outch, 1, k, d, s, pad, bool(bias), i % 7, opr, opb, ow, oh)
            i += 1
    # wide channel counts (the 4x4 stride-2 upsampling heads of detection / segmentation graphs)
    for (c, outch) in [(24, 32), (32, 28), (64, 64)]:
        run_deconv(ref, rng, elemtype, 2, 7, 5, c, outch, 1, 4, 1, 2, 1, True, 1)


@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_deconvolutiondepthwise_grid(ref, elemtype):
    """depthwise and grouped branches (src/layer/deconvolutiondepthwise.cpp:103-205); grid of the reference's
    tests/test_deconvolutiondepthwise.cpp, with a batch axis"""
    rng = np.random.default_rng(72)
    i = 0
    for (k, d, s, pad) in DECONV_KDSP:
        for (w, h, c, outch, group, bias, opr, opb, ow, oh) in [(15, 7, 1, 1, 1, 1, 0, 0, 0, 0), (15, 7, 2, 2, 2, 0, 1, 1, 7, 5), (15, 7, 3, 3, 3, 1, 0, 1, 0, 0),
                                                                (15, 7, 4, 2, 2, 0, 0, 0, 7, 5), (15, 7, 7, 7, 7, 1, 2, 2, 0, 0), (15, 7, 8, 8, 2, 1, 0, 0, 7, 5),
                                                                (15, 7, 12, 12, 4, 0, 2, 2, 0, 0), (15, 7, 16, 32, 8, 1, 2, 0, 0, 0), (15, 7, 64, 64, 64, 1, 0, 0, 0, 0)]:
            if group == 1:
                continue
            if pad > 0 and ((h - 1) * s + d * (k - 1) + 1 + opb - 2 * pad <= 0):
                continue
            if elemtype == F32 or i % 2 == 1:
                run_deconv(ref, rng, elemtype, 1 + i % 2, w, h, c, outch, group, k, d, s, pad, bool(bias), i % 7, opr, opb, ow, oh)
            i += 1
    # the 2x depthwise upsampling of mobilenet_yolo / mobilenetv2_yolov3 (benchmark/models): k2 s2 on 13x13 maps
    run_deconv(ref, rng, elemtype, 3, 13, 13, 96, 96, 96, 2, 1, 2, 0, False, 0)


# ------------------------------------------------------------------------------------------ Pooling
def run_pool(ref, rng, elemtype, n, w, h, c, ptype, k, s, pad, pad_mode, global_pool=0, include_pad=0, adaptive=0, out_wh=(0, 0)):
    L = cabi.lib()
    x = quant(rand(rng, (n, c, h, w)), elemtype)
    params = {0: ptype, 1: k, 2: s, 3: pad, 4: global_pool, 5: pad_mode, 6: include_pad, 7: adaptive, 8: out_wh[0], 18: out_wh[1]}
    want = ref.layer_forward("Pooling", params, [], [x], batched=True)[0]
    if global_pool:
        g = dict(outw=1, outh=1, pad_left=0, pad_top=0, area=(0, w, 0, h))
        top = cabi.Blob((c,), n, elemtype, fill=float("nan"))
    elif adaptive:
        g = dict(outw=out_wh[0], outh=out_wh[1], pad_left=0, pad_top=0, area=(0, w, 0, h))
        top = cabi.Blob((c, g["outh"], g["outw"]), n, elemtype, fill=float("nan"))
    else:
        g = geom.pool_geometry(w, h, k, k, s, s, pad, pad, pad, pad, pad_mode)
        top = cabi.Blob((c, g["outh"], g["outw"]), n, elemtype, fill=float("nan"))
    desc = cabi.PoolDesc(ptype, k, k, s, s, g["pad_left"], g["pad_top"], global_pool, include_pad, adaptive, *g["area"])
    bottom = cabi.Blob.from_numpy(x, elemtype, pad_fill=0.0)
    bd, td = bottom.desc(), top.desc()
    cabi.check(L.ncnn_cuda_pool2d_forward(C.byref(desc), C.byref(bd), C.byref(td), None), "pool2d_forward")
    sync()
    got = top.numpy()
    assert got.shape == want.shape, (got.shape, want.shape)
    # degenerate windows that lie entirely in padding: the reference yields 0/0 = NaN (avg) or -FLT_MAX (max, which
    # a 16-bit blob stores as -inf); both sides must agree on where they are, and they are left out of the error
    deg = ~np.isfinite(want) | (np.abs(want) >= 3e38)
    assert (deg == (~np.isfinite(got) | (np.abs(got) >= 3e38))).all()
    got = np.where(deg, 0, got)
    want = np.where(deg, 0, want)
    tol = 1e-6 if (elemtype == F32) else 1e-5
    if ptype == 0 and elemtype != F32:
        tol = 0.0  # max of stored values is exact
    e = nerr(got, want, elemtype if ptype == 1 else F32)
    assert e <= tol, "pool n%d %dx%dx%d type%d k%d s%d p%d mode%d g%d inc%d ad%d: err %.3g" % (n, w, h, c, ptype, k, s, pad, pad_mode, global_pool, include_pad, adaptive, e)


@pytest.mark.parametrize("elemtype", [F32, BF16])
def test_pooling_grid(ref, elemtype):
    rng = np.random.default_rng(21)
    i = 0
    for ptype in (0, 1):
        for (k, s, pad) in [(2, 1, 0), (2, 2, 1), (3, 1, 0), (3, 2, 1), (4, 2, 1), (5, 1, 2), (5, 2, 2), (7, 3, 1)]:
            for pad_mode in (0, 1, 2, 3):
                c = [1, 3, 4, 8, 12, 16, 31, 64][i % 8]
                run_pool(ref, rng, elemtype, 1 + i % 3, 13 + i % 5, 11 + i % 4, c, ptype, k, s, pad if pad_mode < 2 else 0, pad_mode, include_pad=(i // 3) % 2)
                i += 1
    for ptype in (0, 1):
        run_pool(ref, rng, elemtype, 2, 7, 7, 2048, ptype, 0, 1, 0, 0, global_pool=1)
        run_pool(ref, rng, elemtype, 3, 13, 9, 10, ptype, 0, 1, 0, 0, global_pool=1)
        run_pool(ref, rng, elemtype, 2, 13, 9, 12, ptype, 0, 1, 0, 0, adaptive=1, out_wh=(4, 3))
    # the named models' windows: resnet 3x3 s2 pad_mode 0 (tail pad), avg 7x7 s1, vgg 2x2 s2, sppf 5x5 s1 p2 valid
    run_pool(ref, rng, elemtype, 2, 112, 112, 64, 0, 3, 2, 0, 0)
    run_pool(ref, rng, elemtype, 2, 7, 7, 256, 1, 7, 1, 0, 0)
    run_pool(ref, rng, elemtype, 2, 28, 28, 32, 0, 2, 2, 0, 0)
    run_pool(ref, rng, elemtype, 2, 20, 20, 256, 0, 5, 1, 2, 1)


@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_pooling_max_tma_shapes(ref, elemtype):
    """max pooling on the TMA-staged tile kernel (pool_tma.cuh): every channel-block width, maps that are not a multiple of
    the tile, windows hanging over every edge (NaN out-of-bounds fill must never win), more tiles than ring stages"""
    rng = np.random.default_rng(22)
    for (w, h, c, k, s, pad, mode) in [(112, 112, 64, 3, 2, 0, 0), (57, 33, 128, 2, 2, 0, 0), (56, 56, 96, 3, 2, 1, 0), (27, 27, 256, 3, 2, 0, 0),
                                       (20, 20, 256, 5, 1, 2, 1), (40, 23, 48, 5, 1, 2, 1), (14, 14, 512, 2, 2, 0, 0), (31, 29, 16, 3, 1, 1, 1),
                                       (13, 13, 64, 3, 2, 0, 2), (15, 15, 32, 2, 1, 0, 3)]:
        run_pool(ref, rng, elemtype, 3, w, h, c, 0, k, s, pad, mode)
    run_pool(ref, rng, elemtype, 24, 112, 112, 64, 0, 3, 2, 0, 0)


# ------------------------------------------------------------------------------------------ InnerProduct
def run_fc(ref, rng, elemtype, n, in_shape, num_output, bias, act_type):
    L = cabi.lib()
    x = quant(rand(rng, (n,) + tuple(in_shape)), elemtype)
    num_input = int(np.prod(in_shape)) if len(in_shape) != 2 else in_shape[1]
    wt = quant(rand(rng, (num_output, num_input)) * np.float32(np.sqrt(3.0 / num_input)), elemtype)
    b = rand(rng, (num_output,)) if bias else None
    params = {0: num_output, 1: int(bias), 2: wt.size, 9: act_type}
    if ACT_PARAMS[act_type]:
        params[10] = np.asarray(ACT_PARAMS[act_type], np.float32)
    want = ref.layer_forward("InnerProduct", params, [wt] + ([b] if bias else []), [x], batched=True)[0]
    if len(in_shape) == 3:
        desc = cabi.LinearDesc(num_input, num_output, int(bias), act_of(act_type), elemtype, in_shape[2], in_shape[1], in_shape[0])
    else:
        desc = cabi.LinearDesc(num_input, num_output, int(bias), act_of(act_type), elemtype, 0, 0, 0)
    handle = C.c_void_p()
    wa, wp = cabi.fptr(wt)
    if bias:
        ba, bp = cabi.fptr(b)
    else:
        bp = None
    cabi.check(L.ncnn_cuda_linear_create(C.byref(handle), C.byref(desc), wp, bp, None), "linear_create")
    bottom = cabi.Blob.from_numpy(x, elemtype)
    top_shape = (num_output,) if len(in_shape) != 2 else (in_shape[0], num_output)
    top = cabi.Blob(top_shape, n, elemtype, fill=float("nan"))
    bd, td = bottom.desc(), top.desc()
    cabi.check(L.ncnn_cuda_linear_forward(handle, C.byref(bd), C.byref(td), None), "linear_forward")
    sync()
    got = top.numpy()
    L.ncnn_cuda_linear_destroy(handle)
    e = nerr(got, want, elemtype)
    assert e <= TOL[elemtype], "fc n%d %s->%d: err %.3g" % (n, in_shape, num_output, e)


@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_innerproduct(ref, elemtype):
    rng = np.random.default_rng(31)
    run_fc(ref, rng, elemtype, 3, (2048,), 1000, True, 0)
    run_fc(ref, rng, elemtype, 2, (64, 7, 7), 96, True, 1)      # 3-D bottom: reference flattens c-major (innerproduct.cpp:141-162)
    run_fc(ref, rng, elemtype, 2, (5, 24), 13, True, 4)         # 2-D bottom: row-wise gemm (innerproduct.cpp:102-134)
    run_fc(ref, rng, elemtype, 1, (15,), 7, False, 2)
    run_fc(ref, rng, elemtype, 130, (1280,), 1000, True, 0)


# ------------------------------------------------------------------------------------------ Reshape / Permute
@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_reshape_permute_layout(elemtype):
    """Reshape and Permute are pure data movement; the reference works on the planar (c,h,w) order
    (src/layer/reshape.cpp, permute.cpp:38-73), so numpy's reshape / transpose of the planar array IS the reference
    result.  Covers the tiled-transpose fast paths (3-D -> 2-D reshape, 2-D permute: YOLOv8's head) and the generic gather."""
    L = cabi.lib()
    L.ncnn_cuda_reshape.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ncnn_cuda_permute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    rng = np.random.default_rng(41)
    for (n, c, h, w) in [(2, 144, 20, 20), (3, 65, 7, 9), (1, 8, 33, 31), (2, 33, 1, 70)]:
        x = quant(rand(rng, (n, c, h, w)), elemtype)
        src = cabi.Blob.from_numpy(x, elemtype)
        # (w,h,c) -> (w*h, c): planar reshape to (c, h*w)
        dst = cabi.Blob((c, h * w), n, elemtype, fill=float("nan"))
        sd, dd = src.desc(), dst.desc()
        cabi.check(L.ncnn_cuda_reshape(C.byref(sd), C.byref(dd), None), "reshape 3d->2d")
        sync()
        assert np.array_equal(dst.numpy(), x.reshape(n, c, h * w))
        # 2-D permute order 1: (w,h) -> (h,w)
        per = cabi.Blob((h * w, c), n, elemtype, fill=float("nan"))
        pd = per.desc()
        cabi.check(L.ncnn_cuda_permute(C.byref(dd), C.byref(pd), 1, None), "permute 2d")
        sync()
        assert np.array_equal(per.numpy(), x.reshape(n, c, h * w).transpose(0, 2, 1))
        # back: 2-D (w*h, c) -> 3-D
        back = cabi.Blob((c, h, w), n, elemtype, fill=float("nan"))
        bd = back.desc()
        cabi.check(L.ncnn_cuda_reshape(C.byref(dd), C.byref(bd), None), "reshape 2d->3d")
        sync()
        assert np.array_equal(back.numpy(), x)
        # a reshape the fast path must decline: (w,h,c) -> (w*c, h) style regrouping goes through the generic gather
        if (c * h) % 2 == 0:
            g = cabi.Blob((2, c * h // 2, w), n, elemtype, fill=float("nan"))
            gd = g.desc()
            cabi.check(L.ncnn_cuda_reshape(C.byref(sd), C.byref(gd), None), "reshape generic")
            sync()
            assert np.array_equal(g.numpy(), x.reshape(n, 2, c * h // 2, w))


# ------------------------------------------------------------------------------------------ BatchNorm / Scale / ShuffleChannel
@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_batchnorm_scale_shufflechannel(ref, elemtype):
    """per-channel affine (BatchNorm: b*x + a, Scale: x*s + bias) and the ShuffleChannel permutation against the reference's
    naive layers (src/layer/batchnorm.cpp, scale.cpp, shufflechannel.cpp); 1-D, 2-D (per-row) and 3-D blobs, channel counts
    that are not a multiple of the vector width"""
    import torch
    L = cabi.lib()
    L.ncnn_cuda_channel_affine.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ncnn_cuda_shuffle_channel.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    rng = np.random.default_rng(51)
    tol = 1e-5 if elemtype == F32 else 1e-4
    for shape in [(24, 9, 11), (7, 5, 6), (64, 14, 14), (13, 40), (37,)]:
        n = 3
        ch = shape[0]
        x = quant(rand(rng, (n,) + shape), elemtype)
        slope, mean = rand(rng, (ch,), 0.5, 1.5), rand(rng, (ch,), -0.3, 0.3)
        var, bias = rand(rng, (ch,), 0.4, 1.6), rand(rng, (ch,), -0.2, 0.2)
        eps = np.float32(1e-5)
        # BatchNorm
        want = ref.layer_forward("BatchNorm", {0: ch, 1: float(eps)}, [slope, mean, var, bias], [x], batched=True)[0]
        sq = np.sqrt(var + eps, dtype=np.float32)
        a = (bias - slope * mean / sq).astype(np.float32)
        b = (slope / sq).astype(np.float32)
        ad, bdv = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        blob = cabi.Blob.from_numpy(x, elemtype, pad_fill=0.0)
        d = blob.desc()
        cabi.check(L.ncnn_cuda_channel_affine(C.byref(d), C.byref(d), C.c_void_p(bdv.data_ptr()), C.c_void_p(ad.data_ptr()), None), "batchnorm")
        sync()
        assert nerr(blob.numpy(), want, elemtype) <= tol, ("batchnorm", shape)
        # Scale with and without bias
        for bias_term in (1, 0):
            want = ref.layer_forward("Scale", {0: ch, 1: bias_term}, [slope] + ([bias] if bias_term else []), [x], batched=True)[0]
            sd, bd = torch.from_numpy(slope).cuda(), torch.from_numpy(bias).cuda()
            blob = cabi.Blob.from_numpy(x, elemtype, pad_fill=0.0)
            d = blob.desc()
            cabi.check(L.ncnn_cuda_channel_affine(C.byref(d), C.byref(d), C.c_void_p(sd.data_ptr()), C.c_void_p(bd.data_ptr()) if bias_term else None, None), "scale")
            sync()
            assert nerr(blob.numpy(), want, elemtype) <= tol, ("scale", shape, bias_term)
    for (ch, h, w, group, reverse) in [(24, 9, 11, 3, 0), (24, 9, 11, 3, 1), (116, 7, 7, 2, 0), (30, 5, 4, 5, 0), (8, 3, 3, 1, 0)]:
        x = quant(rand(rng, (2, ch, h, w)), elemtype)
        want = ref.layer_forward("ShuffleChannel", {0: group, 1: reverse}, [], [x], batched=True)[0]
        src = cabi.Blob.from_numpy(x, elemtype)
        dst = cabi.Blob((ch, h, w), 2, elemtype, fill=float("nan"))
        sd, dd = src.desc(), dst.desc()
        g = ch // group if reverse else group
        cabi.check(L.ncnn_cuda_shuffle_channel(C.byref(sd), C.byref(dd), g, None), "shuffle_channel")
        sync()
        assert np.array_equal(dst.numpy(), want), ("shufflechannel", ch, group, reverse)


@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_lrn(ref, elemtype):
    """LRN across channels (AlexNet / GoogLeNet) and within channel against the reference's naive layer (src/layer/lrn.cpp)"""
    L = cabi.lib()
    L.ncnn_cuda_lrn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
    rng = np.random.default_rng(61)
    tol = 1e-5 if elemtype == F32 else 1e-4
    for (c, h, w, region, size, alpha, beta, bias) in [(96, 13, 13, 0, 5, 1e-4, 0.75, 1.0), (7, 5, 6, 0, 3, 0.5, 0.6, 2.0), (16, 9, 11, 1, 3, 0.3, 0.75, 1.0),
                                                      (5, 8, 7, 1, 5, 1e-2, 0.5, 1.5), (3, 4, 4, 0, 9, 1.0, 0.75, 1.0)]:
        x = quant(rand(rng, (2, c, h, w), -2.0, 2.0), elemtype)
        want = ref.layer_forward("LRN", {0: region, 1: size, 2: float(alpha), 3: float(beta), 4: float(bias)}, [], [x], batched=True)[0]
        src = cabi.Blob.from_numpy(x, elemtype)
        dst = cabi.Blob((c, h, w), 2, elemtype, fill=float("nan"))
        sd, dd = src.desc(), dst.desc()
        cabi.check(L.ncnn_cuda_lrn(C.byref(sd), C.byref(dd), region, size, alpha, beta, bias, None), "lrn")
        sync()
        assert nerr(dst.numpy(), want, elemtype) <= tol, ("lrn", c, h, w, region, size)


# ------------------------------------------------------------------------------------------ Reduction
def reduce_shape(shape, axes, keepdims):
    """(flags w,h,d,c), top shape in numpy order -- src/layer/reduction.cpp:753-856; `shape` is numpy order (c,[d,]h,w) etc."""
    dims = len(shape)
    names = {1: ["w"], 2: ["h", "w"], 3: ["c", "h", "w"], 4: ["c", "d", "h", "w"]}[dims]
    if axes is None:
        red = set(names)
    elif dims == 1:
        red = {"w"}
    else:
        red = set(names[a if a >= 0 else a + dims] for a in axes)
    if keepdims:
        out = tuple(1 if nm in red else e for nm, e in zip(names, shape))
    else:
        out = tuple(e for nm, e in zip(names, shape) if nm not in red)
        if not out:
            out = (1,)
    return tuple(int(nm in red) for nm in ["w", "h", "d", "c"]), out


@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_reduction(ref, elemtype):
    """all 11 operations x the axis sets of the reference's tests/test_reduction.cpp:116-172 (1-D .. 4-D, keepdims on/off,
    coeff 1 / 2), batched, against its naive layer"""
    L = cabi.lib()
    L.ncnn_cuda_reduction.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(81)
    cases = [((13,), [None, [0]]), ((5, 12), [None, [0], [1], [0, 1], [-1]]), ((7, 5, 9), [None, [0], [1], [2], [0, 2], [1, 2], [0, 1, 2], [-2, -1]]),
             ((6, 3, 5, 4), [None, [0], [3], [0, 3], [1, 3], [2, 3], [1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3], [0, 1, 2, 3]]), ((40, 14, 14), [[1, 2]]), ((1, 1, 300), [[2]])]
    i = 0
    for shape, axis_sets in cases:
        for axes in axis_sets:
            for op in range(11):
                i += 1
                if elemtype != F32 and i % 3:
                    continue
                keepdims, coeff = i % 2, (2.0 if i % 4 < 2 else 1.0)
                lo, hi = (0.001, 2.0) if op in (9, 10) else ((0.7, 1.3) if op == 6 else (-1.0, 1.0))  # positive for the logs; a product stays O(1)
                x = quant(rand(rng, (2,) + shape, lo, hi), elemtype)
                params = {0: op, 1: int(axes is None), 2: float(coeff), 4: keepdims, 5: 1}
                if axes is not None:
                    params[3] = np.asarray(axes, np.int32)
                want = ref.layer_forward("Reduction", params, [], [x], batched=True)[0]
                flags, oshape = reduce_shape(shape, axes, keepdims)
                assert want.shape == (2,) + oshape, (want.shape, oshape, shape, axes, keepdims)
                src = cabi.Blob.from_numpy(x, elemtype)
                dst = cabi.Blob(oshape, 2, elemtype, fill=float("nan"))
                sd, dd = src.desc(), dst.desc()
                cabi.check(L.ncnn_cuda_reduction(op, flags[0], flags[1], flags[2], flags[3], keepdims, coeff, C.byref(sd), C.byref(dd), None), "reduction")
                sync()
                tol = 1e-5 if elemtype == F32 else 1e-4
                e = nerr(dst.numpy(), want, elemtype)
                assert e <= tol, ("reduction", shape, axes, op, keepdims, e)


# ------------------------------------------------------------------------------------------ LayerNorm / GELU
@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_layernorm(ref, elemtype):
    """LayerNorm over w / w*h / w*h*d with and without the element-wise affine, 1-D .. 4-D blobs (the shapes of the reference's
    tests/test_layernorm.cpp plus a transformer-sized row), in place, batched, against its naive layer"""
    import torch
    L = cabi.lib()
    L.ncnn_cuda_layernorm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(91)
    tol = 1e-5 if elemtype == F32 else 1e-4
    cases = [((15,), 15), ((9, 24), 24), ((145, 768), 768), ((5, 6, 8), 8), ((5, 6, 8), 48), ((12, 4, 7), 7), ((3, 2, 5, 6), 6), ((3, 2, 5, 6), 30), ((3, 2, 5, 6), 60)]
    for i, (shape, size) in enumerate(cases):
        for affine in (1, 0):
            eps = [1e-5, 1e-3, 1e-6][i % 3]
            x = quant(rand(rng, (2,) + shape, -2.0, 3.0), elemtype)
            gamma, beta = rand(rng, (size,), 0.5, 1.5), rand(rng, (size,), -0.5, 0.5)
            want = ref.layer_forward("LayerNorm", {0: size, 1: float(eps), 2: affine}, [gamma, beta] if affine else [], [x], batched=True)[0]
            blob = cabi.Blob.from_numpy(x, elemtype)
            d = blob.desc()
            gd, bd = torch.from_numpy(gamma).cuda(), torch.from_numpy(beta).cuda()
            cabi.check(L.ncnn_cuda_layernorm(C.byref(d), C.byref(d), size, eps, C.c_void_p(gd.data_ptr()) if affine else None, C.c_void_p(bd.data_ptr()) if affine else None, None), "layernorm")
            sync()
            e = nerr(blob.numpy(), want, elemtype)
            assert e <= tol, ("layernorm", shape, size, affine, e)


@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_gelu(ref, elemtype):
    """GELU, erfc and tanh forms (src/layer/gelu.cpp), through the unary entry point"""
    L = cabi.lib()
    L.ncnn_cuda_unary.argtypes = [C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(92)
    tol = 1e-5 if elemtype == F32 else 1e-4
    for shape in [(33,), (7, 40), (12, 9, 11)]:
        for fast in (0, 1):
            x = quant(rand(rng, (3,) + shape, -6.0, 6.0), elemtype)
            want = ref.layer_forward("GELU", {0: fast}, [], [x], batched=True)[0]
            blob = cabi.Blob.from_numpy(x, elemtype)
            d = blob.desc()
            cabi.check(L.ncnn_cuda_unary(11, float(fast), 0.0, C.byref(d), C.byref(d), None), "gelu")
            sync()
            assert nerr(blob.numpy(), want, elemtype) <= tol, ("gelu", shape, fast)


# ------------------------------------------------------------------------------------------ YOLOv8 decode
@pytest.mark.parametrize("elemtype", [F32, BF16, F16])
def test_yolov8_decode(elemtype):
    """device decode of the YOLOv8 head (detect.cu) against the restatement of examples/yolov8.cpp generate_proposals
    (oracle/yolov8_decode.py): three strides, 80 / 5 / 33 classes, a mix of rows above and below the threshold, batched"""
    from oracle import yolov8_decode as oy
    L = cabi.lib()
    L.ncnn_cuda_yolov8_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(101)
    strides = [8, 16, 32]
    for (in_w, in_h, num_class, thr) in [(96, 64, 80, 0.25), (64, 64, 5, 0.4), (160, 96, 33, 0.25)]:
        rows = sum((in_w // s) * (in_h // s) for s in strides)
        n = 3
        pred = rand(rng, (n, rows, 64 + num_class), -3.0, 0.0)
        pred[:, :, :64] = rand(rng, (n, rows, 64), -2.0, 4.0)
        pred[:, :, 64:] += rand(rng, (n, rows, 1), -2.0, 1.5)  # per-anchor offset: some rows pass, some do not
        pred[0, 3, 64:] = -1.0                                    # all classes equal: the first one wins
        pred = quant(pred, elemtype)
        want = np.stack([oy.generate_proposals(pred[b], strides, in_w, in_h, thr) for b in range(n)])
        src = cabi.Blob.from_numpy(pred, elemtype)
        dst = cabi.Blob((rows, 6), n, F32, fill=float("nan"))
        sd, dd = src.desc(), dst.desc()
        st = (C.c_int * 3)(*strides)
        cabi.check(L.ncnn_cuda_yolov8_decode(C.byref(sd), st, 3, in_w, in_h, thr, C.byref(dd), None), "yolov8_decode")
        sync()
        got = dst.numpy()
        score = 1.0 / (1.0 + np.exp(-pred[:, :, 64:].max(-1).astype(np.float64)))
        sure = np.abs(score - thr) > 1e-5   # a score within rounding of the threshold may fall on either side
        assert 0.1 < (want[..., 5] >= 0).mean() < 0.9, "the case should mix accepted and rejected anchors"
        assert np.array_equal(got[..., 5][sure], want[..., 5][sure]), "labels / accepted set differ"
        assert np.abs(got[..., 4] - want[..., 4])[sure].max() <= 1e-6
        assert np.abs(got[..., :4] - want[..., :4])[sure].max() <= 1e-5 * max(in_w, in_h)
