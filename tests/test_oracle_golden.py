"""Pins the oracle (oracle/_ref: the reference's own CPU sources compiled by oracle/build_ref.py) before anything is
checked against it (runs on CPU, no GPU needed):

  * the reference's only literal known answer for this path -- tests/test_squeezenet.cpp:58-92, SqueezeNet v1.1 with its
    real weights on the synthetic logo: top-2 = {532: 0.189459, 920: 0.082801} +-1e-3;
  * the committed fixtures under tests/golden/ (made by tests/golden/make_golden.py from the same library) are
    reproduced, so the GPU box -- where /root/reference does not exist -- checks against the same numbers;
  * batched == per-sample (tests/test_squeezenet.cpp:408-518), the property that lets a per-sample CPU oracle check a
    natively batched GPU kernel;
  * the naive layer (create_layer_naive, the ground truth of tests/testutil.cpp:1301-1339) agrees with the optimised x86
    layer on the hot-path operators within the reference's own epsilon."""
import json
import os

import numpy as np
import pytest

import netutil
from netutil import modelzoo, nerr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ref_net(ref, text, weights, **kw):
    from oracle import ref as oref
    opt = ref.strict_fp32_option(num_threads=ref.cpu_count(), packing=True, **kw)
    return oref.Net(ref, text, weights, opt), opt


def test_reference_known_answer_squeezenet(ref):
    text = open(os.path.join(GOLDEN, "squeezenet_v1.1.param")).read()
    weights = open(os.path.join(GOLDEN, "squeezenet_v1.1.bin"), "rb").read()
    x = netutil.squeezenet_logo_input(np.load(os.path.join(GOLDEN, "ncnn_logo_16x16.npy")))
    expect = json.load(open(os.path.join(GOLDEN, "squeezenet_logo_expect.json")))
    net, opt = ref_net(ref, text, weights)
    prob = net.run({"data": x})["prob"]
    net.close()
    order = np.argsort(-prob)
    assert list(order[:2]) == expect["top2_index"]
    for i, s in zip(expect["top2_index"], expect["top2_score"]):
        assert abs(prob[i] - s) <= expect["epsilon"], (i, prob[i], s)
    want = np.load(os.path.join(GOLDEN, "squeezenet_logo_prob_ref.npy"))
    assert nerr(prob, want) <= 1e-6
    # the winograd/sgemm variants of the same library (the benchncnn configuration) stay within the reference's epsilon
    net, opt = ref_net(ref, text, weights, winograd=True)
    prob_w = net.run({"data": x})["prob"]
    net.close()
    assert nerr(prob_w, want) <= 1e-4


@pytest.mark.parametrize("name", ["squeezenet_v1_1", "mobilenet_v2", "resnet50", "vgg16", "yolov8s"])
def test_fixtures_reproduce(ref, name):
    size = netutil.TEST_SIZES[name]
    text = netutil.with_input_size(modelzoo.param_text(name), size)
    weights = modelzoo.random_model_bytes(text, seed=netutil.WEIGHT_SEED)
    x = netutil.random_input(name, 2, size, seed=1)
    want = np.load(os.path.join(GOLDEN, "%s_ref_n2.npz" % name))
    net, opt = ref_net(ref, text, weights)
    out = net.run({net.input_names[0]: x}, batched=True)
    # batched == per-sample, bit for bit (src/net.cpp:654-705 is a per-sample loop)
    one = net.run({net.input_names[0]: x[1]}, batched=False)
    net.close()
    for k, v in out.items():
        w = want[k.replace("/", "_")]
        assert v.shape == w.shape
        # the fixture was made with the widest ISA of the build box; avx2 vs avx512 kernels differ in summation order
        assert nerr(v, w) <= 2e-5, (k, nerr(v, w))
        assert np.array_equal(v[1], one[k]), k


def test_naive_vs_optimised_layers(ref):
    """the two oracle levels agree on the hot-path operators (1.5e-7 .. 2.8e-6 normalised)"""
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, (16, 20, 20)).astype(np.float32)
    w = (rng.uniform(-1, 1, (24, 16, 3, 3)) * np.sqrt(3.0 / (16 * 9))).astype(np.float32)
    b = rng.uniform(-1, 1, 24).astype(np.float32)
    params = {0: 24, 1: 3, 3: 1, 4: 1, 5: 1, 6: w.size, 9: 1}
    a = ref.layer_forward("Convolution", params, [w, b], [x], naive=True)[0]
    for kw in (dict(winograd=False, sgemm=True), dict(winograd=True, sgemm=True), dict(winograd=False, sgemm=False)):
        opt = ref.strict_fp32_option(num_threads=2, packing=False, **kw)
        c = ref.layer_forward("Convolution", params, [w, b], [x], naive=False, opt=opt)[0]
        ref.lib.ncnn_option_destroy(opt)
        assert nerr(c, a) <= 1e-5, (kw, nerr(c, a))
    wd = (rng.uniform(-1, 1, (16, 1, 3, 3)) / 3).astype(np.float32)
    pd = {0: 16, 1: 3, 3: 2, 4: 1, 5: 0, 6: wd.size, 7: 16}
    a = ref.layer_forward("ConvolutionDepthWise", pd, [wd], [x], naive=True)[0]
    c = ref.layer_forward("ConvolutionDepthWise", pd, [wd], [x], naive=False)[0]
    assert a.shape == (16, 10, 10) and nerr(c, a) <= 1e-6
    pp = {0: 0, 1: 3, 2: 2, 5: 0}
    a = ref.layer_forward("Pooling", pp, [], [x], naive=True)[0]
    c = ref.layer_forward("Pooling", pp, [], [x], naive=False)[0]
    assert a.shape == (16, 10, 10) and np.array_equal(a, c)  # ceil mode: (20-3)/2 -> 9.5 -> 10 windows


def test_yolov8_decode_restatement_matches_reference_example(ref):
    """oracle/yolov8_decode.py (the numpy checker of the device decode) against the reference's OWN post-processing:
    generate_proposals / qsort_descent_inplace / nms_sorted_bboxes of examples/yolov8.cpp:67-273, compiled where they lie into
    oracle/_ref by oracle/build_ref.py (oracle/yolov8_example_driver.cpp).  Pins SURVEY 8f row f4's oracle."""
    import ctypes as C
    from oracle import yolov8_decode as oy
    L = ref.lib
    if not hasattr(L, "ref_yolov8_generate_proposals"):
        pytest.skip("oracle/_ref was built before the YOLOv8 example driver existed: re-run oracle/build_ref.py")
    L.ref_yolov8_generate_proposals.restype = C.c_int
    L.ref_yolov8_generate_proposals.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int]
    L.ref_yolov8_sort_nms.restype = C.c_int
    L.ref_yolov8_sort_nms.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int]
    rng = np.random.default_rng(77)
    strides = [8, 16, 32]
    for (in_w, in_h, num_class, thr) in [(96, 64, 80, 0.25), (64, 64, 5, 0.4), (160, 96, 33, 0.25), (32, 32, 1, 0.5)]:
        rows = sum((in_w // s) * (in_h // s) for s in strides)
        pred = rng.uniform(-3.0, 0.0, (rows, 64 + num_class)).astype(np.float32)
        pred[:, :64] = rng.uniform(-2.0, 4.0, (rows, 64)).astype(np.float32)
        pred[:, 64:] += rng.uniform(-2.0, 1.5, (rows, 1)).astype(np.float32)
        pred[3, 64:] = 1.0  # a tie between all classes: the first one wins in both
        pred = np.ascontiguousarray(pred)
        st = (C.c_int * 3)(*strides)
        out = np.zeros((rows, 6), np.float32)
        n = L.ref_yolov8_generate_proposals(pred.ctypes.data, rows, pred.shape[1], st, 3, in_w, in_h, thr, out.ctypes.data, rows)
        dense = oy.generate_proposals(pred, strides, in_w, in_h, thr)
        mine = dense[dense[:, 5] >= 0]  # the example pushes only the accepted anchors, in anchor order
        score = 1.0 / (1.0 + np.exp(-pred[:, 64:].max(-1).astype(np.float64)))
        assert np.abs(score - thr).min() > 1e-6, "seed puts a score on the threshold: pick another"
        assert n == mine.shape[0] and 0 < n < rows
        got = out[:n]
        assert np.array_equal(got[:, 5], mine[:, 5])
        assert np.abs(got[:, 4] - mine[:, 4]).max() <= 2e-7
        assert np.abs(got[:, :4] - mine[:, :4]).max() <= 2e-5 * max(in_w, in_h)
        # sort + NMS (class-aware and agnostic): same survivors in the same order
        for agnostic in (0, 1):
            boxes = got.copy()
            picked = np.zeros(n, np.int32)
            k = L.ref_yolov8_sort_nms(boxes.ctypes.data, n, 0.45, agnostic, picked.ctypes.data, n)
            order = np.argsort(-got[:, 4], kind="stable")
            srt = got[order]
            if np.unique(got[:, 4]).size == n:  # (the example's quicksort is not stable: compare orders only without ties)
                assert np.array_equal(boxes, srt)
            keep = oy.nms_sorted_bboxes(boxes, 0.45, agnostic=bool(agnostic))
            assert k == len(keep) and list(picked[:k]) == keep
