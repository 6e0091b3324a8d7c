"""Gemm through the Layer API (Net + Extractor, include/c_api.h) against the reference's naive Gemm layer
(src/layer/gemm.cpp:250-315, 579-740), on the parameter tables of the reference's own tests:

  * tests/test_gemm_0.h  (test_gemm_0a..0f.cpp): alpha x transA x transB x output_transpose x constantA x constantB, plus the
    output_N1M forms whose operands are (w, 1, c) blobs;
  * tests/test_gemm_2.h  (test_gemm_2a..2e.cpp): the five broadcast types of C (scalar, M, Mx1, MxN, 1xN / N) x beta x
    constantA / constantB / constantC;
  * tests/test_gemm_1.h  two runtime operands.

fp32 blobs: <= 1e-5 (strided fp32 kernel, or the strict-fp32 implicit-GEMM kernel for constant operands).
fp16 blobs: <= 2e-3 with every operand pre-rounded to fp16 on both sides; the constant-B (any transB) and the constant-A forms
run on the tcgen05 kernel (the constant operand, scaled by alpha, is the re-packed weight matrix; alpha * beta * C the bias).
"""
import numpy as np
import pytest

from netutil import nerr

pytestmark = pytest.mark.gpu

MODES = {
    "fp32": dict(use_fp16_storage=0, use_fp16_packed=0, use_fp16_arithmetic=0, use_bf16_storage=0),
    "fp16": dict(use_fp16_storage=1, use_bf16_storage=0),
}
TOL = {"fp32": 1e-5, "fp16": 2e-3}


def product():
    from ncnn_b200 import capi
    return capi.library()


def q16(a):
    import torch
    return torch.from_numpy(np.asarray(a, np.float32)).to(torch.float16).float().numpy()


def tagged(a):
    """ModelBin::load(..., type 0): a zero fp32 tag, then raw fp32 (src/modelbin.cpp:75-151)"""
    return np.zeros(1, np.uint32).tobytes() + np.ascontiguousarray(a, np.float32).tobytes()


def input_line(name, shape):
    # shape as numpy order: (h, w) 2-D, (c, h, w) 3-D, (w,) 1-D
    if len(shape) == 1:
        return "Input %s 0 1 %s 0=%d" % (name, name, shape[0])
    if len(shape) == 2:
        return "Input %s 0 1 %s 0=%d 1=%d" % (name, name, shape[1], shape[0])
    return "Input %s 0 1 %s 0=%d 1=%d 2=%d" % (name, name, shape[2], shape[1], shape[0])


def run_case(ref, rng, mode, M, N, K, alpha, beta, transA, transB, output_transpose, constantA, constantB, constantC, c_kind, output_N1M=0):
    """c_kind: None (no C: constantC = 1, broadcast type -1), or one of "1", "M", "1M", "NM", "N1", "N" (the RandomMat shapes of test_gemm_2.h)"""
    from ncnn_b200 import capi
    ours = product()
    rnd = (lambda s: q16(rng.uniform(-1, 1, s))) if mode == "fp16" else (lambda s: rng.uniform(-1, 1, s).astype(np.float32))

    def operand(rows, cols, trans):
        # logical rows x cols, stored transposed when trans; N1M operands are (w, 1, c) blobs: numpy (c, 1, w)
        r, c = (cols, rows) if trans else (rows, cols)
        return rnd((r, 1, c)) if output_N1M else rnd((r, c))

    A = operand(M, K, transA)
    B = operand(K, N, transB)  # B is stored [K][N] when transB = 0, [N][K] when transB = 1
    Cm, bt = None, -1
    if c_kind is not None:
        shape = {"1": (1,), "M": (M,), "1M": (M, 1), "NM": (M, N), "N1": (1, N), "N": (N,)}[c_kind]
        Cm = rng.uniform(-1, 1, shape).astype(np.float32)  # C stays fp32 (bias constants are kept fp32)
        if mode == "fp16" and not constantC:
            Cm = q16(Cm)
        bt = {"1": 0, "M": 1, "1M": 2, "NM": 3, "N1": 4, "N": 4}[c_kind]
        if c_kind == "M" and M == N:
            bt = 4  # gemm.cpp:678-689: the N test comes after the M test and wins when M == N
        if c_kind == "1M" and M == 1 and N == 1:
            bt = 4
    params = {0: float(alpha), 1: float(beta), 2: transA, 3: transB, 4: constantA, 5: constantB, 6: 1 if c_kind is None else constantC, 7: M, 8: N, 9: K,
              10: bt if (c_kind is not None) else -1, 11: output_N1M, 14: output_transpose}
    weights, bottoms, names = [], [], []
    if constantA:
        weights.append(A)
    else:
        bottoms.append(A)
        names.append("a")
    if constantB:
        weights.append(B)
    else:
        bottoms.append(B)
        names.append("b")
    if c_kind is not None:
        if constantC:
            weights.append(Cm)
        else:
            bottoms.append(Cm)
            names.append("c")
    if not bottoms:
        return None  # (all-constant Gemm has no input blob: not expressible as a graph)
    want = ref.layer_forward("Gemm", params, weights, bottoms)[0]

    lines = [input_line(nm, b.shape) for nm, b in zip(names, bottoms)]
    ptxt = " ".join("%d=%s" % (k, ("%e" % v) if isinstance(v, float) else str(v)) for k, v in sorted(params.items()))
    lines.append("Gemm gemm %d 1 %s out %s" % (len(bottoms), " ".join(names), ptxt))
    text = "7767517\n%d %d\n%s\n" % (len(lines), len(bottoms) + 1, "\n".join(lines))
    blob = b"".join(tagged(w) for w in weights)
    opt = ours.make_option(1, **MODES[mode])
    net = capi.Net(ours, text, blob, opt)
    try:
        got = net.run(dict(zip(names, bottoms)), outputs=["out"], batched=False)["out"]
    finally:
        net.close()
        ours.lib.ncnn_option_destroy(opt)
    assert got.shape == want.shape, (got.shape, want.shape)
    e = nerr(got, want)
    # a 16-bit top carries its own storage rounding (2^-11 per element) on top of the arithmetic bound
    tol = TOL[mode] + (2.0 ** -11 if mode == "fp16" else 0.0)
    assert e <= tol, "Gemm M%d N%d K%d alpha %.2f beta %.2f tA%d tB%d oT%d cA%d cB%d cC%d C=%s N1M%d %s: err %.3g" % (
        M, N, K, alpha, beta, transA, transB, output_transpose, constantA, constantB, constantC, c_kind, output_N1M, mode, e)
    return e


# tests/test_gemm_0.h: (alpha, transA, transB, output_transpose, constantA, constantB, output_N1M)
GRID0 = [(2.1, 0, 0, 0, 0, 0, 0), (3.1, 0, 1, 0, 0, 0, 0), (4.1, 1, 0, 0, 0, 0, 0), (2.4, 1, 1, 0, 0, 0, 0), (2.1, 0, 0, 1, 0, 0, 0), (3.1, 0, 1, 1, 0, 0, 0),
         (4.1, 1, 0, 1, 0, 0, 0), (2.4, 1, 1, 1, 0, 0, 0),
         (1.7, 0, 1, 0, 0, 0, 1), (1.7, 1, 1, 0, 1, 0, 1), (1.9, 0, 0, 0, 0, 1, 1), (1.7, 0, 1, 1, 1, 0, 1), (1.7, 1, 1, 1, 0, 1, 1), (1.9, 1, 0, 1, 0, 0, 1),
         (2.1, 0, 0, 0, 1, 0, 0), (3.1, 0, 1, 0, 1, 0, 0), (4.1, 1, 0, 0, 1, 0, 0), (2.4, 1, 1, 0, 1, 0, 0), (2.1, 0, 0, 1, 1, 0, 0), (3.1, 0, 1, 1, 1, 0, 0),
         (4.1, 1, 0, 1, 1, 0, 0), (2.4, 1, 1, 1, 1, 0, 0),
         (2.1, 0, 0, 0, 0, 1, 0), (3.1, 0, 1, 0, 0, 1, 0), (4.1, 1, 0, 0, 0, 1, 0), (2.4, 1, 1, 0, 0, 1, 0), (2.1, 0, 0, 1, 0, 1, 0), (3.1, 0, 1, 1, 0, 1, 0),
         (4.1, 1, 0, 1, 0, 1, 0), (2.4, 1, 1, 1, 0, 1, 0)]
SIZES0 = [(1, 1, 1), (2, 2, 2), (3, 3, 3), (5, 5, 5), (8, 8, 8), (15, 16, 16), (16, 20, 15), (31, 47, 23), (128, 200, 72), (300, 136, 264)]


@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_gemm_trans_constant_grid(ref, mode):
    rng = np.random.default_rng(20)
    worst = 0.0
    for (M, N, K) in SIZES0:
        for alpha, tA, tB, oT, cA, cB, n1m in GRID0:
            if n1m and max(M, N, K) > 64:
                continue
            e = run_case(ref, rng, mode, M, N, K, alpha, 1.0, tA, tB, oT, cA, cB, 1, None, output_N1M=n1m)
            worst = max(worst, e or 0.0)
    print("\n[gemm grid0] %s worst err %.3g" % (mode, worst))


# tests/test_gemm_2.h: (C kind, alpha, beta, transA, transB, output_transpose, constantA, constantB, constantC)
GRID2 = []
for cA, cB, cC, perm in [(0, 0, 0, 0), (1, 0, 0, 1), (0, 1, 0, 2), (1, 1, 0, 3), (0, 0, 1, 1), (1, 0, 1, 2), (0, 1, 1, 3), (1, 1, 1, 0)]:
    tt = [(0, 0, 0), (0, 1, 0), (1, 0, 1), (1, 1, 1), (0, 0, 0), (0, 1, 0)]
    for i, (kind, alpha, beta) in enumerate([("1", 2.1, 0.5), ("M", 3.1, 0.6), ("1M", 4.1, 0.7), ("NM", 5.1, 0.8), ("N1", 2.1, 0.5), ("N", 3.1, 0.6)]):
        tA, tB, oT = tt[(i + perm) % 6]
        GRID2.append((kind, alpha, beta, tA, tB, oT, cA, cB, cC))
SIZES2 = [(1, 1, 1), (3, 3, 3), (4, 5, 6), (8, 8, 8), (13, 20, 17), (40, 72, 64), (136, 96, 200)]


@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_gemm_bias_broadcast_grid(ref, mode):
    rng = np.random.default_rng(21)
    worst = 0.0
    for (M, N, K) in SIZES2:
        for kind, alpha, beta, tA, tB, oT, cA, cB, cC in GRID2:
            e = run_case(ref, rng, mode, M, N, K, alpha, beta, tA, tB, oT, cA, cB, cC, kind)
            worst = max(worst, e or 0.0)
    print("\n[gemm grid2] %s worst err %.3g" % (mode, worst))


def test_gemm_constant_operand_forms_use_tcgen05():
    """the constant-B (any transB) and constant-A (transB = 1) forms must be served by the tcgen05 implicit-GEMM kernel
    (ncnn_cuda_tc_launch_count counts its launches), two runtime operands by the strided kernel"""
    import ctypes as C
    from ncnn_b200 import capi
    ours = product()
    ours.lib.ncnn_cuda_tc_launch_count.restype = C.c_ulonglong
    rng = np.random.default_rng(3)
    M, N, K = 256, 192, 128
    for cA, cB, tA, tB, oT, on_tc in [(0, 1, 0, 1, 0, 1), (0, 1, 0, 0, 0, 1), (0, 1, 0, 0, 1, 1), (1, 0, 0, 1, 1, 1), (1, 0, 1, 1, 0, 1), (0, 1, 1, 0, 0, 0)]:
        A = q16(rng.uniform(-1, 1, (K, M) if tA else (M, K)))
        B = q16(rng.uniform(-1, 1, (N, K) if tB else (K, N)))
        params = {0: 1.5, 1: 1.0, 2: tA, 3: tB, 4: cA, 5: cB, 6: 1, 7: M, 8: N, 9: K, 10: -1, 14: oT}
        x, w, nm = (B, A, "b") if cA else (A, B, "a")
        ptxt = " ".join("%d=%s" % (k, ("%e" % v) if isinstance(v, float) else str(v)) for k, v in sorted(params.items()))
        text = "7767517\n2 2\n%s\nGemm gemm 1 1 %s out %s\n" % (input_line(nm, x.shape), nm, ptxt)
        opt = ours.make_option(1, **MODES["fp16"])
        net = capi.Net(ours, text, tagged(w), opt)
        try:
            n0 = int(ours.lib.ncnn_cuda_tc_launch_count())
            got = net.run({nm: x}, outputs=["out"])["out"]
            tc_launches = int(ours.lib.ncnn_cuda_tc_launch_count()) - n0
        finally:
            net.close()
            ours.lib.ncnn_option_destroy(opt)
        opA = A.T if tA else A
        opB = B.T if tB else B
        want = 1.5 * (opA.astype(np.float64) @ opB.astype(np.float64))
        if oT:
            want = want.T
        assert nerr(got, want) <= 2e-3 + 2.0 ** -11
        assert tc_launches == on_tc, (cA, cB, tA, tB, oT, tc_launches)
