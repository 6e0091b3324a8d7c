"""Per-operator grids for the glue layers, through the Layer API (a one-layer Net: load_param + forward on the device) against
the reference's naive layers (oracle/_ref, create_layer_naive), on the shape / parameter tables of the reference's own tests:

  BinaryOp  tests/test_binaryop.cpp:108-310 (same-rank broadcasting grids for ranks 1-4, scalar form, all 12 op types) and the
            rank-mismatch forms of docs/developer-guide/binaryop-broadcasting.md that the models use (3-D x per-channel 1-D ...)
  Eltwise   tests/test_eltwise.cpp (PROD / SUM / MAX, coefficients, 2-4 bottoms)
  Concat    tests/test_concat.cpp (ranks 1-4, every axis, negative axes, 2-3 bottoms of different extents)
  Slice     tests/test_slice.cpp (explicit sizes, -233 remainders, indices, every axis)
  Interp    tests/test_interp.cpp (nearest / bilinear, scale factors, explicit output sizes, align_corner, dynamic target size)
  Softmax   tests/test_softmax.cpp (ranks 1-4, every axis)

fp32 blobs <= 1e-5 (transcendental ops <= 2e-5: expf / powf of the device library against glibc); fp16 blobs: inputs pre-rounded,
<= 2e-3 on the arithmetic plus the top blob's own storage rounding.  Every case also runs batched (n = 2) and must equal the
per-sample results.
"""
import numpy as np
import pytest

from netutil import nerr

pytestmark = pytest.mark.gpu

MODES = {
    "fp32": dict(use_fp16_storage=0, use_fp16_packed=0, use_fp16_arithmetic=0, use_bf16_storage=0),
    "fp16": dict(use_fp16_storage=1, use_bf16_storage=0),
}


def product():
    from ncnn_b200 import capi
    return capi.library()


def q16(a):
    import torch
    return torch.from_numpy(np.asarray(a, np.float32)).to(torch.float16).float().numpy()


def input_line(name, shape):
    # numpy order: (w,), (h, w), (c, h, w), (c, d, h, w)
    ids = [(0, shape[-1])]
    if len(shape) >= 2:
        ids.append((1, shape[-2]))
    if len(shape) == 3:
        ids.append((2, shape[0]))
    if len(shape) == 4:
        ids.append((11, shape[1]))
        ids.append((2, shape[0]))
    return "Input %s 0 1 %s %s" % (name, name, " ".join("%d=%d" % kv for kv in ids))


def param_text(params):
    out = []
    for k, v in sorted(params.items()):
        if isinstance(k, str):
            continue
        if isinstance(v, (list, tuple, np.ndarray)):
            v = list(v)
            is_f = any(isinstance(x, (float, np.floating)) for x in v)
            out.append("-%d=%d,%s" % (23300 + k, len(v), ",".join(("%e" % x) if is_f else str(int(x)) for x in v)))
        elif isinstance(v, (float, np.floating)):
            out.append("%d=%e" % (k, v))
        else:
            out.append("%d=%d" % (k, v))
    return " ".join(out)


def run_layer(ref, mode, type_name, params, bottoms, ntop=1, tol=None, ref_params=None):
    """one layer on the device through a Net, unbatched and batched, against the reference naive layer"""
    from ncnn_b200 import capi
    ours = product()
    if mode == "fp16":
        bottoms = [q16(b) for b in bottoms]
    rp = dict(ref_params if ref_params is not None else params)
    rp = dict((k, (np.asarray(v, np.float32) if any(isinstance(x, (float, np.floating)) for x in v) else np.asarray(v, np.int32)) if isinstance(v, (list, tuple)) else v)
              for k, v in rp.items())
    if ntop > 1:
        rp["_ntop"] = ntop
    want = ref.layer_forward(type_name, rp, [], bottoms)
    names = ["in%d" % i for i in range(len(bottoms))]
    tops = ["out%d" % i for i in range(ntop)]
    lines = [input_line(nm, b.shape) for nm, b in zip(names, bottoms)]
    lines.append("%s op %d %d %s %s %s" % (type_name, len(bottoms), ntop, " ".join(names), " ".join(tops), param_text(params)))
    text = "7767517\n%d %d\n%s\n" % (len(lines), len(bottoms) + ntop, "\n".join(lines))
    opt = ours.make_option(1, **MODES[mode])
    net = capi.Net(ours, text, b"", opt)
    try:
        got = net.run(dict(zip(names, bottoms)), outputs=tops, batched=False)
        # batched: sample 0 = the case, sample 1 = a scaled copy; row 0 must reproduce the unbatched result bit for bit
        stacked = [np.stack([b, b * np.float32(0.5)]) for b in bottoms]
        gotb = net.run(dict(zip(names, stacked)), outputs=tops, batched=True)
    finally:
        net.close()
        ours.lib.ncnn_option_destroy(opt)
    base = tol if tol is not None else (1e-5 if mode == "fp32" else 2e-3)
    worst = 0.0
    for i, t in enumerate(tops):
        assert got[t].shape == want[i].shape, (type_name, params, got[t].shape, want[i].shape)
        assert np.isfinite(got[t]).all()
        d = np.abs(got[t].astype(np.float64) - want[i]) - (2.0 ** -11 if mode == "fp16" else 0.0) * np.abs(want[i])
        e = max(d.max(), 0.0) / max(np.abs(want[i]).max(), 1e-30) if want[i].size else 0.0
        assert e <= base, "%s %s %s shapes %s: err %.3g" % (type_name, params, mode, [b.shape for b in bottoms], e)
        assert np.array_equal(gotb[t][0], got[t]), "%s %s: batched row differs from the unbatched result" % (type_name, params)
        worst = max(worst, e)
    return worst


def rnd(rng, shape, lo=-1.0, hi=1.0):
    return rng.uniform(lo, hi, shape).astype(np.float32)


# ------------------------------------------------------------------------------------------ BinaryOp
def _same_rank_grid(full):
    """every combination of (extent or 1) per axis for a and b: tests/test_binaryop.cpp test_binaryop_1..4"""
    import itertools
    variants = [tuple(full[i] if keep[i] else 1 for i in range(len(full))) for keep in itertools.product([1, 0], repeat=len(full))]
    return [(a, b) for a in variants for b in variants]


@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_binaryop_grid(ref, mode):
    rng = np.random.default_rng(31)
    worst = 0.0
    shapes = []
    for full in [(31,), (24,), (31, 13), (24, 15), (31, 3, 7), (24, 5, 5), (28, 4, 6), (5, 3, 7, 2)]:
        shapes += _same_rank_grid(full)
    # rank-mismatch forms (binaryop.cpp:345-406): lower rank aligned to the outer axes
    shapes += [((31, 3, 7), (31,)), ((31,), (31, 3, 7)), ((24, 5, 5), (1,)), ((24, 5, 5), (24, 5)), ((31, 13), (31,)), ((31,), (31, 13)),
               ((5, 3, 7, 2), (5,)), ((5, 3, 7, 2), (5, 3)), ((5, 3, 7, 2), (5, 3, 7)), ((28, 4, 6), (6,))]
    ops = list(range(12))
    k = 0
    for (sa, sb) in shapes:
        # every op on the small grids would be thousands of nets: rotate the op over the shape pairs, all 12 ops on a fixed subset
        todo = ops if k % 29 == 0 else [ops[k % 12]]
        k += 1
        for op in todo:
            # positive operands: pow / rpow / div are then well-conditioned (the reference's tests do the same for pow)
            a, b = rnd(rng, sa, 0.3, 2.0), rnd(rng, sb, 0.3, 2.0)
            tol = None
            if op in (6, 9, 10, 11):
                tol = 3e-5 if mode == "fp32" else 4e-3  # powf / atan2f: device library vs glibc
            worst = max(worst, run_layer(ref, mode, "BinaryOp", {0: op}, [a, b], tol=tol))
    # scalar form (with_scalar = 1), in place
    for op in ops:
        for shape in [(31,), (13, 31), (24, 5, 5), (5, 3, 7, 2)]:
            tol = (3e-5 if mode == "fp32" else 4e-3) if op in (6, 9, 10, 11) else None
            worst = max(worst, run_layer(ref, mode, "BinaryOp", {0: op, 1: 1, 2: 0.2}, [rnd(rng, shape, 0.3, 2.0)], tol=tol))
    print("\n[binaryop] %s worst err %.3g over %d shape pairs" % (mode, worst, len(shapes)))


# ------------------------------------------------------------------------------------------ Eltwise
@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_eltwise_grid(ref, mode):
    rng = np.random.default_rng(32)
    worst = 0.0
    for shape in [(16,), (7, 12), (12, 5, 7), (16, 6, 6), (3, 2, 5, 4)]:
        for nb in (2, 3, 4):
            bottoms = [rnd(rng, shape) for _ in range(nb)]
            for op in (0, 1, 2):
                worst = max(worst, run_layer(ref, mode, "Eltwise", {0: op}, bottoms))
            coeffs = [float(x) for x in rng.uniform(-2, 2, nb)]
            worst = max(worst, run_layer(ref, mode, "Eltwise", {0: 1, 1: coeffs}, bottoms))
    print("\n[eltwise] %s worst err %.3g" % (mode, worst))


# ------------------------------------------------------------------------------------------ Concat / Slice
@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_concat_grid(ref, mode):
    rng = np.random.default_rng(33)
    n = 0
    for base in [(13,), (5, 9), (8, 5, 7), (16, 4, 6), (3, 4, 5, 6)]:
        dims = len(base)
        for axis in list(range(dims)) + [-1]:
            pa = axis % dims
            for extents in ([3, 5], [8, 16, 4], [1, 2]):
                bottoms = []
                for e in extents:
                    s = list(base)
                    s[pa] = e
                    bottoms.append(rnd(rng, tuple(s)))
                e = run_layer(ref, mode, "Concat", {0: axis}, bottoms, tol=0.0 if mode == "fp32" else None)
                assert e == 0.0 or mode != "fp32"
                n += 1
    print("\n[concat] %s %d cases (copies are exact)" % (mode, n))


@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_slice_grid(ref, mode):
    rng = np.random.default_rng(34)
    n = 0
    for shape in [(24,), (12, 18), (16, 6, 12), (24, 4, 6), (6, 4, 8, 12)]:
        dims = len(shape)
        for axis in list(range(dims)) + [-1]:
            ext = shape[axis % dims]
            x = rnd(rng, shape)
            cases = [({0: [ext // 2, -233], 1: axis}, 2), ({0: [-233, -233, -233], 1: axis}, 3), ({0: [2, ext // 3, -233], 1: axis}, 3),
                     ({0: [-233, -233], 1: axis, 2: [ext // 3]}, 2), ({0: [-233, -233, -233], 1: axis, 2: [2, -2]}, 3)]
            if ext < 6:
                cases = cases[:4]  # (the 2 / -2 index pair needs a non-empty middle slice)
            for params, ntop in cases:
                run_layer(ref, mode, "Slice", params, [x], ntop=ntop, tol=0.0 if mode == "fp32" else None)
                n += 1
    print("\n[slice] %s %d cases (copies are exact)" % (mode, n))


# ------------------------------------------------------------------------------------------ Interp
@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_interp_grid(ref, mode):
    rng = np.random.default_rng(35)
    worst = 0.0
    for shape in [(4, 13, 15), (16, 7, 9), (3, 20, 20), (8, 1, 6)]:
        x = rnd(rng, shape)
        c, h, w = shape
        for rt in (1, 2):
            # scale factors (interp.cpp:439-443), explicit output sizes, up and down, identity
            for hs, ws in [(2.0, 2.0), (4.0, 0.5), (0.8, 1.2), (1.0, 1.0), (1.5, 3.0)]:
                if int(h * hs) < 1 or int(w * ws) < 1:
                    continue
                worst = max(worst, run_layer(ref, mode, "Interp", {0: rt, 1: float(hs), 2: float(ws)}, [x]))
            for oh, ow in [(2, 2), (7, 5), (h, w), (2 * h + 1, 3 * w - 1), (1, 1)]:
                worst = max(worst, run_layer(ref, mode, "Interp", {0: rt, 3: oh, 4: ow}, [x]))
        for oh, ow in [(2 * h, 2 * w), (5, 9), (h, w)]:
            if oh > 1 and ow > 1:
                worst = max(worst, run_layer(ref, mode, "Interp", {0: 2, 3: oh, 4: ow, 6: 1}, [x]))  # bilinear, align_corner
        # dynamic target size: the second bottom lends its w / h (interp.cpp:455-470)
        like = rnd(rng, (2, h + 3, 2 * w))
        for rt in (1, 2):
            worst = max(worst, run_layer(ref, mode, "Interp", {0: rt, 5: 1}, [x, like]))
    print("\n[interp] %s worst err %.3g" % (mode, worst))


# ------------------------------------------------------------------------------------------ Softmax
@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_softmax_grid(ref, mode):
    rng = np.random.default_rng(36)
    worst = 0.0
    for shape in [(1000,), (37,), (12, 24), (24, 5, 7), (5, 24, 6), (3, 4, 6, 8), (1, 1, 91)]:
        dims = len(shape)
        x = rnd(rng, shape, -4.0, 4.0)
        for axis in list(range(dims)) + [-1, -dims]:
            # outputs are probabilities: normalised by max|ref| <= 1 the bound is absolute; fp16 storage of p adds its own rounding
            worst = max(worst, run_layer(ref, mode, "Softmax", {0: axis, 1: 1}, [x], tol=2e-6 if mode == "fp32" else 2e-3))
    print("\n[softmax] %s worst err %.3g" % (mode, worst))
