#!/usr/bin/env python3
"""Author the .param graphs of the five benchmark configurations (BASELINE.json `configs`) and seeded random .bin
weights for them.

The graphs are written from the published architectures with a small builder (`Graph`), in the conventions the
reference's converters use (Split inserted for multi-consumer blobs, BN folded into the conv bias, activation fused
as param 9 where the benchmark files do so).  tests/test_modelzoo.py checks, where the reference tree is present,
that the four graphs the reference ships (benchmark/models/{squeezenet,mobilenet_v2,resnet50,vgg16}.param) and ours are
the same graph up to blob/layer names.  YOLOv8s is not in the reference tree; it follows the public
Ultralytics yolov8.yaml (scale s: depth 0.33, width 0.50) in pnnx's ncnn conventions (Conv = Convolution + Swish layer).

    python tools/modelzoo.py            # (re)writes models/*.param
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS_DIR = os.path.join(ROOT, "models")


class Graph(object):
    def __init__(self):
        self.layers = []  # [type, name, bottoms, tops, params(dict, ordered)]
        self.counter = 0

    def add(self, type_, name, bottoms, ntop=1, tops=None, **params):
        tops = tops or ([name] if ntop == 1 else ["%s_%d" % (name, i) for i in range(ntop)])
        self.layers.append([type_, name, list(bottoms), list(tops), params])
        return tops[0] if len(tops) == 1 else tops

    # ---- operator helpers (param ids as src/layer/<op>.cpp load_param)
    def input(self, name, w, h, c):
        return self.add("Input", name, [], p0=w, p1=h, p2=c)

    def conv(self, name, x, cin, cout, k, s=1, p=0, act=0, bias=1):
        params = {"p0": cout, "p1": k}
        if s != 1:
            params["p3"] = s
        if p != 0:
            params["p4"] = p
        params["p5"] = bias
        params["p6"] = cout * cin * k * k
        if act:
            params["p9"] = act
        return self.add("Convolution", name, [x], **params)

    def dwconv(self, name, x, c, k, s=1, p=0, act=0):
        params = {"p0": c, "p1": k}
        if s != 1:
            params["p3"] = s
        if p != 0:
            params["p4"] = p
        params["p5"] = 1
        params["p6"] = c * k * k
        params["p7"] = c
        if act:
            params["p9"] = act
        return self.add("ConvolutionDepthWise", name, [x], **params)

    def pool(self, name, x, **kw):
        return self.add("Pooling", name, [x], **kw)

    def fc(self, name, x, cin, cout, act=0):
        params = {"p0": cout, "p1": 1, "p2": cin * cout}
        if act:
            params["p9"] = act
        return self.add("InnerProduct", name, [x], **params)

    # ---- finalisation: Split insertion + text
    def finalize(self):
        consumers = {}
        for li, (t, n, bs, ts, p) in enumerate(self.layers):
            for bi, b in enumerate(bs):
                consumers.setdefault(b, []).append((li, bi))
        out = []
        split_id = 0
        rename = {}  # (layer index, bottom slot) -> blob name
        for li, (t, n, bs, ts, p) in enumerate(self.layers):
            out.append([t, n, [rename.get((li, bi), b) for bi, b in enumerate(bs)], ts, p])
            for top in ts:
                cons = consumers.get(top, [])
                if len(cons) > 1:
                    names = ["%s_splitncnn_%d" % (top, i) for i in range(len(cons))]
                    out.append(["Split", "splitncnn_%d" % split_id, [top], names, {}])
                    split_id += 1
                    # the reference's converters hand the LAST split output to the FIRST consumer
                    for (cl, cb), nm in zip(cons, reversed(names)):
                        rename[(cl, cb)] = nm
        self.layers = out
        return self

    def text(self):
        blobs = set()
        for t, n, bs, ts, p in self.layers:
            blobs.update(bs)
            blobs.update(ts)
        lines = ["7767517", "%d %d" % (len(self.layers), len(blobs))]
        for t, n, bs, ts, p in self.layers:
            items = ["%-24s %-24s %d %d" % (t, n, len(bs), len(ts))] + bs + ts
            for k, v in p.items():
                kid = int(k[1:])
                if isinstance(v, (list, tuple)):
                    items.append("-%d=%d,%s" % (23300 + kid, len(v), ",".join(fmt(x) for x in v)))
                else:
                    items.append("%d=%s" % (kid, fmt(v)))
            lines.append(" ".join(items))
        return "\n".join(lines) + "\n"


def fmt(v):
    if isinstance(v, float):
        return "%e" % v
    return str(int(v))


# ------------------------------------------------------------------------------------------------ models
def squeezenet_v1_1(size=227):
    g = Graph()
    x = g.input("data", size, size, 3)
    x = g.conv("conv1", x, 3, 64, 3, s=2, act=1)
    x = g.pool("pool1", x, p1=3, p2=2)

    def fire(name, x, cin, sq, ex):
        s = g.conv(name + "/squeeze1x1", x, cin, sq, 1, act=1)
        a = g.conv(name + "/expand1x1", s, sq, ex, 1, act=1)
        b = g.conv(name + "/expand3x3", s, sq, ex, 3, p=1, act=1)
        return g.add("Concat", name + "/concat", [a, b])

    x = fire("fire2", x, 64, 16, 64)
    x = fire("fire3", x, 128, 16, 64)
    x = g.pool("pool3", x, p1=3, p2=2)
    x = fire("fire4", x, 128, 32, 128)
    x = fire("fire5", x, 256, 32, 128)
    x = g.pool("pool5", x, p1=3, p2=2)
    x = fire("fire6", x, 256, 48, 192)
    x = fire("fire7", x, 384, 48, 192)
    x = fire("fire8", x, 384, 64, 256)
    x = fire("fire9", x, 512, 64, 256)
    x = g.conv("conv10", x, 512, 1000, 1, p=1, act=1)
    x = g.pool("pool10", x, p0=1, p4=1)
    g.add("Softmax", "prob", [x], tops=["output"])
    return g.finalize()


def mobilenet_v2():
    g = Graph()
    x = g.input("data", 224, 224, 3)
    x = g.conv("conv1", x, 3, 32, 3, s=2, p=1, act=1)
    cin = 32
    # (expansion t, channels, repeats, stride of the first)
    cfg = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
    bi = 0
    for t, c, n, s in cfg:
        for i in range(n):
            bi += 1
            name = "block%d" % bi
            stride = s if i == 0 else 1
            hid = cin * t
            y = g.conv(name + "/expand", x, cin, hid, 1, act=1)
            y = g.dwconv(name + "/dwise", y, hid, 3, s=stride, p=1, act=1)
            y = g.conv(name + "/linear", y, hid, c, 1)
            if stride == 1 and cin == c:
                y = g.add("Eltwise", name + "/add", [x, y], p0=1)
            x = y
            cin = c
    x = g.conv("conv_last", x, 320, 1280, 1, act=1)
    x = g.pool("pool_last", x, p0=1, p4=1)
    x = g.fc("fc", x, 1280, 1000)
    g.add("Softmax", "prob", [x], tops=["output"])
    return g.finalize()


def resnet50():
    g = Graph()
    x = g.input("data", 224, 224, 3)
    x = g.conv("conv1", x, 3, 64, 7, s=2, p=3, act=1)
    x = g.pool("pool1", x, p1=3, p2=2)
    cin = 64
    for stage, (mid, n, s) in enumerate([(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]):
        cout = mid * 4
        for i in range(n):
            name = "res%d%s" % (stage + 2, chr(ord("a") + i))
            stride = s if i == 0 else 1
            # caffe topology: the projection shortcut comes first and the stride sits on the first 1x1
            sc = g.conv(name + "_branch1", x, cin, cout, 1, s=stride) if i == 0 else x
            y = g.conv(name + "_branch2a", x, cin, mid, 1, s=stride, act=1)
            y = g.conv(name + "_branch2b", y, mid, mid, 3, p=1, act=1)
            y = g.conv(name + "_branch2c", y, mid, cout, 1)
            y = g.add("Eltwise", name, [sc, y], p0=1)
            x = g.add("ReLU", name + "_relu", [y])
            cin = cout
    x = g.pool("pool5", x, p0=1, p1=7)
    x = g.fc("fc1000", x, 2048, 1000)
    g.add("Softmax", "prob", [x], tops=["output"])
    return g.finalize()


def vgg16():
    g = Graph()
    x = g.input("data", 224, 224, 3)
    cin = 3
    for bi, (c, n) in enumerate([(64, 2), (128, 2), (256, 3), (512, 3), (512, 3)]):
        for i in range(n):
            x = g.conv("conv%d_%d" % (bi + 1, i + 1), x, cin, c, 3, p=1, act=1)
            cin = c
        x = g.pool("pool%d" % (bi + 1), x, p1=2, p2=2)
    x = g.fc("fc6", x, 512 * 7 * 7, 4096, act=1)
    x = g.fc("fc7", x, 4096, 4096, act=1)
    x = g.fc("fc8", x, 4096, 1000)
    g.add("Softmax", "prob", [x], tops=["output"])
    return g.finalize()


def yolov8s(size=640, nc=80):
    g = Graph()
    uid = [0]

    def cbs(x, cin, cout, k, s=1):
        """ultralytics Conv = Conv2d(no bias) + BN (folded into a bias) + SiLU; pnnx emits Convolution + Swish"""
        uid[0] += 1
        y = g.conv("conv_%d" % uid[0], x, cin, cout, k, s=s, p=k // 2)
        return g.add("Swish", "silu_%d" % uid[0], [y])

    def c2f(x, cin, cout, n, shortcut):
        uid[0] += 1
        tag = "c2f_%d" % uid[0]
        c = cout // 2
        y = cbs(x, cin, cout, 1)
        a, b = g.add("Slice", tag + "_chunk", [y], ntop=2, p0=[-233, -233], p1=0)
        parts = [a, b]
        cur = b
        for i in range(n):
            z = cbs(cur, c, c, 3)
            z = cbs(z, c, c, 3)
            if shortcut:
                z = g.add("BinaryOp", "%s_add_%d" % (tag, i), [cur, z], p0=0)
            parts.append(z)
            cur = z
        y = g.add("Concat", tag + "_cat", parts, p0=0)
        return cbs(y, (2 + n) * c, cout, 1)

    def sppf(x, cin, cout, k=5):
        c = cin // 2
        y = cbs(x, cin, c, 1)
        p1 = g.pool("sppf_pool1", y, p0=0, p1=k, p2=1, p3=k // 2, p5=1)
        p2 = g.pool("sppf_pool2", p1, p0=0, p1=k, p2=1, p3=k // 2, p5=1)
        p3 = g.pool("sppf_pool3", p2, p0=0, p1=k, p2=1, p3=k // 2, p5=1)
        y = g.add("Concat", "sppf_cat", [y, p1, p2, p3], p0=0)
        return cbs(y, c * 4, cout, 1)

    x = g.input("in0", size, size, 3)
    x = cbs(x, 3, 32, 3, 2)
    x = cbs(x, 32, 64, 3, 2)
    x = c2f(x, 64, 64, 1, True)
    x = cbs(x, 64, 128, 3, 2)
    p3 = c2f(x, 128, 128, 2, True)
    x = cbs(p3, 128, 256, 3, 2)
    p4 = c2f(x, 256, 256, 2, True)
    x = cbs(p4, 256, 512, 3, 2)
    x = c2f(x, 512, 512, 1, True)
    p5 = sppf(x, 512, 512)
    # neck
    u = g.add("Interp", "up1", [p5], p0=1, p1=2.0, p2=2.0)
    x = g.add("Concat", "cat_p4", [u, p4], p0=0)
    n4 = c2f(x, 768, 256, 1, False)
    u = g.add("Interp", "up2", [n4], p0=1, p1=2.0, p2=2.0)
    x = g.add("Concat", "cat_p3", [u, p3], p0=0)
    o3 = c2f(x, 384, 128, 1, False)
    x = cbs(o3, 128, 128, 3, 2)
    x = g.add("Concat", "cat_n4", [x, n4], p0=0)
    o4 = c2f(x, 384, 256, 1, False)
    x = cbs(o4, 256, 256, 3, 2)
    x = g.add("Concat", "cat_p5", [x, p5], p0=0)
    o5 = c2f(x, 768, 512, 1, False)
    # detect head: per scale box branch (4 * reg_max = 64 ch) and class branch (nc), concatenated, flattened
    outs = []
    for i, (f, c) in enumerate([(o3, 128), (o4, 256), (o5, 512)]):
        b = cbs(f, c, 64, 3)
        b = cbs(b, 64, 64, 3)
        b = g.conv("head%d_box" % i, b, 64, 64, 1)
        k = cbs(f, c, 128, 3)
        k = cbs(k, 128, 128, 3)
        k = g.conv("head%d_cls" % i, k, 128, nc, 1)
        y = g.add("Concat", "head%d_cat" % i, [b, k], p0=0)
        y = g.add("Reshape", "head%d_flat" % i, [y], p0=-1, p1=64 + nc)   # (w = H*W, h = 144)
        y = g.add("Permute", "head%d_perm" % i, [y], p0=1)                # (w = 144, h = H*W)
        outs.append(y)
    g.add("Concat", "out_cat", outs, tops=["out0"], p0=0)                 # (w = 144, h = 8400)
    return g.finalize()


MODELS = {
    "squeezenet_v1_1": squeezenet_v1_1,
    "mobilenet_v2": mobilenet_v2,
    "resnet50": resnet50,
    "vgg16": vgg16,
    "yolov8s": yolov8s,
}

# (w, h, c) of the input blob and the benchmark batch of each configuration (BASELINE.json)
INPUTS = {
    "squeezenet_v1_1": ((227, 227, 3), 1),
    "mobilenet_v2": ((224, 224, 3), 128),
    "resnet50": ((224, 224, 3), 256),
    "vgg16": ((224, 224, 3), 256),
    "yolov8s": ((640, 640, 3), 64),
}


def param_path(name):
    return os.path.join(MODELS_DIR, name + ".param")


def param_text(name):
    p = param_path(name)
    if os.path.exists(p):
        return open(p).read()
    return MODELS[name]().text()


# ------------------------------------------------------------------------------------------------ weights
def parse_param(text):
    """-> list of (type, name, bottoms, tops, {id: value})"""
    lines = [l for l in text.splitlines() if l.strip()]
    out = []
    for l in lines[2:]:
        tok = l.split()
        t, n, nb, nt = tok[0], tok[1], int(tok[2]), int(tok[3])
        bs = tok[4:4 + nb]
        ts = tok[4 + nb:4 + nb + nt]
        params = {}
        for kv in tok[4 + nb + nt:]:
            k, v = kv.split("=", 1)
            k = int(k)
            if k <= -23300:
                k = -k - 23300
                vals = v.split(",")[1:]
                params[k] = [float(x) if ("." in x or "e" in x.lower()) else int(x) for x in vals]
            elif "," in v:
                params[k] = [float(x) if ("." in x or "e" in x.lower()) else int(x) for x in v.split(",")]
            elif v and (v[0].isalpha() or v[0] == '"'):
                params[k] = v
            else:
                params[k] = float(v) if ("." in v or "e" in v.lower()) else int(v)
        out.append((t, n, bs, ts, params))
    return out


def random_model_bytes(text, seed=7767517, bias_scale=0.1, dtype=np.float32):
    """A .bin byte stream (src/modelbin.cpp layout: 4-byte tag 0 + raw fp32 for weights, raw fp32 for biases) with
    seeded uniform(-a, a) weights, a = sqrt(3 / fan_in) * gain so that activations keep O(1) scale through ReLU/SiLU
    stacks (He-style), and small uniform biases.  The same bytes feed the reference and this runtime."""
    rng = np.random.default_rng(seed)
    chunks = []
    layers = parse_param(text)
    consumer = {}
    for t, n, bs, ts, p in layers:
        for b in bs:
            consumer[b] = t
    for t, n, bs, ts, p in layers:
        if t in ("Convolution", "ConvolutionDepthWise", "InnerProduct", "Deconvolution", "DeconvolutionDepthWise"):
            if t == "InnerProduct":
                num_output, bias_term, wsize = p[0], p.get(1, 0), p[2]
            else:
                num_output, bias_term, wsize = p[0], p.get(5, 0), p[6]
            fan_in = wsize // num_output
            if t.startswith("Deconvolution"):
                # an output pixel only sees the taps congruent to it modulo the stride
                fan_in = max(1, fan_in // (p.get(3, 1) * p.get(13, p.get(3, 1))))
            nxt = consumer.get(ts[0], "")
            if p.get(9, 0) != 0 or nxt in ("ReLU", "Swish"):
                a = np.sqrt(6.0 / fan_in)          # He-uniform: keeps the second moment through ReLU / SiLU
            elif nxt in ("Eltwise", "BinaryOp"):
                a = 0.5 * np.sqrt(3.0 / fan_in)    # end of a residual branch: damped so 16 adds do not blow up
            else:
                a = np.sqrt(3.0 / fan_in)          # linear layer: variance preserving
            w = rng.uniform(-a, a, wsize).astype(np.float32)
            chunks.append(np.zeros(1, np.uint32).tobytes())
            chunks.append(w.tobytes())
            if bias_term:
                chunks.append(rng.uniform(-bias_scale, bias_scale, num_output).astype(np.float32).tobytes())
        elif t == "BatchNorm":
            # src/layer/batchnorm.cpp:22-38: slope, mean, var, bias -- raw fp32, no tag (ModelBin type 1)
            c = p[0]
            chunks.append(rng.uniform(0.5, 1.5, c).astype(np.float32).tobytes())
            chunks.append(rng.uniform(-0.3, 0.3, c).astype(np.float32).tobytes())
            chunks.append(rng.uniform(0.4, 1.6, c).astype(np.float32).tobytes())
            chunks.append(rng.uniform(-0.2, 0.2, c).astype(np.float32).tobytes())
        elif t == "Scale":
            # src/layer/scale.cpp:25-42: scale (+ bias), raw fp32
            c = p[0]
            chunks.append(rng.uniform(0.5, 1.5, c).astype(np.float32).tobytes())
            if p.get(1, 0):
                chunks.append(rng.uniform(-0.2, 0.2, c).astype(np.float32).tobytes())
        elif t == "MultiHeadAttention":
            # src/layer/multiheadattention.cpp:191-238: q, k, v, out weights (tagged) each followed by its bias (raw)
            embed, wsize = p[0], p[2]
            qdim, kdim, vdim = wsize // embed, p.get(3, embed), p.get(4, embed)
            for (rows, cols) in [(embed, qdim), (embed, kdim), (embed, vdim), (qdim, embed)]:
                a = np.sqrt(3.0 / cols)
                chunks.append(np.zeros(1, np.uint32).tobytes())
                chunks.append(rng.uniform(-a, a, rows * cols).astype(np.float32).tobytes())
                chunks.append(rng.uniform(-bias_scale, bias_scale, rows).astype(np.float32).tobytes())
        elif t == "LayerNorm":
            # src/layer/layernorm.cpp:23-36: gamma, beta raw fp32 when affine (id 2, default 1)
            if p.get(2, 1):
                chunks.append(rng.uniform(0.5, 1.5, p[0]).astype(np.float32).tobytes())
                chunks.append(rng.uniform(-0.2, 0.2, p[0]).astype(np.float32).tobytes())
        elif t == "MemoryData":
            # src/layer/memorydata.cpp:26-53: w*h*d*c raw fp32 (load type 1, no tag)
            count = max(p.get(0, 0), 1) * max(p.get(1, 0), 1) * max(p.get(11, 0), 1) * max(p.get(2, 0), 1)
            if p.get(21, 1) != 1:
                raise NotImplementedError("MemoryData load_type %d" % p.get(21, 1))
            chunks.append(rng.uniform(0.5, 1.5, count).astype(np.float32).tobytes())
        elif t == "Gemm":
            # src/layer/gemm.cpp:160-212: constant A / B / C in that order, each tagged (ModelBin type 0)
            M, N, K = p.get(7, 0), p.get(8, 0), p.get(9, 0)
            if p.get(4, 0):
                chunks.append(np.zeros(1, np.uint32).tobytes())
                chunks.append(rng.uniform(-1, 1, M * K).astype(np.float32).tobytes())
            if p.get(5, 0):
                a = np.sqrt(3.0 / K)
                chunks.append(np.zeros(1, np.uint32).tobytes())
                chunks.append(rng.uniform(-a, a, N * K).astype(np.float32).tobytes())
            bt = p.get(10, 0)
            if p.get(6, 0) and bt != -1:
                count = {0: 1, 1: M, 2: M, 3: N * M, 4: N}[bt]
                chunks.append(np.zeros(1, np.uint32).tobytes())
                chunks.append(rng.uniform(-bias_scale, bias_scale, count).astype(np.float32).tobytes())
        elif t in ("PReLU",):
            raise NotImplementedError("random weights for " + t)
    return b"".join(chunks)


def main():
    os.makedirs(MODELS_DIR, exist_ok=True)
    for name, fn in MODELS.items():
        g = fn()
        with open(param_path(name), "w") as f:
            f.write(g.text())
        print("%-18s %3d layers" % (name, len(g.layers)))


if __name__ == "__main__":
    sys.exit(main())
