"""Shared helpers of the network-level parity tests, smoke() and bench.py: seeded inputs, model variants."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import modelzoo  # noqa: E402

WEIGHT_SEED = 7767517

# spatial size used by the parity tests (the graphs are fully convolutional up to the classifier; VGG16's fc6 pins 224)
TEST_SIZES = {"squeezenet_v1_1": 227, "mobilenet_v2": 224, "resnet50": 224, "vgg16": 224, "yolov8s": 320}


def with_input_size(text, size):
    """rewrite the Input layer's w/h (params 0/1)"""
    lines = text.splitlines()
    for i, l in enumerate(lines):
        if l.startswith("Input"):
            tok = l.split()
            tok = [("0=%d" % size) if t.startswith("0=") else (("1=%d" % size) if t.startswith("1=") else t) for t in tok]
            lines[i] = " ".join(tok)
    return "\n".join(lines) + "\n"


def random_input(name, n, size, seed=1):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, (n, 3, size, size)).astype(np.float32)


def squeezenet_logo_input(logo16):
    """tests/test_squeezenet.cpp:14-49 + :203-206: gray -> BGR (three equal planes), nearest resize to 227x227
    (src/layer/interp.cpp:606-625: in_x = min((int)(x * (w / (float)outw)), w - 1), float arithmetic), minus the means"""
    w = h = 227
    scale = np.float32(16) / np.float32(w)
    idx = np.minimum((np.arange(w, dtype=np.float32) * scale).astype(np.int32), 15)
    img = logo16[idx][:, idx].astype(np.float32)
    means = np.asarray([104.0, 117.0, 123.0], np.float32)
    return np.stack([img - m for m in means]).astype(np.float32)


def nerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
