"""The authored graphs (tools/modelzoo.py -> models/*.param) against the ones the reference ships
(benchmark/models/*.param): same layer sequence, same parameters, same wiring, up to names and shape hints."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import modelzoo  # noqa: E402

REF_MODELS = "/root/reference/benchmark/models"


def canonical(text):
    layers = modelzoo.parse_param(text)
    ids = {}
    out = []
    for t, n, bs, ts, p in layers:
        for b in bs + ts:
            ids.setdefault(b, len(ids))
        p = {k: v for k, v in p.items() if k != 30}
        out.append((t, [ids[b] for b in bs], [ids[b] for b in ts], sorted(p.items())))
    return out


@pytest.mark.parametrize("ours,theirs", [("squeezenet_v1_1", "squeezenet"), ("mobilenet_v2", "mobilenet_v2"), ("resnet50", "resnet50"), ("vgg16", "vgg16")])
def test_graph_matches_reference(ours, theirs):
    path = os.path.join(REF_MODELS, theirs + ".param")
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    a = canonical(modelzoo.MODELS[ours]().text())
    b = canonical(open(path).read())
    assert len(a) == len(b)
    for i, (x, y) in enumerate(zip(a, b)):
        assert x == y, "layer %d differs:\n ours   %s\n theirs %s" % (i, x, y)


def test_committed_params_are_current():
    for name, fn in modelzoo.MODELS.items():
        assert open(modelzoo.param_path(name)).read() == fn().text(), "models/%s.param is stale: run tools/modelzoo.py" % name


def test_yolov8s_shape():
    layers = modelzoo.parse_param(modelzoo.param_text("yolov8s"))
    types = [l[0] for l in layers]
    assert types.count("Convolution") == 63 and types.count("Swish") == 57
    assert layers[-1][3] == ["out0"]
