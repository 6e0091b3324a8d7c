"""Host-side pieces of bench.py that the driver depends on and that need no GPU: the JSON line must be strict JSON whatever a
measurement produced, the workloads are BASELINE.json's configurations, the reference arm's line has the contract's keys."""
import json
import math
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_json_line_is_strict_json():
    line = {"value": float("nan"), "roofline": {"traffic": {"dram_bytes_per_step": float("inf")}, "frac": np.float32(0.5)},
            "list": [1, float("-inf"), np.int64(3), "s"], "ok": 1.25, "none": None}
    out = bench._finite(line)
    text = json.dumps(out, allow_nan=False)  # raises on NaN / inf
    back = json.loads(text)
    assert back["value"] is None and back["roofline"]["traffic"]["dram_bytes_per_step"] is None
    assert back["list"] == [1, None, 3, "s"] and back["ok"] == 1.25 and math.isclose(back["roofline"]["frac"], 0.5)


def test_workloads_are_the_baseline_configurations():
    cfg = " ".join(json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"])
    assert bench.WORKLOADS["resnet50"][1:] == ("resnet50", 256, 224)[1:] and "ResNet-50 224x224 batch 256" in cfg
    assert bench.WORKLOADS["mobilenet_v2"][1] == 128 and "MobileNetV2 fp32 224x224 batch 128" in cfg
    assert bench.WORKLOADS["yolov8s"][1:] == (64, 640) and "YOLOv8s 640x640 batch 64" in cfg
    assert bench.WORKLOADS["squeezenet_v1_1"][1] == 1 and "SqueezeNet v1.1" in cfg
    for name in bench.WORKLOADS:
        text = bench.with_input_size(bench.modelzoo.param_text(bench.WORKLOADS[name][0]), bench.WORKLOADS[name][2])
        inp = [l for l in text.splitlines() if l.startswith("Input")][0]
        assert "0=%d" % bench.WORKLOADS[name][2] in inp and "1=%d" % bench.WORKLOADS[name][2] in inp


def test_traffic_merge_tool(tmp_path):
    a = {"launches": 2, "dram_read_bytes": 10.0, "dram_write_bytes": 5.0, "time_us": 3.0, "kernels": {"k": {"launches": 2, "time_us": 3.0}}}
    b = {"launches": 17, "dram_read_bytes": 7.0, "dram_write_bytes": 1.0, "time_us": 2.0, "kernels": {}}
    pa, pb = tmp_path / "a.json", tmp_path / "b.json"
    pa.write_text(json.dumps(a))
    pb.write_text(json.dumps(b))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "merge_traffic.py"), str(pa), str(pb)], stdout=subprocess.PIPE, check=True, text=True).stdout
    t = json.loads(out)
    assert t["resnet50"]["storage"] == "fp16" and t["resnet50"]["dram_read_bytes"] + t["resnet50"]["dram_write_bytes"] == 15.0
    assert t["mobilenet_v2"]["launches"] == 17
    # the committed capture is what bench.py reads for roofline.traffic: finite numbers only
    c = json.load(open(os.path.join(ROOT, "profiles", "r2", "traffic.json")))
    for wl in ("resnet50", "mobilenet_v2"):
        for k in ("dram_read_bytes", "dram_write_bytes"):
            assert isinstance(c[wl][k], (int, float)) and c[wl][k] == c[wl][k] and c[wl][k] > 0


def test_reference_arm_line_shape(ref):
    """bench.py --impl reference on a tiny sample: the reference CPU path through oracle/_ref, contract keys present"""
    r = bench.cpu_reference_run("squeezenet_v1_1", 227, 1, 1, threads=2)
    assert r["images_per_s"] > 0 and r["threads"] == 2 and r["lib"].startswith("libncnn_ref_")
