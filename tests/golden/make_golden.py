#!/usr/bin/env python3
"""Builds the committed golden fixtures from the reference tree (run where /root/reference exists):

  squeezenet_v1.1.param / .bin   the real-weight model the reference's own end-to-end test uses
                                 (tests/test_squeezenet.cpp:150-231, files from examples/)
  ncnn_logo_16x16.npy            the synthetic 16x16 input of that test (test_squeezenet.cpp:14-49), uint8
  squeezenet_logo_expect.json    the literal known answer of that test (:58-92): top-2 = {532: 0.189459, 920: 0.082801}
  squeezenet_logo_prob_ref.npy   the 1000 probabilities the reference CPU path (oracle/_ref) produces here for it
  <model>_ref_n2.npz             reference CPU outputs (strict fp32 options) for the seeded random-weight graphs in models/,
                                 batch 2, input seed 1 -- so the GPU parity tests also have a fixed vector to hit

The GPU parity tests re-run the oracle live as well; these files pin both sides."""
import json
import os
import re
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("NCNN_REFERENCE", "/root/reference")


def main():
    from oracle import ref as oref
    import modelzoo
    import netutil

    for f in ("squeezenet_v1.1.param", "squeezenet_v1.1.bin"):
        shutil.copyfile(os.path.join(REF, "examples", f), os.path.join(HERE, f))
    src = open(os.path.join(REF, "tests/test_squeezenet.cpp")).read()
    body = src[src.index("ncnn_logo_data[16][16]"):]
    body = body[body.index("{"):body.index("};")]
    vals = [int(x) for x in re.findall(r"\d+", body)]
    assert len(vals) == 256
    logo = np.asarray(vals, np.uint8).reshape(16, 16)
    np.save(os.path.join(HERE, "ncnn_logo_16x16.npy"), logo)
    json.dump({"top2_index": [532, 920], "top2_score": [0.189459, 0.082801], "epsilon": 0.001,
               "source": "tests/test_squeezenet.cpp:58-92"}, open(os.path.join(HERE, "squeezenet_logo_expect.json"), "w"), indent=1)

    R = oref.reference()
    x = netutil.squeezenet_logo_input(logo)
    opt = R.strict_fp32_option(num_threads=R.cpu_count(), packing=True)
    net = oref.Net(R, open(os.path.join(HERE, "squeezenet_v1.1.param")).read(), open(os.path.join(HERE, "squeezenet_v1.1.bin"), "rb").read(), opt)
    prob = net.run({"data": x})["prob"]
    net.close()
    order = np.argsort(-prob)
    print("reference top-2:", order[:2], prob[order[:2]])
    assert list(order[:2]) == [532, 920]
    np.save(os.path.join(HERE, "squeezenet_logo_prob_ref.npy"), prob.astype(np.float32))

    for name in ("squeezenet_v1_1", "mobilenet_v2", "resnet50", "vgg16", "yolov8s"):
        text = modelzoo.param_text(name)
        size = netutil.TEST_SIZES[name]
        text = netutil.with_input_size(text, size)
        weights = modelzoo.random_model_bytes(text, seed=netutil.WEIGHT_SEED)
        x = netutil.random_input(name, 2, size, seed=1)
        net = oref.Net(R, text, weights, opt)
        out = net.run({net.input_names[0]: x}, batched=True)
        net.close()
        np.savez_compressed(os.path.join(HERE, "%s_ref_n2.npz" % name), **{k.replace("/", "_"): v for k, v in out.items()})
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
