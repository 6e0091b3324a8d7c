"""CUDA-graph replay of a recorded forward walk (ncnn_cuda_graph_begin_capture / _end_capture / _launch of include/ncnn_cuda.h around
the ordinary recorder + Extractor calls; the analogue of re-submitting a recorded VkCompute command buffer, src/command.cpp:1834).

A replay must be the SAME computation as the eager walk it was recorded from: bit-identical output, recomputed from whatever the
input buffer holds at replay time (not a cached result), for the launch-bound batch-1 case it exists for (SqueezeNet v1.1, the
reference's own test network) and for a batched network that goes through every tcgen05 operand mode, the folds and the
in-place Concat plan.
"""
import os

import numpy as np
import pytest

import netutil
from netutil import modelzoo

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _session(text, weights, storage):
    from ncnn_b200 import runner
    return runner.Session(text, weights, storage=storage, device=0)


@pytest.mark.parametrize("storage", ["fp16", "fp32"])
def test_graph_replay_whole_extract_squeezenet(storage):
    """host form: H2D + upload kernel + walk + download kernel + D2H in ONE graph; the input Mat is overwritten in place between
    replays and every replay equals the eager extract on the same bytes"""
    text = open(os.path.join(GOLDEN, "squeezenet_v1.1.param")).read()
    weights = open(os.path.join(GOLDEN, "squeezenet_v1.1.bin"), "rb").read()
    logo = np.load(os.path.join(GOLDEN, "ncnn_logo_16x16.npy"))
    xa = netutil.squeezenet_logo_input(logo)[None].astype(np.float32)
    xb = np.random.default_rng(5).uniform(0, 255, xa.shape).astype(np.float32)
    sess = _session(text, weights, storage)
    try:
        want_a = sess.run_host(xa)
        want_b = sess.run_host(xb)
        assert not np.array_equal(want_a, want_b)
        hin = sess.pinned_input(xa)
        g = sess.capture(host_mat=hin)
        assert g.kernels > 10
        n0 = sess.launch_count()
        for _ in range(3):
            g.replay()
        got_a = g.result()
        assert sess.launch_count() == n0  # a replay is one graph launch: no kernel goes through the launch path
        assert np.array_equal(got_a, want_a)
        assert list(np.argsort(-got_a[0])[:2]) == [532, 920]  # tests/test_squeezenet.cpp:58-92
        sess.L._view(hin, force_batch=True)[...] = xb
        g.replay()
        assert np.array_equal(g.result(), want_b)
        sess.L._view(hin, force_batch=True)[...] = xa
        g.replay()
        assert np.array_equal(g.result(), want_a)
        g.close()
        sess.L.lib.ncnn_mat_destroy(hin)
    finally:
        sess.close()


@pytest.mark.parametrize("model,batch,size", [("resnet50", 2, 224), ("mobilenet_v2", 4, 64), ("yolov8s", 2, 96)])
def test_graph_replay_device_walk(model, batch, size):
    """device form: the walk only, input resident; replay == eager.  An eager walk recorded on the same recorder between two
    replays shares the graph's pool blocks (the stream orders them): its result is read before the next replay, which may
    overwrite it, and the replay after it is undisturbed"""
    text = netutil.with_input_size(modelzoo.param_text(model), size)
    weights = modelzoo.random_model_bytes(text, seed=11)
    x = np.random.default_rng(3).uniform(-1, 1, (batch, 3, size, size)).astype(np.float32)
    sess = _session(text, weights, "fp16")
    lib = sess.L.lib
    try:
        hin = sess.pinned_input(x)
        din = sess.upload(hin)
        eager = sess.enqueue_device(din)
        want = sess.download(eager)
        lib.ncnn_cuda_mat_destroy(eager)
        g = sess.capture(dev_in=din)
        g.replay()
        got1 = sess.download(g.dev_out)
        other = sess.enqueue_device(din)  # an eager walk on the same stream and pool
        got_other = sess.download(other)
        lib.ncnn_cuda_mat_destroy(other)
        g.replay()
        got2 = sess.download(g.dev_out)
        assert np.array_equal(got1, want)
        assert np.array_equal(got2, want)
        assert np.array_equal(got_other, want)
        g.close()
        lib.ncnn_cuda_mat_destroy(din)
        lib.ncnn_mat_destroy(hin)
    finally:
        sess.close()
