"""Multi-GPU mode, host side, on CPU: world_size-2 (and 3, ragged) `gloo` runs of the replica plumbing that
bench.py --gpus N uses (ncnn_b200/replicas.py).  Inference shards by batch only -- replicas, no data-path collective
 -- so what needs covering is the split (Mat::batch_range views, src/mat.h:241-242), the barrier and
the max-over-ranks timing."""
import json
import os
import socket
import subprocess
import sys

import pytest

from ncnn_b200 import replicas

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("total,world", [(256, 1), (256, 2), (256, 8), (7, 3), (2, 4), (0, 2)])
def test_shard_partitions_the_batch(total, world):
    spans = [replicas.shard(total, world, r) for r in range(world)]
    assert spans[0][0] == 0
    for (s0, c0), (s1, _c1) in zip(spans, spans[1:]):
        assert s0 + c0 == s1
    assert spans[-1][0] + spans[-1][1] == total
    counts = [c for _, c in spans]
    assert max(counts) - min(counts) <= 1


def test_shard_rejects_bad_rank():
    with pytest.raises(ValueError):
        replicas.shard(8, 2, 2)


def run_world(world, total):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tests", "replica_worker.py"), str(total)]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout  # rank 0 alone prints
    return json.loads(lines[0])


@pytest.mark.parametrize("world,total", [(2, 16), (3, 7)])
def test_gloo_replicas(world, total):
    out = run_world(world, total)
    assert out["world"] == world
    assert out["images"] == total
    assert out["ms_max"] == 10.0 + 5.0 * (world - 1)  # max over ranks, not rank 0's own time
    shards = sorted(out["shards"], key=lambda s: s["rank"])
    assert [s["rank"] for s in shards] == list(range(world))
    assert all(s["aliases"] for s in shards)
    assert sum(s["count"] for s in shards) == total
    assert abs(sum(s["checksum"] for s in shards) - out["full_checksum"]) < 1e-6  # the shards cover the batch exactly once
    assert abs(out["value"] - total * 4 / (out["ms_max"] / 1000.0)) < 1e-6
