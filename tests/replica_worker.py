"""World-size-N worker for tests/test_replicas_gloo.py (launched by torch.distributed.run, gloo backend, CPU only).

Exercises exactly the host-side logic bench.py uses for N > 1 -- rendezvous, batch sharding through zero-copy
Mat::batch_range views of ONE seeded global batch, the barrier, the max-over-ranks reduction and the whole-job
throughput formula -- with no compute call (there is no GPU here).  Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ncnn_b200 import capi, replicas  # noqa: E402


def main():
    total = int(sys.argv[1])
    g = replicas.Group(backend="gloo")
    L = capi.library()
    rng = np.random.default_rng(7767517)
    x = rng.uniform(-1, 1, (total, 3, 5, 7)).astype(np.float32)  # the same global batch on every rank (seeded)
    m = L.mat_from_numpy(x, batched=True)
    start, count = replicas.shard(total, g.world, g.rank)
    info = {"rank": g.rank, "start": start, "count": count}
    if count > 0:
        v = replicas.batch_view(L, m, start, count)
        mine = L._view(v, force_batch=True)
        assert mine.shape == (count, 3, 5, 7), mine.shape
        assert np.array_equal(mine, x[start:start + count])
        # zero copy: the view aliases the parent's memory
        info["aliases"] = bool(L.lib.ncnn_mat_get_data(v) == L.lib.ncnn_mat_get_data(m) + start * L.lib.ncnn_mat_get_nstep(m) * 4)
        info["checksum"] = float(np.asarray(mine, np.float64).sum())
        L.lib.ncnn_mat_destroy(v)
    else:
        info["aliases"] = True
        info["checksum"] = 0.0
    g.barrier()
    fake_ms = 10.0 + 5.0 * g.rank  # the slowest replica bounds the job
    ms = g.max(fake_ms)
    images = g.sum(count)
    infos = g.gather_objects(info)
    g.barrier()
    if g.rank == 0:
        print(json.dumps({"world": g.world, "ms_max": ms, "images": images, "shards": infos, "full_checksum": float(np.asarray(x, np.float64).sum()),
                          "value": replicas.throughput(total / g.world, 4, g.world, ms)}))
    L.lib.ncnn_mat_destroy(m)
    g.close()


if __name__ == "__main__":
    main()
