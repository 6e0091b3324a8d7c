"""ncnn_b200/replicas.py -- host-side plumbing of the multi-GPU mode: N independent replicas of one Net, batch split.

Inference shards by batch only: every sample is independent through every layer, so there is NO
collective on the data path.  torch.distributed is plumbing here -- rendezvous, the barrier that brackets the timed
region and the max-over-ranks reduction of device times -- never compute.  On a GPU box the backend is "nccl"
(one process per GPU, launched by torchrun); the same code runs under "gloo" on CPU, which is how tests/ cover it.

The reference splits a host batch with zero-copy views, Mat::batch_range (src/mat.h:241-242); shard() gives the
[start, start + count) range of rank r and batch_view() takes that view through the C API.
"""
import os


def env_rank():
    """(rank, local_rank, world) from the torchrun environment; (0, 0, 1) when launched plainly"""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard(total, world, rank):
    """contiguous split of `total` samples over `world` replicas, remainder to the lowest ranks -> (start, count)"""
    if world <= 0 or rank < 0 or rank >= world or total < 0:
        raise ValueError("bad shard request: total %d world %d rank %d" % (total, world, rank))
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def batch_view(L, mat, start, count):
    """zero-copy view of samples [start, start + count) of a batched host Mat (ncnn_mat_batch_range)"""
    import ctypes as C
    L.lib.ncnn_mat_batch_range.restype = C.c_void_p
    L.lib.ncnn_mat_batch_range.argtypes = [C.c_void_p, C.c_int, C.c_int]
    v = L.lib.ncnn_mat_batch_range(mat, start, count)
    if not v:
        raise ValueError("batch range [%d, %d) out of bounds" % (start, start + count))
    return C.c_void_p(v)


class Group(object):
    """the process group of one bench / serving job: barrier + scalar reductions; a no-op for world == 1"""

    def __init__(self, backend=None, device=None):
        self.rank, self.local_rank, self.world = env_rank()
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if backend is None:
                import torch
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                import torch
                torch.cuda.set_device(self.local_rank)
                self.device = "cuda"
            else:
                self.device = "cpu"
            dist.init_process_group(backend=backend)
            self.dist = dist
        self.backend = backend

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, v, op):
        if self.dist is None:
            return float(v)
        import torch
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v):
        """device times are reported as the MAX over ranks (the slowest replica bounds the job)"""
        return self._reduce(v, self.dist.ReduceOp.MAX) if self.dist is not None else float(v)

    def sum(self, v):
        return self._reduce(v, self.dist.ReduceOp.SUM) if self.dist is not None else float(v)

    def gather_objects(self, obj):
        """diagnostics only (tests, per-rank reports); never on the data path"""
        if self.dist is None:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None


def throughput(units_per_rank_step, steps, world, ms_max):
    """whole-job units/s: what all replicas processed / the slowest replica's device time"""
    return units_per_rank_step * world * steps / (ms_max / 1000.0)
