"""Multi-GPU plumbing: streams are independent plug-in instances
(PluginProcessor.h:69-73 -- all state is per instance), so the batch shards by
contiguous stream ranges, one process per GPU, with NO data-path collective
(SURVEY.md 8(e)). torch.distributed is only used for the launch rendezvous,
the barrier around the timed region and the max/sum of per-rank scalars."""
import os


def stream_range(rank, world, n_streams):
    """GPU `rank` of `world` owns streams [lo, hi) of the global batch: lo = rank*S/G, hi = (rank+1)*S/G."""
    if world <= 0 or not (0 <= rank < world) or n_streams < 0:
        raise ValueError("bad shard arguments rank=%r world=%r n_streams=%r" % (rank, world, n_streams))
    return (rank * n_streams) // world, ((rank + 1) * n_streams) // world


def dist_env():
    """(rank, local_rank, world_size) from the torchrun environment (1 process when absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


class Group:
    """Thin wrapper over torch.distributed for scalar reductions; a no-op for world_size 1."""

    def __init__(self, backend=None, device=None):
        self.rank, self.local_rank, self.world = dist_env()
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            if backend is None:
                # scalars only (barrier, max / sum of per-rank timings): gloo on the host is all it takes, and it keeps
                # NCCL (and its banner on stdout) out of a path that has no collective
                backend = "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                self.device = torch.device("cuda", self.local_rank)
            else:
                self.device = torch.device("cpu")
            if not dist.is_initialized():
                dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world)
            self.dist = dist
            self.torch = torch

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, value, op):
        if self.dist is None:
            return float(value)
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, value):
        return self._reduce(value, self.dist.ReduceOp.MAX if self.dist else None)

    def min(self, value):
        return self._reduce(value, self.dist.ReduceOp.MIN if self.dist else None)

    def sum(self, value):
        return self._reduce(value, self.dist.ReduceOp.SUM if self.dist else None)

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()
            self.dist = None


def aggregate_throughput(group, units_this_rank, seconds_this_rank):
    """Whole-job throughput = units processed by all ranks / the slowest rank's time."""
    total = group.sum(units_this_rank)
    t = group.max(seconds_this_rank)
    return total / t if t > 0 else 0.0, t, total
