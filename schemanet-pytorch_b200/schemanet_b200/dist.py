"""One-process-per-GPU helpers: batch sharding (no collective) and class sharding (one all-gather).

The head is embarrassingly parallel over images (stages 0-2 and the instance side of stage 3), so data-parallel
ranks simply take contiguous slices of the batch -- the reference does the same through DistributedSampler
(schema_inference/data/__init__.py:106-122).  Note the reference semantics this preserves: the pooling mean divides
by the SHARD-local maximum graph size (gnn.py:96 under DDP), so logits of a sharded batch equal those of the
reference run with the same per-GPU batches.  `global_max_vertices` adds the one-int all-reduce(MAX) that reproduces
a single-process big batch instead.
"""
import os

import torch


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of n items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int, dim: int = 0) -> torch.Tensor:
    lo, hi = shard_range(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


def global_max_vertices(max_vertices: torch.Tensor) -> torch.Tensor:
    """all_reduce(MAX) of the per-shard maximum graph size (int32 [1]) -- optional, see module docstring."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(max_vertices, op=dist.ReduceOp.MAX)
    return max_vertices


def gather_class_features(local: torch.Tensor, K: int) -> torch.Tensor:
    """All-gather equally sized [per, D] class-embedding slices into [K, D]."""
    import torch.distributed as dist
    world = dist.get_world_size()
    full = torch.empty(world * local.shape[0], local.shape[1], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, local.contiguous())
    return full[:K]
