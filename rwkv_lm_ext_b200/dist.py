"""Batch sharding helpers for the multi-GPU configs.  Nothing is sharded inside the
recurrence: ranks own contiguous batch slices (what MyBatchSampler does with rank*batch_size,
data/custom_datasets.py:54) and there is no data-path collective; the optional all_gather below only
assembles per-rank [B_local, D] results on every rank."""
import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced split of n items: the first n % world ranks get one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, rank: int = None, world: int = None) -> torch.Tensor:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(x.size(0), rank, world)
    return x[lo:hi]


def gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """all_gather of ragged [B_local, D] shards back into [n_total, D] in batch order."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    maxb = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(maxb, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.size(0)] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def max_over_ranks(ms: float, device="cpu") -> float:
    """Device-time reduction used by bench.py: the job is as slow as its slowest rank."""
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
