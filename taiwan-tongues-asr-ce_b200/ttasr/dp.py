"""Chunk-level data parallelism over the GPUs of one box: independent 30 s chunks, replicated weights, no collective on
the data path.  One process per GPU (torchrun); the only cross-rank traffic is a host-side
gather of per-shard metadata / results.
"""
from __future__ import annotations


def shard_bounds(n_items: int, rank: int, world: int, keep_together: int = 1) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `n_items` for `rank`.  `keep_together` > 1 keeps groups of that many consecutive
    items (the chunks of one file) on one rank so a file's decoder finds them on one GPU."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    groups = (n_items + keep_together - 1) // keep_together
    base, extra = divmod(groups, world)
    glo = rank * base + min(rank, extra)
    ghi = glo + base + (1 if rank < extra else 0)
    return min(glo * keep_together, n_items), min(ghi * keep_together, n_items)


def gather_host(obj, dst: int = 0):
    """Host-side gather of a picklable per-rank object (shard bounds, checksums, timings) to rank `dst`.
    Works on any torch.distributed backend (NCCL ranks gather through the CPU object path)."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out


def all_max(value: float) -> float:
    """Max over ranks of a host float (timing rule: multi-GPU numbers are the max over ranks)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
