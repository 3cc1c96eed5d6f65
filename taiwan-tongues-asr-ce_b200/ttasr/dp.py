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


def probe_digest(pipeline, n_chunks: int = 3, seed: int = 4242):
    """Cross-rank determinism check (SURVEY.md 8e "same chunk => bit-identical output on any rank"): every rank encodes
    the same seeded probe batch (noise, a short clip behind n_valid, an all-zero chunk) through `pipeline` and the
    SHA-256 digests of the bf16 hidden states are compared across the control group.
    Returns (hex digest of this rank, True iff every rank reported the same digest)."""
    import hashlib

    import torch

    dev = pipeline.device
    g = torch.Generator(device="cpu").manual_seed(seed)          # CPU generator: the same samples on every rank
    n = pipeline.feature_extractor.n_samples
    pcm = (0.1 * torch.randn((n_chunks, n), generator=g)).clamp_(-1, 1)
    pcm[-1] = 0.0
    nv = torch.full((n_chunks,), n, dtype=torch.int32)
    if n_chunks > 1:
        nv[1] = 3 * 16000 + 77
    hidden = pipeline.encode_device(pcm.to(dev), n_valid=nv.to(dev))
    torch.cuda.synchronize(dev)
    digest = hashlib.sha256(hidden.view(torch.int16).cpu().numpy().tobytes()).hexdigest()
    return digest, len(set(gather_all(digest))) == 1


def gather_all(obj):
    """All-gather of a picklable per-rank object over the control group (a one-element list without one)."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
