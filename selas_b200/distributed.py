"""Multi-GPU plumbing for the one place the path needs it: the scene is replicated on every GPU,
sample passes are sharded across ranks, and ONE reduce of the float accumulation buffers merges the films.

torch.distributed is used as plumbing only (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations


def shard_samples(total_spp: int, rank: int, world: int):
    """Contiguous split of sample indices [0, total_spp) over `world` ranks; returns (first_sample, count).
    The first total_spp % world ranks get one extra sample, so every sample index is rendered exactly once."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(total_spp), world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def shard_iterations(total_iterations: int, rank: int, world: int):
    """VCM shards whole iterations (each pass has its own light paths, hash grid and radius; SelasVcmSettings.firstIteration):
    returns (first_iteration, count) with the same contiguous split as shard_samples."""
    return shard_samples(total_iterations, rank, world)


def reduce_film(film, dst: int = 0):
    """Sum the per-rank accumulation buffers onto rank `dst` (the path's only collective); no-op without a process group."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film
