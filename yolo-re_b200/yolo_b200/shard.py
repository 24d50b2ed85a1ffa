"""Image-batch sharding across ranks.  Eval-mode inference has no cross-image
dependency (BN uses running stats, NMS is per image -- src/yolo/utils/nms.py:46), so the batch is
split contiguously, weights are replicated and the data path has NO collective; torch.distributed
is only used to agree on the timing (max over ranks) and to count images."""
from __future__ import annotations

import torch


def shard_range(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of the global batch owned by `rank` (first ranks take the remainder)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def aggregate_throughput(n_images_local: int, elapsed_s_local: float, device=None) -> tuple[float, int, float]:
    """(images/s of the whole job, total images, max elapsed seconds).  Uses all_reduce(SUM) on the
    image count and all_reduce(MAX) on the elapsed time when a process group is initialised."""
    import torch.distributed as dist
    n = torch.tensor([float(n_images_local)], dtype=torch.float64, device=device)
    t = torch.tensor([float(elapsed_s_local)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total, tmax = int(n.item()), float(t.item())
    return (total / tmax if tmax > 0 else 0.0), total, tmax
