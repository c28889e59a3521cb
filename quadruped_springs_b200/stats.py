"""Rollout statistics across GPUs.  Envs are independent, so the step path has
no collective; the only exchange is one all-gather of a 16-float vector per
shard, over NCCL on GPUs or gloo in the CPU tests."""
import torch
import torch.distributed as dist

NAMES = ["num_envs", "episodes", "sum_max_height", "max_max_height", "sum_rel_max_height", "sum_max_fwd",
         "max_max_fwd", "sum_max_flight_time", "sum_flip_completion", "sum_return", "sum_length", "terminated", "nonfinite"]
_MAX_ROWS = (3, 6)


def combine(vectors: torch.Tensor) -> dict:
    """[G, 16] per-shard vectors -> global statistics dict"""
    v = vectors.to(torch.float64)
    tot = v.sum(0)
    for r in _MAX_ROWS:
        tot[r] = v[:, r].max()
    out = {k: float(tot[i]) for i, k in enumerate(NAMES)}
    ep = max(out["episodes"], 1.0)
    out.update(
        mean_max_height=out["sum_max_height"] / ep, mean_rel_max_height=out["sum_rel_max_height"] / ep,
        mean_max_fwd=out["sum_max_fwd"] / ep, mean_max_flight_time=out["sum_max_flight_time"] / ep,
        mean_flip_completion=out["sum_flip_completion"] / ep, mean_return=out["sum_return"] / ep,
        mean_length=out["sum_length"] / ep, terminated_fraction=out["terminated"] / ep,
    )
    return out


def gather_rollout_stats(local: torch.Tensor) -> dict:
    """all-gather the shard vectors (if a process group exists) and combine"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        bufs = [torch.empty_like(local) for _ in range(dist.get_world_size())]
        dist.all_gather(bufs, local.contiguous())
        return combine(torch.stack(bufs).cpu())
    return combine(local[None].cpu())


def shard_range(num_envs_total: int, rank: int, world_size: int):
    """contiguous slice of the global env ids owned by `rank`"""
    per = num_envs_total // world_size
    rem = num_envs_total % world_size
    start = rank * per + min(rank, rem)
    return start, per + (1 if rank < rem else 0)
