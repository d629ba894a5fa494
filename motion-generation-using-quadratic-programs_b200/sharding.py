"""QP-index sharding for the multi-GPU path: the QPs are independent, so ranks take
contiguous index ranges and never exchange problem data; the only collective is the max-over-ranks
of the timed region that bench.py reports."""
from __future__ import annotations


def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the QP indices rank `rank` of `world` owns; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_seed(config_seed: int, rank: int) -> int:
    """Distinct, reproducible workload seed per rank (weak scaling: every rank makes its own batch)."""
    return config_seed + 1000 * rank


def whole_job_rate(units_this_rank: int, seconds_this_rank: float, group=None) -> float:
    """units of ALL ranks / max-over-ranks time (torch.distributed; identity when not initialised)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return units_this_rank / seconds_this_rank
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([seconds_this_rank], dtype=torch.float64, device=dev)
    u = torch.tensor([float(units_this_rank)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(u.item()) / float(t.item())
