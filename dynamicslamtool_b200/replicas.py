"""Multi-GPU = independent replicas: a sequence is the unit of work, frames of one
sequence are strictly sequential, there is no collective on the data path. This module holds the
host-side bookkeeping bench.py uses: which sequences a rank owns and how per-rank measurements are
combined (max of times, sum of frames) over torch.distributed - NCCL on the GPU box, gloo in tests."""
from __future__ import annotations


def sequences_for_rank(n_sequences: int, rank: int, world: int) -> list[int]:
    """Sequence s runs on rank (s mod world)."""
    return [s for s in range(n_sequences) if s % world == rank]


def seed_for_sequence(base_seed: int, sequence: int) -> int:
    return base_seed + sequence


def combine(frames_done: int, elapsed_ms: float, device=None):
    """Returns (total frames over all ranks, max elapsed ms over all ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return frames_done, elapsed_ms
    t = torch.tensor([float(frames_done)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    e = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    return int(t.item()), float(e.item())


def aggregate_throughput(frames_done: int, elapsed_ms: float, device=None) -> float:
    total, worst = combine(frames_done, elapsed_ms, device)
    return total / (worst * 1e-3)
