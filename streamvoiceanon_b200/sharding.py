"""Multi-GPU plumbing: utterance streams are independent, so the path shards embarrassingly -- "replicas only"
(SURVEY.md section 8e).  One process per GPU, stream s lives on rank s mod G, weights are replicated, and there is
NO collective on the data path; torch.distributed is used only to barrier and to take the max-over-ranks of the
timed region."""
from __future__ import annotations

from typing import List

import torch


def streams_for_rank(n_streams: int, rank: int, world: int) -> List[int]:
    """Stream ids owned by `rank` (stream s -> GPU s mod G)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_streams, world))


def max_over_ranks(values, device="cpu") -> List[float]:
    """Element-wise max of a small list of floats over all ranks (identity without a process group)."""
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def aggregate_frames_per_sec(frames_per_rank: int, world: int, max_elapsed_ms: float) -> float:
    """Whole-job throughput: all ranks' frames over the slowest rank's time."""
    return world * frames_per_rank / (max_elapsed_ms / 1e3)
