"""Partitioning of independent frames / streams over the GPUs of one box.

The foveation path has no cross-frame dependency (video_server.cc:287-345: nothing is carried from
one iteration to the next except the gaze-independent grid), so multi-GPU operation is pure
sharding: frame f -> rank f % G, stream s -> rank s % G.  No collective touches pixel data; the only
communication is the barrier and the max-over-ranks of the timed region used for reporting.
"""
from __future__ import annotations

from typing import List, Sequence


def frames_for_rank(n_frames: int, world: int, rank: int) -> List[int]:
    """Round-robin frame ownership: frame f belongs to rank f % world."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, n_frames, world))


def streams_for_rank(n_streams: int, world: int, rank: int) -> List[int]:
    """Serving config: stream s is pinned to GPU s % world for its whole life (its SAT, reduced
    and output buffers live there), exactly one owner per stream."""
    return frames_for_rank(n_streams, world, rank)


def owner_of(index: int, world: int) -> int:
    return index % world


class SessionPlacement:
    """Least-loaded placement of client sessions on the GPUs of one box; mirrors
    include/fov360/session_placement.h (the reference binds every connection to device 0,
    video_server.cc:62-66).  Without disconnects it is the static rule s -> s % G."""

    def __init__(self, device_count: int):
        self._load = [0] * max(1, int(device_count))

    def acquire(self) -> int:
        best = min(range(len(self._load)), key=lambda d: (self._load[d], d))
        self._load[best] += 1
        return best

    def release(self, device: int) -> None:
        if 0 <= device < len(self._load) and self._load[device] > 0:
            self._load[device] -= 1

    def sessions(self, device: int) -> int:
        return self._load[device] if 0 <= device < len(self._load) else 0


def aggregate_throughput(units_per_rank: Sequence[int], seconds_per_rank: Sequence[float]) -> float:
    """Whole-job units/s: all units divided by the slowest rank's time (never a sum of rates)."""
    return float(sum(units_per_rank)) / max(seconds_per_rank)


def reduce_max_seconds(local_seconds: float, dist=None, device=None) -> float:
    """max over ranks of a locally measured duration (torch.distributed, any backend)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_seconds)
    import torch

    t = torch.tensor([local_seconds], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum_int(local: int, dist=None, device=None) -> int:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(local)
    import torch

    t = torch.tensor([local], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
