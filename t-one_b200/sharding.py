"""Multi-GPU host logic: streams are independent recurrences, so the step shards by stream with no data-path
collective.  One process per GPU; torch.distributed is used only for the barrier and for the
max-over-ranks of the timed region."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def owner_of(stream_id: int, world: int) -> int:
    """Static partition: stream -> rank (stream mod G)."""
    return int(stream_id) % int(world)


def local_streams(n_streams: int, rank: int, world: int) -> np.ndarray:
    """Global ids of the streams rank `rank` owns."""
    return np.arange(rank, n_streams, world, dtype=np.int64)


def partition(n_streams: int, world: int) -> List[np.ndarray]:
    return [local_streams(n_streams, r, world) for r in range(world)]


def batches(stream_ids: np.ndarray, max_batch: int) -> List[np.ndarray]:
    """Split a rank's streams into step batches of at most max_batch (sub-batching when latency demands)."""
    return [stream_ids[i:i + max_batch] for i in range(0, len(stream_ids), max_batch)]


def aggregate_throughput(audio_seconds_local: float, elapsed_local: float, dist=None) -> Tuple[float, float, float]:
    """Whole-job throughput = sum over ranks of audio seconds / max over ranks of elapsed time.
    Returns (throughput, total_audio_seconds, max_elapsed).  `dist` is torch.distributed (initialised) or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return audio_seconds_local / elapsed_local, audio_seconds_local, elapsed_local
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    s = torch.tensor([audio_seconds_local], dtype=torch.float64, device=dev)
    t = torch.tensor([elapsed_local], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(s[0] / t[0]), float(s[0]), float(t[0])
