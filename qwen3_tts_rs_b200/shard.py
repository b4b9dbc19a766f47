"""Data-parallel sharding of utterances over ranks: contiguous split, weights replicated, no
exchange during decode; one all_gather of per-utterance frame counts (and optionally PCM) at the end.
torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced: the first n_total % world ranks get one extra utterance."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def utterance_seeds(base_seed: int, lo: int, hi: int) -> List[int]:
    """Utterance i always gets seed base_seed + i, whatever rank it lands on (row i == independent run i)."""
    return [base_seed + i for i in range(lo, hi)]


def gather_frame_counts(local_counts: Sequence[int], n_total: int, rank: int, world: int, device="cpu") -> np.ndarray:
    """all_gather of the per-utterance frame counts -> array of length n_total on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return np.asarray(local_counts, dtype=np.int64)
    width = (n_total + world - 1) // world
    buf = torch.full((width,), -1, dtype=torch.int64, device=device)
    buf[: len(local_counts)] = torch.as_tensor(list(local_counts), dtype=torch.int64)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    res = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        res.extend(out[r][: hi - lo].tolist())
    return np.asarray(res, dtype=np.int64)
