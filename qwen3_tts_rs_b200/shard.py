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


def gather_pcm(local_pcm, local_counts: Sequence[int], all_counts: Sequence[int], rank: int, world: int,
               samples_per_frame: int = 1920, dst: int = 0, device="cpu"):
    """The second (and last) collective of SURVEY.md §8e: PCM rows gathered to rank `dst`.

    local_pcm: f32 [n_local, >= max(local_counts) * samples_per_frame] (numpy or torch; rows zero past their own length);
    all_counts: the result of gather_frame_counts (every rank has it, so every rank sizes the same buffer).
    Returns, on `dst`, a list of n_total 1-D f32 numpy arrays trimmed to each utterance's own length (utterance order,
    whatever rank produced them); None on the other ranks.  63 MB for 32 x 256 frames: bandwidth is irrelevant here."""
    import torch
    import torch.distributed as dist
    n_total = len(all_counts)
    pcm = torch.as_tensor(np.asarray(local_pcm) if not isinstance(local_pcm, torch.Tensor) else local_pcm, dtype=torch.float32)
    trim = lambda row, frames: np.ascontiguousarray(row[: int(frames) * samples_per_frame].cpu().numpy())
    if world == 1:
        return [trim(pcm[i], c) for i, c in enumerate(local_counts)]
    width = (n_total + world - 1) // world
    smax = max(1, int(max(all_counts)) * samples_per_frame)
    buf = torch.zeros((width, smax), dtype=torch.float32, device=device)
    n_local = len(local_counts)
    if n_local:
        w = min(smax, pcm.shape[1])
        buf[:n_local, :w] = pcm[:n_local, :w].to(device)
    # one receive buffer, one device -> host copy on `dst` (a copy per row would synchronise n_total times)
    big = torch.empty((world, width, smax), dtype=torch.float32, device=device) if rank == dst else None
    dist.gather(buf, list(big.unbind(0)) if rank == dst else None, dst=dst)
    if rank != dst:
        return None
    host = big.cpu().numpy()
    out = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        for j in range(hi - lo):
            out.append(np.ascontiguousarray(host[r, j, : int(all_counts[lo + j]) * samples_per_frame]))
    return out


def synthesize_sharded(tts, all_text_ids: Sequence[Sequence[int]], options, base_seed: int, rank: int, world: int,
                       speaker: str = "ryan", language: str = "english", device="cuda", dst: int = 0):
    """Batched multi-utterance synthesis over `world` ranks (BASELINE configs[4]): rank r synthesizes utterances
    shard_range(n, r, world) with seeds base_seed + i on its own GPU (`tts` is that rank's replica of the model), then the
    frame counts are all-gathered and the PCM gathered to `dst`.  -> (all_counts on every rank, list of AudioBuffer-like
    numpy rows on dst / None elsewhere).  No collective runs during decode."""
    n_total = len(all_text_ids)
    lo, hi = shard_range(n_total, rank, world)
    local_ids = [list(t) for t in all_text_ids[lo:hi]]
    if local_ids:
        audio = tts.synthesize_with_voice(local_ids, speaker, language, options, seeds=utterance_seeds(base_seed, lo, hi))
        spf = tts.spec.vocoder.total_upsample
        counts = [len(a) // spf for a in audio]
        width = max(1, max(len(a) for a in audio))
        pcm = np.zeros((len(audio), width), dtype=np.float32)
        for i, a in enumerate(audio):
            pcm[i, : len(a)] = a.samples
    else:
        spf = tts.spec.vocoder.total_upsample
        counts, pcm = [], np.zeros((0, 1), dtype=np.float32)
    all_counts = gather_frame_counts(counts, n_total, rank, world, device=device)
    return all_counts, gather_pcm(pcm, counts, all_counts.tolist(), rank, world, spf, dst, device)
