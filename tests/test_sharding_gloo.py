"""N > 1 host logic on CPU: two gloo ranks shard a 13-utterance prompt set, run a stand-in per-utterance
function, and gather the frame counts; the result must equal the single-rank result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qwen3_tts_rs_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_frames(i, seed):
    return (seed * 7 + i * 3) % 50 + 1


class _FakeAudio:
    def __init__(self, samples):
        self.samples = samples

    def __len__(self):
        return len(self.samples)


class _FakeTTS:
    """Stand-in for api.Qwen3TTS: utterance with first token i and seed s -> _fake_frames(i, s) frames of value i."""
    class spec:
        class vocoder:
            total_upsample = 8

    def __init__(self):
        self.seen_seeds = []

    def synthesize_with_voice(self, batch_ids, speaker, language, options, seeds=None):
        self.seen_seeds = list(seeds)
        return [_FakeAudio(np.full(_fake_frames(ids[0], s) * 8, float(ids[0]), dtype=np.float32)) for ids, s in zip(batch_ids, seeds)]


def _worker(rank, world, port, n_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(n_total, rank, world)
    seeds = shard.utterance_seeds(42, lo, hi)
    local = [_fake_frames(i, s) for i, s in zip(range(lo, hi), seeds)]
    allc = shard.gather_frame_counts(local, n_total, rank, world)
    # PCM gather (the second collective of SURVEY 8e): row i is a ramp that encodes (utterance, sample index)
    spf = 8
    pcm = np.zeros((hi - lo, max(local) * spf + 5), dtype=np.float32)          # wider than needed: must be trimmed
    for j, (i, c) in enumerate(zip(range(lo, hi), local)):
        pcm[j, : c * spf] = i * 1000.0 + np.arange(c * spf, dtype=np.float32)
    rows = shard.gather_pcm(pcm, local, allc.tolist(), rank, world, samples_per_frame=spf)
    assert (rows is None) == (rank != 0)
    # the whole sharded call with a stand-in model: rank r only ever sees its own utterances and seeds
    fake = _FakeTTS()
    counts2, rows2 = shard.synthesize_sharded(fake, [[i] * (1 + i % 3) for i in range(n_total)], None, 42, rank, world, device="cpu")
    assert fake.seen_seeds == seeds
    t = torch.tensor([float(sum(local))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put((allc.tolist(), float(t.item()), [r.tolist() for r in rows], counts2.tolist(), [r.tolist() for r in rows2]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_without_overlap():
    for n in (1, 8, 13, 32):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                lo, hi = shard.shard_range(n, r, w)
                seen.extend(range(lo, hi))
                assert 0 <= hi - lo <= (n + w - 1) // w
            assert seen == list(range(n))
    assert shard.shard_range(32, 3, 8) == (12, 16)            # BASELINE configs[4]: 4 utterances per GPU


def test_two_rank_gloo_gather_matches_single_rank():
    n_total, world = 13, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    counts, total, rows, counts2, rows2 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    expect = [_fake_frames(i, 42 + i) for i in range(n_total)]
    assert counts == expect and total == float(sum(expect))
    assert len(rows) == n_total
    for i, (r, c) in enumerate(zip(rows, expect)):
        assert len(r) == c * 8 and r == (i * 1000.0 + np.arange(c * 8)).tolist()
    assert counts2 == expect
    assert [len(r) for r in rows2] == [c * 8 for c in expect] and all(set(r) == {float(i)} for i, r in enumerate(rows2))


def test_gather_pcm_single_rank_trims_rows():
    pcm = np.arange(2 * 40, dtype=np.float32).reshape(2, 40)
    rows = shard.gather_pcm(pcm, [2, 4], [2, 4], 0, 1, samples_per_frame=8)
    assert [len(r) for r in rows] == [16, 32] and rows[1][0] == 40.0
