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


def _worker(rank, world, port, n_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(n_total, rank, world)
    seeds = shard.utterance_seeds(42, lo, hi)
    local = [_fake_frames(i, s) for i, s in zip(range(lo, hi), seeds)]
    allc = shard.gather_frame_counts(local, n_total, rank, world)
    t = torch.tensor([float(sum(local))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put((allc.tolist(), float(t.item())))
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
    counts, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    expect = [_fake_frames(i, 42 + i) for i in range(n_total)]
    assert counts == expect and total == float(sum(expect))
