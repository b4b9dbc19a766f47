"""Pins the oracle's sampler against every weight-free known answer in the reference's own unit tests
(SURVEY.md §8c): src/generation/sampling.rs:441-770, src/generation/tts.rs:76-169, src/lib.rs:2015-2134."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import sampling as s
from oracle import generate as OG

F = np.float32


def test_cumsum():                                     # sampling.rs:441-471
    r = s.cumsum_f32(np.array([0.1, 0.2, 0.3, 0.4], F))
    assert np.allclose(r, [0.1, 0.3, 0.6, 1.0], atol=1e-5)


def test_greedy_sample():                              # sampling.rs:473-496
    assert s.greedy_sample(np.array([[1, 2, 5, 1]], F)).tolist() == [2]
    assert s.greedy_sample(np.array([[1, 5, 2], [3, 1, 2], [1, 2, 10]], F)).tolist() == [1, 0, 2]


def test_sample_very_low_temperature():                # sampling.rs:498-511, 611-624
    cfg = s.GenerationConfig(temperature=0.001)
    assert s.sample(np.array([[1, 10, 2, 1]], F), cfg, s.SamplingContext(42)).tolist() == [1]
    assert s.sample(np.array([[10, 1, 1], [1, 10, 1]], F), cfg, s.SamplingContext(42)).tolist() == [0, 1]


def test_sample_valid_index():                         # sampling.rs:513-538
    cfg = s.GenerationConfig(temperature=0.7, repetition_penalty=1.0, eos_token_id=None)
    assert s.sample(np.ones((1, 4), F), cfg, s.SamplingContext(1))[0] < 4
    cfg1 = s.GenerationConfig(temperature=1.0)
    assert s.sample(np.full((1, 3), 2.0, F), cfg1, s.SamplingContext(2))[0] < 3


def test_repetition_penalty():                         # sampling.rs:540-579, 758-770
    l = np.array([[1, 2, 3]], F)
    assert np.allclose(s.apply_repetition_penalty(l, [0], 1.0), l)
    assert np.allclose(s.apply_repetition_penalty(np.array([[2, 3, 4]], F), [0], 2.0), [[1, 3, 4]])
    assert np.allclose(s.apply_repetition_penalty(np.array([[-2, 3, 4]], F), [0], 2.0), [[-4, 3, 4]])
    assert np.allclose(s.apply_repetition_penalty(np.array([[2, 3, 4, 5]], F), [0, 2], 2.0), [[1, 3, 2, 5]])
    # x == 0 takes the `* penalty` branch: result stays 0
    assert s.apply_repetition_penalty(np.array([[0, 3]], F), [0], 2.0)[0, 0] == 0


def test_multinomial_deterministic_probs():            # sampling.rs:600-609
    for seed in range(10):
        u = s.SamplingContext(seed).rand_f32()
        assert s.multinomial_sample(np.array([0, 1, 0, 0], F), u) == 1


def test_seeded_determinism_and_reset():               # sampling.rs:626-706
    a = [s.SamplingContext(12345).rand_f32() for _ in range(1)]
    c1, c2 = s.SamplingContext(12345), s.SamplingContext(12345)
    v1 = [c1.rand_f32() for _ in range(10)]
    v2 = [c2.rand_f32() for _ in range(10)]
    assert v1 == v2
    c3 = s.SamplingContext(67890)
    assert [c3.rand_f32() for _ in range(10)] != v1
    c = s.SamplingContext(42)
    first, second = c.rand_f32(), c.rand_f32()
    c.reset(42)
    assert (c.rand_f32(), c.rand_f32()) == (first, second)
    cfg = s.GenerationConfig(temperature=1.0)
    l = np.ones((1, 5), F)
    ca, cb = s.SamplingContext(99999), s.SamplingContext(99999)
    assert [int(s.sample(l, cfg, ca)[0]) for _ in range(5)] == [int(s.sample(l, cfg, cb)[0]) for _ in range(5)]
    assert all(0.0 <= v <= 1.0 for v in v1)


def test_top_k_filter():                               # sampling.rs:708-732
    v = s.top_k_filter(np.array([1, 5, 3, 2, 4], F), 3)
    assert v[1] == 5 and v[4] == 4 and v[2] == 3 and np.isneginf(v[0]) and np.isneginf(v[3])
    assert np.allclose(s.top_k_filter(np.array([1, 2, 3], F), 100), [1, 2, 3])
    # ties at the threshold keep extras (`>=`)
    assert np.isfinite(s.top_k_filter(np.array([3, 3, 3, 1], F), 2)).sum() == 3


@pytest.mark.parametrize("mode", ["gpu", "cpu"])
def test_top_p_filter(mode):                           # sampling.rs:734-756
    v = s.top_p_filter(np.array([10, 0, 0, 0], F), 0.9, mode)
    assert v[0] == 10
    kept = np.isfinite(s.top_p_filter(np.ones(4, F), 0.5, mode)).sum()
    assert 2 <= kept <= 4


def test_suppression_mask():                           # tts.rs:76-169
    out = s.apply_token_suppression(np.ones((1, 3072), F), 3072, 2150)[0]
    assert out[0] == 1 and out[2047] == 1
    assert np.isneginf(out[2048]) and np.isneginf(out[2149]) and np.isneginf(out[2151]) and np.isneginf(out[3071])
    assert out[2150] == 1
    out3 = s.apply_token_suppression(np.ones((3, 3072), F), 3072, 2150)
    assert np.isneginf(out3[:, 2048]).all() and (out3[:, 2150] == 1).all()
    m1, m2 = s.build_suppression_mask(3072, 2150), s.build_suppression_mask(3072, 2150)
    assert (m1 == m2).all() and m1.sum() == 1023


def test_update_penalty_mask():                        # lib.rs:2093-2118
    m = np.zeros((1, 3072), F)
    s.update_penalty_mask(m, 42)
    assert m[0, 42] == 1 and m[0, 41] == 0 and m[0, 43] == 0
    m = np.zeros((1, 3072), F)
    s.update_penalty_mask(m, 9999)
    assert m.sum() == 0


def test_generation_penalties_order_and_min_new_tokens():   # lib.rs:1271-1322
    cfg = s.GenerationConfig()
    l = np.zeros((1, 3072), F)
    l[0, 2150] = 5.0
    l[0, 10] = 2.0
    seen = np.zeros((1, 3072), F)
    seen[0, 10] = 1
    sup = s.build_suppression_mask()
    out0 = s.apply_generation_penalties(l, seen, cfg, 0, sup)
    out2 = s.apply_generation_penalties(l, seen, cfg, 2, sup)
    assert np.isneginf(out0[0, 2150]) and out2[0, 2150] == 5.0
    assert out0[0, 10] == F(2.0) * (F(1.0) / F(1.05))
    assert np.isneginf(out0[0, 2048:2150]).all() and np.isneginf(out0[0, 2151:]).all()


def test_codes_to_tensor_layout():                     # lib.rs:2015-2050
    assert OG.codes_to_tensor([]).shape == (1, 16, 0)
    assert OG.codes_to_tensor([[0] * 16]).shape == (1, 16, 1)
    t = OG.codes_to_tensor([list(range(16)), list(range(100, 116))]).flatten()
    assert t[:4].tolist() == [0, 100, 1, 101]


def test_bench_fixture_is_frozen():
    """benches/sampling.rs:12-64 fixture (logits = sin(0.1 i)*5, T=0.9, seed 42): oracle outputs frozen in
    tests/golden/sampler_fixture.json by tests/golden/make_golden.py."""
    import json
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", "sampler_fixture.json")))
    i = np.arange(3072, dtype=F)
    logits = (np.sin(i * F(0.1)) * F(5.0)).astype(F)[None]
    for case in gold["cases"]:
        cfg = s.GenerationConfig(temperature=0.9, top_k=case["top_k"], top_p=case["top_p"], repetition_penalty=1.0)
        ctx = s.SamplingContext(42)
        toks = [int(s.sample(logits, cfg, ctx)[0]) for _ in range(len(case["tokens"]))]
        assert toks == case["tokens"], case
    ctx = s.SamplingContext(42)
    assert [ctx.next_u32() for _ in range(8)] == gold["pcg_seed42_u32"]


def test_c_restatement_matches_python():
    """oracle/c (plain C) == oracle (python) for the integer / byte-exact parts."""
    from oracle import build_ref
    lib = C.CDLL(build_ref.build_c())
    lib.q3o_seed_state.restype = C.c_uint64
    lib.q3o_seed_state.argtypes = [C.c_uint64]
    lib.q3o_pcg_next.restype = C.c_uint32
    lib.q3o_pcg_next.argtypes = [C.POINTER(C.c_uint64)]
    lib.q3o_rand_f32.restype = C.c_float
    lib.q3o_rand_f32.argtypes = [C.POINTER(C.c_uint64)]
    for seed in (0, 1, 42, 2 ** 63 + 12345, 2 ** 64 - 1):
        ctx = s.SamplingContext(seed)
        st = C.c_uint64(lib.q3o_seed_state(seed))
        assert st.value == ctx.state
        for _ in range(50):
            assert lib.q3o_pcg_next(C.byref(st)) == ctx.next_u32()
        st2 = C.c_uint64(ctx.state)
        assert F(lib.q3o_rand_f32(C.byref(st2))) == ctx.rand_f32()
    mask = (C.c_uint8 * 3072)()
    lib.q3o_suppression_mask(3072, 2150, mask)
    assert (np.frombuffer(mask, dtype=np.uint8).astype(bool) == s.build_suppression_mask()).all()
    codes = np.arange(48, dtype=np.uint32).reshape(3, 16)
    out = np.zeros((16, 3), dtype=np.int64)
    lib.q3o_codes_to_tensor(codes.ctypes.data_as(C.c_void_p), 3, out.ctypes.data_as(C.c_void_p))
    assert (out == OG.codes_to_tensor(codes.tolist())[0]).all()
    x = np.array([-2.0, -1.0, -0.5, 0.0, 0.25, 1.0, 3.0], dtype=F)
    pcm = np.zeros(7, dtype=np.int16)
    lib.q3o_pcm16(x.ctypes.data_as(C.c_void_p), 7, pcm.ctypes.data_as(C.c_void_p))
    assert pcm.tolist() == [-32767, -32767, -16383, 0, 8191, 32767, 32767]   # audio/io.rs:143-165
