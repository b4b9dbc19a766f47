"""GPU parity tests for the talker step, the code-predictor frame, prefill / prompt assembly and the
generation loop, through the C ABI, against the oracle on the same seeded synthetic weights.

Tolerances (stated here, used below):
  * bf16 activations: the CUDA kernels and the oracle both accumulate in f32 but in a different order,
    so individual bf16 outputs land on a neighbouring bf16 value now and then and the difference
    propagates through the following layers.  Hidden states and logits are compared element-wise with
    |d| <= 2**-6 |ref| + 2**-4 rms(ref)  (4 bf16 ulp at the tensor's rms) and on average with
    mean|d| <= 2**-7 rms(ref).
  * tokens: an arg-max code must equal the oracle's unless the oracle's own top-2 logit margin is below
    2**-5 |top1| (4 bf16 ulp); a sampled token must equal the oracle's or be a CDF NEIGHBOUR of the oracle's draw
    (its interval of the oracle's own CDF meets [u - 0.05, u + 0.05]: 3-6 candidates of 3072; bf16 logit noise of a
    few ulp moves the CDF by a few percent).  The free-running tests here classify the first fork only; every frame of
    a free-running run is held to the oracle, with no token exemptions, by tests/test_gpu_parity.py.
"""
import numpy as np
import pytest
import torch

from qwen3_tts_rs_b200 import api, spec as S, weights as W
from helpers import bf16_ulp_diff, first_divergence_is_a_near_tie as _first_divergence_is_a_near_tie, gpu_tts, oracle_cfg, oracle_models, oracle_run

pytestmark = pytest.mark.gpu

SPECS = [S.SPEC_TINY, S.SPEC_TINY_PROJ, S.SPEC_MID]


def close_bf16(a, b, what):
    a, b = a.float().flatten(), b.float().flatten()
    rms = float(b.pow(2).mean().sqrt())
    tol = 2.0 ** -6 * b.abs() + 2.0 ** -4 * rms
    bad = ((a - b).abs() > tol)
    assert not bool(bad.any()), f"{what}: {int(bad.sum())}/{a.numel()} outside tolerance, max |d|={float((a-b).abs().max()):.4g}, rms={rms:.4g}"
    assert float((a - b).abs().mean()) <= 2.0 ** -7 * rms, f"{what}: mean |d| {float((a-b).abs().mean()):.4g} vs rms {rms:.4g}"


@pytest.mark.parametrize("spec", SPECS, ids=lambda s: s.name)
def test_prefill_step_cp_teacher_forced(spec):
    """Teacher-forced: every frame the CUDA path gets the ORACLE's inputs (last hidden, semantic token, step
    input) and its outputs are compared with the oracle's."""
    opts = api.SynthesisOptions(max_length=12, seed=42)
    text_ids = W.synthetic_prompt(3, spec)
    frames, tr, emb = oracle_run(spec, text_ids, 42, opts, trace=True)
    assert len(tr.frames) >= 8
    tts = gpu_tts(spec)
    sess = api.Session(tts.model, 1, opts, [42], max_seq=64)
    # prefill from the oracle's embeddings, then check the device-side prompt assembly separately
    sess.prefill_embeds([emb[0]])
    exempt_cp = exempt_total = 0
    for fr in tr.frames:
        codes, lg = sess.code_predictor_frame(fr["cp_in_hidden"][0, 0], [fr["tok"]], want_logits=True)
        ol = fr["cp_logits"].float()                       # [15, V]
        close_bf16(torch.from_numpy(lg[0]), ol, f"cp logits frame {fr['frame']}")
        for g in range(15):
            exempt_total += 1
            if int(codes[0, g]) != fr["codes"][g]:
                top2 = torch.topk(ol[g], 2).values
                margin = float(top2[0] - top2[1])
                assert margin <= 2.0 ** -5 * abs(float(top2[0])) + 1e-6, (fr["frame"], g, margin)
                exempt_cp += 1
                break                                      # later codes depend on this one
        hid, logits = sess.talker_step(fr["step_input"][0, 0])
        close_bf16(hid[0], fr["hidden"][0, 0], f"hidden frame {fr['frame']}")
        close_bf16(torch.from_numpy(logits[0]), torch.from_numpy(fr["logits"][0]), f"logits frame {fr['frame']}")
    assert exempt_cp <= 2, exempt_cp
    sess.close()


@pytest.mark.parametrize("spec", SPECS, ids=lambda s: s.name)
def test_prompt_assembly_and_trailing_text_on_device(spec):
    """q3_prefill_ids / q3_set_trailing_ids (text embedding gather, text projection, codec-embedding add)
    against the oracle's prefill_custom_voice + build_trailing_text, observed through the first sampled
    frames: free-running generation must reproduce the oracle's frames."""
    opts = api.SynthesisOptions(max_length=6, seed=7)
    text_ids = W.synthetic_prompt(5, spec)
    frames, tr, _ = oracle_run(spec, text_ids, 7, opts, trace=True)
    tts = gpu_tts(spec)
    got = tts.generate_codes([text_ids], options=opts, seeds=[7])[0]
    m, ok, why = _first_divergence_is_a_near_tie(got, frames, tr, oracle_cfg(opts))
    print("match", m, why)
    assert ok, (m, why)


@pytest.mark.parametrize("spec", [S.SPEC_TINY, S.SPEC_MID], ids=lambda s: s.name)
def test_generate_free_running_batch_vs_oracle(spec):
    """Free-running loop, batch of 4 utterances with different prompts and seeds: row i is compared with an
    independent oracle run (the reference has no batching).  With random synthetic weights the 2048-way
    arg-max has a near-tie every few frames, so a row may fork; the test requires that every fork happens
    at a position where the oracle's own margin is inside the stated exemption band, and reports the
    match lengths."""
    B, F = 4, 16
    opts = api.SynthesisOptions(max_length=F)
    prompts = [W.synthetic_prompt(i, spec) for i in range(B)]
    seeds = [42 + i for i in range(B)]
    tts = gpu_tts(spec)
    got = tts.generate_codes(prompts, options=opts, seeds=seeds)
    report = []
    for b in range(B):
        ref, tr, _ = oracle_run(spec, prompts[b], seeds[b], opts, trace=True)
        m, ok, why = _first_divergence_is_a_near_tie(got[b], ref, tr, oracle_cfg(opts))
        report.append((m, len(ref), ok, why))
    print("free-running (match_len, oracle_frames, fork_is_near_tie, detail):", report)
    assert all(ok for _, _, ok, _ in report), report
    assert all(len(g) == F for g in got)


def test_batch_rows_are_independent_and_deterministic():
    """Size-independent property: row i of a batch-8 run is bit-identical to a batch-1 run with the same
    prompt and seed, and two identical runs give identical codes (graph replay == eager first frame)."""
    spec = S.SPEC_MID
    B, F = 8, 24
    opts = api.SynthesisOptions(max_length=F)
    prompts = [W.synthetic_prompt(i, spec) for i in range(B)]
    seeds = [1000 + i for i in range(B)]
    tts = gpu_tts(spec)
    a = tts.generate_codes(prompts, options=opts, seeds=seeds)
    b = tts.generate_codes(prompts, options=opts, seeds=seeds)
    assert a == b
    for i in (0, 3, 7):
        single = tts.generate_codes([prompts[i]], options=opts, seeds=[seeds[i]])[0]
        assert single == a[i], i


def test_eos_stops_rows_and_eos_frame_is_not_emitted():
    """EOS path (lib.rs:581-585): with a codec_head whose EOS row dominates, min_new_tokens = 2 forbids EOS
    for the first two samples, the third sample is EOS, so exactly 2 frames are emitted per row and the
    EOS token itself never appears in the codes.  Oracle and CUDA path agree."""
    spec = S.SPEC_TINY
    from conftest import talker_weights
    w = dict(talker_weights(spec))
    head = w["talker.codec_head.weight"].clone().float()
    head[2150] = 0.0
    w2 = dict(w)
    # EOS logit = 40 * mean(|h|)-ish: use the final-norm weight direction so it is large and positive
    tts0 = None
    from oracle import model as OM, generate as OG, sampling as osmp
    tk, cp = OM.Talker(spec, w, OM.BF16P), OM.CodePredictor(spec, w, OM.BF16P)
    ids = W.synthetic_prompt(0, spec)
    emb = tk.custom_voice_embeds(ids, S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"])
    caches = tk.new_kv_caches(64)
    hidden, _ = tk.run_prefill_layers(emb, caches)
    direction = hidden[0, -1] / hidden[0, -1].norm()
    head[2150] = 8.0 * direction          # logit ~ 8*|h| at prefill; later hiddens are similar in norm
    w2["talker.codec_head.weight"] = head.to(torch.bfloat16)
    opts = api.SynthesisOptions(max_length=12)
    tk2, cp2 = OM.Talker(spec, w2, OM.BF16P), OM.CodePredictor(spec, w2, OM.BF16P)
    cfg = osmp.GenerationConfig(max_new_tokens=12)
    ref = OG.prefill_and_generate(tk2, cp2, emb, ids, cfg, 42)
    tts = api.Qwen3TTS.from_weights(spec, w2)
    got = tts.generate_codes([ids, ids], options=opts, seeds=[42, 43])
    assert len(ref) == 2 and ref[0][0] != 2150           # EOS is the third sampled token (min_new_tokens = 2)
    for row in got:
        assert len(row) == 2                             # EOS step equals the oracle's
        assert all(f[0] != 2150 for f in row)
    assert got[0][0] == ref[0]


def test_kv_cache_overflow_is_an_error():
    """kv_cache.rs:293-300: appending past max_seq fails with 'KV cache overflow'."""
    spec = S.SPEC_TINY
    tts = gpu_tts(spec)
    opts = api.SynthesisOptions(max_length=64, seed=1)
    sess = api.Session(tts.model, 1, opts, [1], max_seq=16)
    prompts = [tts.custom_voice_prompt(W.synthetic_prompt(0, spec), "ryan", "english")]
    sess.prefill_ids([p[0] for p in prompts], [p[1] for p in prompts])
    sess.set_trailing_ids([[1, 2, 3]])
    with pytest.raises(api.L.Q3Error) as e:
        sess.generate(32)
    assert e.value.status == "Q3_ERR_KV_OVERFLOW" and "KV cache overflow" in str(e.value)
    sess.close()


def test_missing_weight_is_an_error():
    """decoder_12hz.rs:176-181 style: a missing tensor is reported by name."""
    spec = S.SPEC_TINY
    from conftest import talker_weights
    w = dict(talker_weights(spec))
    del w["talker.model.layers.1.mlp.down_proj.weight"]
    with pytest.raises(api.L.Q3Error) as e:
        api.Qwen3TTS.from_weights(spec, w)
    assert e.value.status == "Q3_ERR_MISSING_WEIGHT" and "layers.1.mlp.down_proj" in str(e.value)


def test_multi_kernel_path_rows_independent_and_matches_oracle_tolerance(monkeypatch):
    """The multi-kernel (CUDA-graph) decode path -- the fallback when the persistent kernel cannot be used, and the
    prefill engine -- is kept honest: with Q3_MEGA=0 a batch of 20 is bit-identical, row by row, to batch-1 runs on the
    same path, and its forks from the oracle happen only at near-ties."""
    spec = S.SPEC_TINY_PROJ
    B, F = 20, 10
    opts = api.SynthesisOptions(max_length=F)
    prompts = [W.synthetic_prompt(i, spec) for i in range(B)]
    seeds = [7 + i for i in range(B)]
    tts = gpu_tts(spec)
    monkeypatch.setenv("Q3_MEGA", "0")
    big = tts.generate_codes(prompts, options=opts, seeds=seeds)
    for i in (0, 5, 19):
        single = tts.generate_codes([prompts[i]], options=opts, seeds=[seeds[i]])[0]
        assert single == big[i], i
    ref, tr, _ = oracle_run(spec, prompts[2], seeds[2], opts, trace=True)
    m, ok, why = _first_divergence_is_a_near_tie(big[2], ref, tr, oracle_cfg(opts))
    assert ok, (m, why)


def test_batches_above_16_run_as_row_groups_on_the_persistent_kernel():
    """Batch 37 = row groups of 16 + 16 + 5 launched back to back on the dataflow kernel: every row equals the batch-1
    run with the same prompt and seed (bit-identical), including rows at group boundaries, and a second run repeats."""
    spec = S.SPEC_MID
    B, F = 37, 18
    opts = api.SynthesisOptions(max_length=F)
    prompts = [W.synthetic_prompt(60 + i, spec) for i in range(B)]
    seeds = [900 + i for i in range(B)]
    tts = gpu_tts(spec)
    a = tts.generate_codes(prompts, options=opts, seeds=seeds)
    assert a == tts.generate_codes(prompts, options=opts, seeds=seeds)
    assert all(len(r) == F for r in a)
    for i in (0, 15, 16, 31, 32, 36):
        single = tts.generate_codes([prompts[i]], options=opts, seeds=[seeds[i]])[0]
        assert single == a[i], i


def test_full_size_1p7b_batch8_properties():
    """BASELINE configs[2] size (1.7B, batch 8): properties that need no oracle -- determinism, row independence
    (row 3 of the batch == a batch-1 run), frame structure, no suppressed ids, frame count == max_length."""
    spec = S.SPEC_1_7B
    tts = gpu_tts(spec)
    B, F = 8, 12
    opts = api.SynthesisOptions(max_length=F)
    prompts = [W.synthetic_prompt(i, spec) for i in range(B)]
    seeds = [42 + i for i in range(B)]
    a = tts.generate_codes(prompts, options=opts, seeds=seeds)
    b = tts.generate_codes(prompts, options=opts, seeds=seeds)
    assert a == b
    single = tts.generate_codes([prompts[3]], options=opts, seeds=[seeds[3]])[0]
    assert single == a[3]
    for row in a:
        assert len(row) <= F and all(len(f) == 16 for f in row)
        assert all((f[0] < 2048) for f in row) and all(max(f[1:]) < 2048 for f in row)


@pytest.mark.parametrize("mega", ["1", "2", "3", "4", "5"])
def test_persistent_kernel_generations_agree_with_the_oracle(monkeypatch, mega):
    """The generations of the persistent frame kernel -- fence-based grid barriers (Q3_MEGA=1), tagged dataflow phases (2),
    the round-1 TMA weight-ring variant (3), the warp-specialised TMA ring of round 2 (4) and the dataflow kernel with a TMA
    prefetch buffer (5) -- are all held to the same bar on a model whose
    dimensions exercise the register-resident, streaming and ring code paths (hidden 2048 / CP hidden 1024):
    free-running forks from the oracle only at near-ties, identical results on a second run, and across the
    16-frame launch boundary (40 frames = 3 launches, so the session's tag counter is carried between launches)."""
    from qwen3_tts_rs_b200 import lib as L
    if mega in ("1", "3") and not L.IS_DEV:
        pytest.skip("historical generation: only in libq3tts_b200_dev.so (run with Q3TTS_LIB=dev)")
    monkeypatch.setenv("Q3_MEGA", mega)
    # the ring generations (3, 4, 5) need every skinny-GEMM K to be a multiple of 1024: SPEC_RING has the 1.7B's matrix shapes
    # (K = 1024, 2048, 3072, 6144; 48-row gate/up tiles) with 2 + 2 layers; generations 1 and 2 run the mid spec as in round 1
    # (on which round 1's generation-3 run silently fell back to generation 2 -- hence the decode_generation() assertion)
    spec = S.SPEC_RING if mega in ("3", "4", "5") else S.SPEC_MID
    B, F = 4, 40
    opts = api.SynthesisOptions(max_length=F)
    prompts = [W.synthetic_prompt(10 + i, spec) for i in range(B)]
    seeds = [77 + i for i in range(B)]
    tts = gpu_tts(spec)
    probe = api.Session(tts.model, B, opts, seeds)
    assert probe.decode_generation() == int(mega), "the requested generation was replaced by a fallback"
    probe.close()
    a = tts.generate_codes(prompts, options=opts, seeds=seeds)
    b = tts.generate_codes(prompts, options=opts, seeds=seeds)
    assert a == b
    assert all(len(r) == F for r in a)
    ref, tr, _ = oracle_run(spec, prompts[1], seeds[1], api.SynthesisOptions(max_length=16), trace=True)
    m, ok, why = _first_divergence_is_a_near_tie([f for f in a[1][:16]], ref, tr, oracle_cfg(api.SynthesisOptions(max_length=16)))
    assert ok, (mega, m, why)
    single = tts.generate_codes([prompts[2]], options=opts, seeds=[seeds[2]])[0]
    assert single == a[2]


def test_batch_16_runs_on_the_persistent_kernel_and_rows_stay_independent():
    """Batches 9..16 stay on the dataflow kernel (code-predictor pass 0 in two row groups of 8): rows of a batch-16
    and of a batch-11 run are bit-identical to batch-1 runs with the same prompt and seed, and a second run repeats."""
    spec = S.SPEC_MID
    F = 20
    opts = api.SynthesisOptions(max_length=F)
    tts = gpu_tts(spec)
    for B in (16, 11):
        prompts = [W.synthetic_prompt(30 + i, spec) for i in range(B)]
        seeds = [500 + i for i in range(B)]
        a = tts.generate_codes(prompts, options=opts, seeds=seeds)
        assert a == tts.generate_codes(prompts, options=opts, seeds=seeds)
        for i in (0, 7, 8, B - 1):
            single = tts.generate_codes([prompts[i]], options=opts, seeds=[seeds[i]])[0]
            assert single == a[i], (B, i)
