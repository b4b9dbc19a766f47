"""CPU tests of the oracle against the weight-free known answers in the reference's unit tests
 and against an independent implementation of the same model family
(transformers' qwen3_omni_moe Code2Wav blocks) for block semantics."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import generate as OG
from oracle import model as OM
from oracle import sampling as osmp
from oracle import vocoder as OV
from qwen3_tts_rs_b200 import spec as S, weights as W
from conftest import talker_weights, vocoder_weights


# ---- causal conv / trans conv / snake (causal_conv.rs:154-238, causal_trans_conv.rs:119-197, snake_beta.rs:107-132)
@pytest.mark.parametrize("k,dil", [(1, 1), (3, 1), (7, 1), (7, 3), (7, 9)])
def test_causal_conv_length_and_causality(k, dil):
    g = torch.Generator().manual_seed(0)
    w, b = torch.randn(5, 4, k, generator=g), torch.randn(5, generator=g)
    x = torch.randn(2, 4, 20, generator=g)
    y = OV.causal_conv1d(x, w, b, dil)
    assert y.shape == (2, 5, 20)
    x2 = x.clone()
    x2[:, :, 12:] += 1.0                      # changing the future must not change the past
    y2 = OV.causal_conv1d(x2, w, b, dil)
    assert torch.equal(y[:, :, :12], y2[:, :, :12]) and not torch.equal(y[:, :, 12:], y2[:, :, 12:])


@pytest.mark.parametrize("k,s", [(16, 8), (10, 5), (8, 4), (6, 3), (4, 2), (2, 2)])
def test_causal_trans_conv_lengths(k, s):
    x = torch.randn(1, 3, 7)
    w = torch.randn(3, 2, k)
    y = OV.causal_trans_conv1d(x, w, None, s)
    assert y.shape == (1, 2, 7 * s)           # right_trim = k - s (causal_trans_conv.rs:86)
    x2 = x.clone()
    x2[:, :, 4:] += 1.0
    y2 = OV.causal_trans_conv1d(x2, w, None, s)
    assert torch.equal(y[:, :, : 4 * s], y2[:, :, : 4 * s])


def test_snake_beta_zero_params():
    x = torch.linspace(-3, 3, 50).reshape(1, 2, 25)
    y = OV.snake_beta(x, torch.zeros(2), torch.zeros(2))
    assert torch.allclose(y, x + torch.sin(x) ** 2, atol=1e-6)     # alpha = beta = 0 -> x + sin^2 x


def test_upsample_total_and_stage_shapes():
    """decoder_12hz.rs:713-722 (total 1920) and the T=2 stage shapes asserted in reference_validation.rs
    (SURVEY.md §4): quantized [1,512,2], pre_conv [1,1024,2], pre_transformer [1,2,512], output_proj
    [1,2,1024], upsample_0 [1,1024,4], decoder.0 [1,1536,8], decoder.1 [1,768,64]; 1 frame -> 1920, 2 -> 3840."""
    vs = S.VocoderSpec()
    assert vs.total_upsample == 1920
    vw = vocoder_weights(vs, "full")
    voc = OV.Vocoder(vs, vw)
    st = {}
    out = voc.decode(np.zeros((1, 16, 2), dtype=np.int64), st)
    assert out.shape == (1, 1, 3840)
    assert tuple(st["quantized"].shape) == (1, 512, 2)
    assert tuple(st["pre_conv"].shape) == (1, 1024, 2)
    assert tuple(st["pre_transformer"].shape) == (1, 2, 512)
    assert tuple(st["output_proj"].shape) == (1, 2, 1024)
    assert tuple(st["upsample_0"].shape) == (1, 1024, 4)
    assert tuple(st["decoder.0"].shape) == (1, 1536, 8)
    assert tuple(st["decoder.1"].shape) == (1, 768, 64)
    assert voc.decode(np.zeros((1, 16, 1), dtype=np.int64)).shape == (1, 1, 1920)
    assert float(out.abs().max()) <= 1.0


def test_decoder_block_length():                  # decoder_block.rs:298-376
    vs = S.TINY_VOCODER
    vw = vocoder_weights(vs, "tiny")
    x = torch.randn(1, vs.decoder_dim, 5)
    y = OV.decoder_block(x, vw, "decoder.decoder.1.block", 8)
    assert y.shape == (1, vs.decoder_dim // 2, 40)


def test_vocoder_blocks_match_transformers_code2wav():
    """Independent cross-check of block semantics against transformers' qwen3_omni_moe (same model family):
    SnakeBeta, the causal conv net, the causal transposed conv (right trim), the residual unit and the
    ConvNeXt block.  Where the two could differ the Rust reference wins, so only blocks whose
    definitions coincide are compared."""
    mod = pytest.importorskip("transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe")
    g = torch.Generator().manual_seed(1)
    # SnakeBeta
    sb = mod.SnakeBeta(6)
    with torch.no_grad():
        sb.alpha.copy_(0.1 * torch.randn(6, generator=g))
        sb.beta.copy_(0.1 * torch.randn(6, generator=g))
    x = torch.randn(2, 6, 11, generator=g)
    assert torch.allclose(sb(x), OV.snake_beta(x, sb.alpha.detach(), sb.beta.detach()), atol=1e-6)
    # causal conv, dilation 3
    cc = mod.Qwen3OmniMoeCausalConvNet(6, 4, kernel_size=7, dilation=3)
    y = cc(x)
    assert torch.allclose(y, OV.causal_conv1d(x, cc.conv.weight.detach(), cc.conv.bias.detach(), 3), atol=1e-5)
    # causal transposed conv k = 2*s
    tc = mod.Qwen3OmniMoeCausalTransConvNet(6, 4, 10, 5)
    y = tc(x)
    mine = OV.causal_trans_conv1d(x, tc.conv.weight.detach(), tc.conv.bias.detach(), 5)
    # transformers trims k - s from BOTH sides (50 samples); the Rust reference trims the right side only and
    # keeps exactly T*s samples (causal_trans_conv.rs:86-100, docs/VALIDATION.md:79-96) -- the reference wins;
    # the two agree on the overlapping samples.
    assert mine.shape == (2, 4, 55) and y.shape == (2, 4, 50) and torch.allclose(y, mine[:, :, 5:], atol=1e-5)
    # residual unit
    ru = mod.Qwen3OmniMoeCode2WavDecoderResidualUnit(6, 9)
    w = {"u.act1.alpha": ru.act1.alpha.detach(), "u.act1.beta": ru.act1.beta.detach(),
         "u.conv1.conv.weight": ru.conv1.conv.weight.detach(), "u.conv1.conv.bias": ru.conv1.conv.bias.detach(),
         "u.act2.alpha": ru.act2.alpha.detach(), "u.act2.beta": ru.act2.beta.detach(),
         "u.conv2.conv.weight": ru.conv2.conv.weight.detach(), "u.conv2.conv.bias": ru.conv2.conv.bias.detach()}
    xx = torch.randn(1, 6, 40, generator=g)
    assert torch.allclose(ru(xx), OV.residual_unit(xx, w, "u", 9), atol=1e-5)
    # ConvNeXt block
    cn = mod.Qwen3OmniMoeConvNeXtBlock(6)
    with torch.no_grad():
        cn.gamma.copy_(0.1 + 0.02 * torch.randn(6, generator=g))
    w = {"c.dwconv.conv.weight": cn.dwconv.conv.weight.detach(), "c.dwconv.conv.bias": cn.dwconv.conv.bias.detach(),
         "c.norm.weight": cn.norm.weight.detach(), "c.norm.bias": cn.norm.bias.detach(),
         "c.pwconv1.weight": cn.pwconv1.weight.detach(), "c.pwconv1.bias": cn.pwconv1.bias.detach(),
         "c.pwconv2.weight": cn.pwconv2.weight.detach(), "c.pwconv2.bias": cn.pwconv2.bias.detach(),
         "c.gamma": cn.gamma.detach()}
    assert torch.allclose(cn(xx), OV.convnext_block(xx, w, "c"), atol=1e-5)


# ---- transformer pieces ---------------------------------------------------------------------------------
def test_fused_equals_sequential_rmsnorm_f32():      # fused_ops.rs:269-313
    g = torch.Generator().manual_seed(3)
    x, r = torch.randn(4, 64, generator=g), torch.randn(4, 64, generator=g)
    w = 1 + 0.1 * torch.randn(64, generator=g)
    n1, s1 = OM.fused_residual_rmsnorm(OM.F32P, x, r, w, 1e-6, cuda_kernel_semantics=True)
    n2, s2 = OM.fused_residual_rmsnorm(OM.F32P, x, r, w, 1e-6, cuda_kernel_semantics=False)
    assert torch.equal(s1, x + r) and torch.allclose(n1, n2, atol=1e-5)
    assert torch.allclose(n1, OM.rms_norm(OM.F32P, x + r, w, 1e-6), atol=1e-6)


def test_fused_bf16_quirk_sum_of_unrounded():
    """fused_residual_rmsnorm.cu:60-65,86: sum of squares from the un-rounded f32 sum, output from the rounded one."""
    x = torch.tensor([[1.0, 2.0 ** -9]]).to(torch.bfloat16).float()
    r = torch.tensor([[2.0 ** -9, 1.0]]).to(torch.bfloat16).float()
    w = torch.ones(2)
    n, s = OM.fused_residual_rmsnorm(OM.BF16P, x, r, w, 0.0)
    assert torch.equal(s, torch.tensor([[1.0, 1.0]]))           # 1 + 2^-9 rounds to 1 in bf16
    exact = (1 + 2.0 ** -9)
    assert torch.allclose(n, OM.BF16P.r(torch.tensor([[1.0, 1.0]]) / exact), atol=0)


def test_kv_cache_semantics():                       # kv_cache.rs:378-389, 293-300, 350-352
    c = OM.KVCache(max_seq=7)
    k1 = torch.randn(1, 2, 4, 16)
    k, v = c.update(k1, k1)
    assert k.shape == (1, 2, 4, 16)
    k, v = c.update(torch.randn(1, 2, 3, 16), torch.randn(1, 2, 3, 16))
    assert k.shape == (1, 2, 7, 16) and torch.equal(k[:, :, :4], k1)
    with pytest.raises(RuntimeError, match="KV cache overflow"):
        c.update(torch.randn(1, 2, 1, 16), torch.randn(1, 2, 1, 16))
    c.reset()
    assert len(c) == 0


def test_causal_mask():                              # transformer.rs:21-36
    m = OM.causal_mask(3, 2)[0, 0]
    assert m.shape == (3, 5)
    assert (m[0] == torch.tensor([0, 0, 0, float("-inf"), float("-inf")])).all()
    assert (m[2] == 0).all()


def test_rope_is_rotation_and_position_zero_is_identity():
    cos, sin = OM.rope_cos_sin([0, 5], 128, 1e6)
    x = torch.randn(1, 2, 2, 128)
    y = OM.apply_rope_rotation(OM.F32P, x, cos, sin)
    assert torch.allclose(y[:, :, 0], x[:, :, 0], atol=1e-6)
    assert torch.allclose(y.norm(dim=-1), x.norm(dim=-1), atol=1e-4)
    assert float(OM.inv_freq(128, 1e6)[0]) == 1.0


def test_prefill_then_steps_equals_full_prefill_f32():
    """Decode steps over the KV cache reproduce a longer causal prefill (the cache/rope/mask plumbing)."""
    spec = S.SPEC_TINY
    tk, _ = OM.Talker(spec, talker_weights(spec), OM.F32P), None
    g = torch.Generator().manual_seed(5)
    emb = 0.05 * torch.randn(1, 6, spec.hidden, generator=g)
    c1 = tk.new_kv_caches()
    h_full, logits_full = tk.run_prefill_layers(emb, c1)
    c2 = tk.new_kv_caches()
    tk.run_prefill_layers(emb[:, :4], c2)
    h4, _ = tk.generate_step_with_embed(emb[:, 4:5], c2, 4)
    h5, l5 = tk.generate_step_with_embed(emb[:, 5:6], c2, 5)
    assert torch.allclose(h5, h_full[:, 5:6], atol=2e-5) and torch.allclose(l5, logits_full, atol=2e-4)


def test_custom_voice_prefill_is_10_positions_and_voice_design_layout():
    """talker.rs:437-449: 3 role + 6 codec/tts + 1 first-text = 10 positions whatever the text length;
    VoiceDesign: N_instruct + 9 (talker.rs:566-583)."""
    spec = S.SPEC_TINY
    tk = OM.Talker(spec, talker_weights(spec), OM.BF16P)
    for n in (1, 5, 40):
        e = tk.custom_voice_embeds(list(range(n)), S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"])
        assert e.shape == (1, 10, spec.hidden)
    e = tk.voice_design_embeds([1, 2, 3], list(range(24)), S.LANGUAGE_IDS["english"])
    assert e.shape == (1, 24 + 9, spec.hidden)
    t, n, pad = tk.build_trailing_text([7])
    assert n == 1 and torch.equal(t, tk.tts_eos_embed())            # lib.rs:509-515
    t, n, pad = tk.build_trailing_text([7, 8, 9])
    assert n == 3 and torch.equal(t[:, 2:3], tk.tts_eos_embed())


def test_generate_loop_semantics_f32():
    """generate_codes (lib.rs:530-656): frames are [semantic, a0..a14]; acoustic codes < 2048; semantic ids never
    in the suppressed range; frame count == max_new_tokens without EOS; seeded determinism; F32 and BF16 modes
    both run."""
    spec = S.SPEC_TINY
    w = talker_weights(spec)
    ids = W.synthetic_prompt(1, spec)
    for prec in (OM.F32P, OM.BF16P):
        tk, cp = OM.Talker(spec, w, prec), OM.CodePredictor(spec, w, prec)
        emb = tk.custom_voice_embeds(ids, S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"])
        cfg = osmp.GenerationConfig(max_new_tokens=5)
        a = OG.prefill_and_generate(tk, cp, emb, ids, cfg, 42)
        b = OG.prefill_and_generate(tk, cp, emb, ids, cfg, 42)
        assert a == b and len(a) == 5 and all(len(f) == 16 for f in a)
        assert all(f[0] < 2048 or f[0] == 2150 for f in a) and all(max(f[1:]) < 2048 for f in a)
        c = OG.prefill_and_generate(tk, cp, emb, ids, cfg, 43)
        assert c != a


def test_forced_eos_stops_the_loop():
    """+40 on logit 2150 at frame 3 -> the loop stops; the EOS frame is not emitted."""
    spec = S.SPEC_TINY
    w = talker_weights(spec)
    tk, cp = OM.Talker(spec, w, OM.F32P), OM.CodePredictor(spec, w, OM.F32P)
    ids = W.synthetic_prompt(0, spec)
    emb = tk.custom_voice_embeds(ids, S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"])

    def hook(frame_idx, logits):
        if frame_idx == 3:
            logits = logits.copy()
            logits[:, 2150] += 40.0
        return logits
    fr = OG.prefill_and_generate(tk, cp, emb, ids, osmp.GenerationConfig(max_new_tokens=10), 42, logit_hook=hook)
    assert len(fr) == 4 and all(f[0] != 2150 for f in fr)


def test_streaming_session_chunks():
    """lib.rs:1650-1759: chunk_frames buffering, flush of the remainder, frame count equals non-streaming."""
    spec = S.SPEC_TINY
    w = talker_weights(spec)
    tk, cp = OM.Talker(spec, w, OM.F32P), OM.CodePredictor(spec, w, OM.F32P)
    ids = W.synthetic_prompt(2, spec)
    emb = tk.custom_voice_embeds(ids, S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"])
    cfg = osmp.GenerationConfig(max_new_tokens=7)
    s = OG.StreamingSession(tk, cp, lambda c: np.zeros(c.shape[2] * 1920, np.float32), emb, ids, cfg, 5, chunk_frames=3)
    chunks = list(s)
    assert [len(c) // 1920 for c in chunks] == [3, 3, 1]
    ref = OG.prefill_and_generate(tk, cp, emb, ids, cfg, 5)
    assert s.all_frames == ref


def test_streaming_left_context_restores_the_non_streamed_waveform():
    """The opt-in left-context streaming mode (q3_session_set_stream_context) on the oracle: every
    vocoder op is causal, so re-decoding the whole history in front of each chunk and dropping its samples gives the
    non-streamed waveform, while the reference's stateless chunks (left_context = 0) differ from it at chunk starts."""
    spec = S.SPEC_TINY
    w = talker_weights(spec)
    tk, cp = OM.Talker(spec, w, OM.F32P), OM.CodePredictor(spec, w, OM.F32P)
    voc = OV.Vocoder(spec.vocoder, vocoder_weights(spec.vocoder, spec.name))
    dec = lambda c: voc.decode(c)[0, 0].numpy()
    ids = W.synthetic_prompt(2, spec)
    emb = tk.custom_voice_embeds(ids, S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"])
    cfg = osmp.GenerationConfig(max_new_tokens=7)
    runs = {}
    for lc in (0, 2, -1):
        s = OG.StreamingSession(tk, cp, dec, emb, ids, cfg, 5, chunk_frames=3, left_context=lc)
        chunks = list(s)
        assert [len(c) // 1920 for c in chunks] == [3, 3, 1]
        runs[lc] = np.concatenate(chunks)
    whole = dec(OG.codes_to_tensor(s.all_frames))
    rms = float(np.sqrt(np.mean(whole ** 2)))
    err = {lc: float(np.sqrt(np.mean((runs[lc] - whole) ** 2))) for lc in runs}
    assert err[-1] <= 1e-6 * max(rms, 1e-3)                        # whole history: the non-streamed waveform
    assert np.array_equal(runs[0][: 3 * 1920], runs[-1][: 3 * 1920])   # the first chunk has no history either way
    assert err[0] > 1e-3 * rms                                     # stateless chunks do differ (the reference's behaviour)
    assert err[2] < err[0]                                         # a bounded context already removes most of it


def test_vocoder_back_half_receptive_field():
    """How much left context an exact streamed decode needs.  The front half of the vocoder (RVQ, pre_conv, pre-transformer)
    runs at the frame rate and attends to the whole history; the back half is a causal conv stack whose look-back adds up
    to 6/2 + 6/4 + 6/4 (ConvNeXt k7 x2, init conv k7) + 1/4 + 78/32 + 1/32 + 78/160 + 1/160 + 78/640 + 1/640 + 78/1920 +
    6/1920 = 9.4 frames (residual units: k7 with dilation 1, 3, 9 = 78 samples per block; transposed convs k = 2s: one input
    sample).  So: front half over the whole history + back half over the chunk and the 10 frames before it reproduces the
    non-streamed waveform exactly, and 9 frames do not.  (Design input for the stateful streaming vocoder)"""
    v = S.TINY_VOCODER                                              # same kernel sizes, dilations and rates as the full model
    assert (v.upsampling_ratios, v.upsample_rates) == (S.VocoderSpec().upsampling_ratios, S.VocoderSpec().upsample_rates)
    # float64 copy of the weights: in f32 the taps furthest back contribute less than the summation-order noise of the conv
    # routines (~5e-7), which would blur the boundary this test is about
    voc = OV.Vocoder(v, vocoder_weights(v, "tiny"))
    g = torch.Generator().manual_seed(8)
    codes = torch.randint(0, v.codebook_size, (1, 16, 26), generator=g).numpy()
    front = voc.decode_front(codes).double()                        # [1, latent, 26], whole history
    voc.w = {k: t.double() for k, t in voc.w.items()}              # back half in f64 from here on
    whole = voc.decode_back(front)[0, 0]
    f0, t = 20, 6                                                   # stream the last 6 frames
    want = whole[f0 * 1920:]
    err = {}
    for c in (0, 4, 9, 10, 12):
        part = voc.decode_back(front[:, :, f0 - c:])[0, 0][c * 1920:]
        assert part.shape == want.shape
        err[c] = float((part - want).abs().max())
    assert err[10] <= 1e-13 and err[12] <= 1e-13, err               # exact: nothing outside the window is read
    assert err[9] > 1e-11 and err[4] > err[9] and err[0] > err[4], err


def test_fork_rules_of_the_gpu_parity_tests():
    """The exemption rules the GPU tests apply to a free-running fork (tests/helpers.py), exercised on the oracle's own trace
    with hand-made forks: an arg-max fork passes only at a near-tie; a sampled fork (the first token included) passes the
    CDF-window rule only for a neighbour of the draw."""
    from helpers import cdf_window, first_divergence_is_a_near_tie, oracle_cfg, oracle_run
    from qwen3_tts_rs_b200 import api
    spec = S.SPEC_TINY
    opts = api.SynthesisOptions(max_length=6)
    ids = W.synthetic_prompt(1, spec)
    ref, tr, _ = oracle_run(spec, ids, 11, opts, trace=True)
    cfg = oracle_cfg(opts)
    assert first_divergence_is_a_near_tie(ref, ref, tr, cfg)[:2] == (len(ref), True)
    # sampled fork at frame 3: to a CDF neighbour, and to a token far from the draw
    fr = tr.frames[2]
    probe = osmp.SamplingContext(0)
    probe.state = fr["rng_state"]
    tok, win = cdf_window(fr["penalised"][0], cfg, float(probe.rand_f32()))
    assert tok == ref[3][0] and 2 <= len(win) <= 12
    neighbour = next(t for t in sorted(win) if t != tok)
    far = next(t for t in range(2048) if t not in win)
    for other, want in ((neighbour, True), (far, False)):
        got = [list(f) for f in ref]
        got[3][0] = other
        m, ok, why = first_divergence_is_a_near_tie(got, ref, tr, cfg)
        assert (m, ok) == (3, want), why
    # the same at the FIRST token (sampled from the prefill logits): only a neighbour passes
    probe.state = tr.first["rng_state"]
    tok0, win0 = cdf_window(tr.first["penalised"][0], cfg, float(probe.rand_f32()))
    assert tok0 == ref[0][0]
    far0 = next(t for t in range(2048) if t not in win0)
    got = [list(f) for f in ref]
    got[0][0] = far0
    assert first_divergence_is_a_near_tie(got, ref, tr, cfg)[:2] == (0, False)
    # arg-max fork at a code whose top-2 margin is wide: refused by both rules
    got = [list(f) for f in ref]
    g = max(range(15), key=lambda i: float(torch.topk(tr.frames[1]["cp_logits"][i].float(), 2).values.diff().abs()))
    got[1][g + 1] = (ref[1][g + 1] + 1) % 2048
    assert first_divergence_is_a_near_tie(got, ref, tr, cfg)[:2] == (1, False)
    assert first_divergence_is_a_near_tie(ref[:4], ref, tr, cfg)[1] is False       # a length mismatch is not a fork


def test_voice_clone_prompt_shapes_and_overlay_rules():
    """prefill_voice_clone (talker.rs:511-564: 10 positions, 9 in ICL mode; position 7 carries the continuous speaker embedding
    under tts_pad) and build_icl_prompt (talker.rs:646-705): streaming overlay over the codec length with the text remainder as
    trailing text, or tts_pad-padded text and a [tts_pad] trailing; non-streaming form = [text + codec_pad ++ codec + tts_pad]."""
    from qwen3_tts_rs_b200 import spec as S, weights as W
    spec = S.SPEC_TINY_PROJ
    w = W.make_talker_weights(spec)
    tk, cp = OM.Talker(spec, w, OM.F32P), OM.CodePredictor(spec, w, OM.F32P)
    g = torch.Generator().manual_seed(0)
    spk = torch.randn(spec.hidden, generator=g)
    lang = S.LANGUAGE_IDS["english"]
    ids = [5, 6, 7, 8]
    xv = tk.voice_clone_embeds(ids, spk, lang, icl_mode=False)
    icl9 = tk.voice_clone_embeds(ids, spk, lang, icl_mode=True)
    assert xv.shape[1] == 10 and icl9.shape[1] == 9 and torch.equal(xv[:, :9], icl9)
    assert torch.allclose(xv[0, 7], tk.tts_pad_embed()[0, 0] + spk, atol=1e-6)                    # speaker embedding under tts_pad
    cv = tk.custom_voice_embeds(ids, S.SPEAKER_IDS["ryan"], lang)
    keep = [i for i in range(10) if i != 7]
    assert torch.equal(xv[0, keep], cv[0, keep])                                       # only the speaker position differs
    ref = torch.randint(0, 2048, (6, 16), generator=g).tolist()
    rce = OM.sum_ref_codec_embeddings(tk, cp, ref)
    assert rce.shape == (1, 6, spec.hidden)
    want0 = tk.codec_embedding[ref[0][0]] + sum(cp.codec_embeddings[gi - 1][ref[0][gi]] for gi in range(1, 16))
    assert torch.allclose(rce[0, 0], want0, atol=1e-5)
    eos = OM.special_id(spec, S.TTS_EOS)
    # text (3 + 4 + 1 = 8) longer than codec (1 + 6 = 7): overlay 7 positions, 1 trailing row = tts_eos
    emb, tr = tk.build_icl_prompt(ids, [1, 2, 3], rce)
    assert emb.shape[1] == 7 and tr.shape[1] == 1 and torch.allclose(tr, tk.projected_text([eos]), atol=1e-6)
    assert torch.allclose(emb[0, 0], tk.projected_text([1])[0, 0] + tk.codec_embedding[S.CODEC_BOS], atol=1e-6)
    assert torch.allclose(emb[0, 3], tk.projected_text([5])[0, 0] + rce[0, 2], atol=1e-6)
    # text (1 + 1 + 1 = 3) shorter: padded with tts_pad, trailing = [tts_pad]
    emb, tr = tk.build_icl_prompt([5], [1], rce)
    assert emb.shape[1] == 7 and torch.equal(tr, tk.tts_pad_embed())
    assert torch.allclose(emb[0, 2], tk.projected_text([eos])[0, 0] + rce[0, 1], atol=1e-6)
    assert torch.allclose(emb[0, 6], tk.tts_pad_embed()[0, 0] + rce[0, 5], atol=1e-6)
    # non-streaming form
    emb, tr = tk.build_icl_prompt(ids, [1, 2, 3], rce, non_streaming=True)
    assert emb.shape[1] == 8 + 7 and torch.equal(tr, tk.tts_pad_embed())
    assert torch.allclose(emb[0, 0], tk.projected_text([1])[0, 0] + tk.codec_embedding[S.CODEC_PAD], atol=1e-6)
    assert torch.allclose(emb[0, 8], tk.codec_embedding[S.CODEC_BOS] + tk.tts_pad_embed()[0, 0], atol=1e-6)
    full, tr = OM.voice_clone_prompt(tk, cp, ids, spk, lang, ref, [1, 2, 3])
    assert full.shape[1] == 9 + 7
