"""Second witness for the oracle's transformer stack (SURVEY.md §8(c) "independent cross-check available offline").

`transformers` 5.5 ships the Qwen3 dense decoder (`models/qwen3/modeling_qwen3.py`): RMSNorm -> q/k/v with per-head
q_norm / k_norm before RoPE -> rotate-half RoPE -> GQA attention -> o_proj -> residual -> RMSNorm -> SwiGLU -> residual.
That is the architecture the reference's talker and code predictor implement (transformer.rs:247-467), written by
different people.  Loading the SAME synthetic weights into it and into the oracle (F32 mode) must give the same
hidden states and logits up to F32 summation order; a structural mistake in the oracle (norm placement, RoPE
pairing, GQA head mapping, scale, mask, cache offsets) would show up as an O(1) difference.

This pins the op ORDER of the oracle, not candle's bf16 rounding points (nothing offline can: SURVEY §8(c)).
Where HF and the Rust reference differ, the Rust reference wins; for this stack they do not differ.
"""
import pytest
import torch

from qwen3_tts_rs_b200 import spec as S, weights as W
from oracle import model as OM

HF = pytest.importorskip("transformers.models.qwen3.modeling_qwen3")


def _hf_stack(spec, w, prefix, hidden, inter, layers, heads, kv_heads, vocab, embed, head):
    cfg = HF.Qwen3Config(vocab_size=vocab, hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                         num_attention_heads=heads, num_key_value_heads=kv_heads, head_dim=spec.head_dim,
                         rms_norm_eps=spec.rms_eps, rope_theta=spec.rope_theta, attention_bias=False,
                         max_position_embeddings=4096, tie_word_embeddings=False, use_sliding_window=False,
                         attn_implementation="eager")
    cfg.rope_parameters = {"rope_type": "default", "rope_theta": spec.rope_theta}
    m = HF.Qwen3ForCausalLM(cfg).eval().to(torch.float32)
    sd = {"model." + k[len(prefix) + 1:]: v for k, v in w.items()
          if k.startswith(prefix + ".layers.") or k == prefix + ".norm.weight"}
    sd["model.embed_tokens.weight"] = embed
    sd["lm_head.weight"] = head
    m.load_state_dict(sd, strict=True)
    return m


@pytest.mark.parametrize("spec", [S.SPEC_TINY, S.SPEC_TINY_PROJ], ids=lambda s: s.name)
def test_talker_prefill_and_steps_match_hf_qwen3(spec):
    """run_prefill_layers (talker.rs:823-841) and generate_step_with_embed (talker.rs:716-736), F32."""
    w = W.make_talker_weights(spec, dtype=torch.float32)
    m = _hf_stack(spec, w, "talker.model", spec.hidden, spec.inter, spec.layers, spec.heads, spec.kv_heads,
                  spec.codec_vocab, w["talker.model.codec_embedding.weight"], w["talker.codec_head.weight"])
    g = torch.Generator().manual_seed(5)
    emb = 0.05 * torch.randn(1, 12, spec.hidden, generator=g)
    tk = OM.Talker(spec, w, OM.F32P)
    caches = tk.new_kv_caches()
    with torch.no_grad():
        out = m(inputs_embeds=emb[:, :9], use_cache=True)
        hs = m.model(inputs_embeds=emb[:, :9]).last_hidden_state
    h, lg = tk.run_prefill_layers(emb[:, :9], caches)
    scale = float(hs.abs().max())
    assert float((h - hs).abs().max()) <= 2e-5 * scale
    assert float((lg[0, 0] - out.logits[0, -1]).abs().max()) <= 2e-5 * float(out.logits.abs().max())
    past = out.past_key_values
    for pos in range(9, 12):                                      # decode steps over the KV cache, offset = position
        with torch.no_grad():
            o = m(inputs_embeds=emb[:, pos:pos + 1], past_key_values=past, use_cache=True,
                  position_ids=torch.tensor([[pos]]), output_hidden_states=True)
        past = o.past_key_values
        h1, l1 = tk.generate_step_with_embed(emb[:, pos:pos + 1], caches, pos)
        assert float((l1[0, 0] - o.logits[0, -1]).abs().max()) <= 2e-5 * float(o.logits.abs().max()), pos
        assert int(torch.argmax(l1)) == int(torch.argmax(o.logits[0, -1]))


def test_code_predictor_passes_match_hf_qwen3():
    """generate_acoustic_codes (code_predictor.rs:320-416): the two-token causal first pass at offset 0, then 14
    single-token passes at offsets 2..15 over the same caches, greedy arg-max per head -- driven here through HF's
    Qwen3 stack with the oracle's embeddings and heads, and compared code for code and logit for logit."""
    spec = S.SPEC_TINY_PROJ                                       # has small_to_mtp_projection, like the 1.7B
    w = W.make_talker_weights(spec, dtype=torch.float32)
    cp = OM.CodePredictor(spec, w, OM.F32P)
    pre = "talker.code_predictor"
    m = _hf_stack(spec, w, pre + ".model", spec.cp_hidden, spec.cp_inter, spec.cp_layers, spec.cp_heads, spec.cp_kv_heads,
                  spec.cp_vocab, torch.zeros(spec.cp_vocab, spec.cp_hidden), w[pre + ".lm_head.0.weight"])
    g = torch.Generator().manual_seed(11)
    talker_hidden = torch.randn(1, 1, spec.hidden, generator=g)
    sem = w["talker.model.codec_embedding.weight"][123][None, None]
    codes, logits = cp.generate_acoustic_codes(talker_hidden, sem, cp.new_kv_caches(), return_logits=True)
    pw, pb = w[pre + ".small_to_mtp_projection.weight"], w[pre + ".small_to_mtp_projection.bias"]
    proj = lambda x: x @ pw.T + pb
    with torch.no_grad():
        o = m.model(inputs_embeds=proj(torch.cat([talker_hidden, sem], 1)), use_cache=True)
        past = o.past_key_values
        hf_codes, h = [], o.last_hidden_state[:, 1:2]
        for grp in range(15):
            lg = h[0, 0] @ w[f"{pre}.lm_head.{grp}.weight"].T
            assert float((lg - logits[grp]).abs().max()) <= 5e-5 * float(lg.abs().max()), grp
            hf_codes.append(int(torch.argmax(lg)))
            if grp == 14:
                break
            e = w[f"{pre}.model.codec_embedding.{grp}.weight"][hf_codes[-1]][None, None]
            o = m.model(inputs_embeds=proj(e), past_key_values=past, use_cache=True,
                        position_ids=torch.tensor([[grp + 2]]))
            past, h = o.past_key_values, o.last_hidden_state
    assert hf_codes == codes


def test_vocoder_pre_transformer_matches_hf_code2wav():
    """decoder_12hz.rs:536-699 (8-layer pre-transformer: RMSNorm eps 1e-5, rotate-half RoPE, scale after QK^T, full causal
    mask, layer-scale then residual) against transformers' `Qwen3OmniMoeCode2WavTransformerModel` on the same weights.
    HF applies a 72-position sliding window; the Rust reference attends to the full causal prefix (decoder_12hz.rs:
    557-564, 642) and wins -- so HF is given a window wider than the sequence, and the layer scales are set to O(1) so that a
    mistake inside a branch is not hidden behind the 0.01 factor."""
    mm = pytest.importorskip("transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe")
    cc = pytest.importorskip("transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe")
    from oracle import vocoder as OV
    v = S.TINY_VOCODER
    w = dict(W.make_vocoder_weights(v))
    g = torch.Generator().manual_seed(2)
    for k in list(w):
        if k.endswith("_layer_scale.scale"):
            w[k] = 0.5 + torch.rand(w[k].shape, generator=g)
    cfg = cc.Qwen3OmniMoeCode2WavConfig(
        hidden_size=v.hidden_size, num_attention_heads=v.num_heads, num_key_value_heads=v.num_heads,
        intermediate_size=v.intermediate_size, num_hidden_layers=v.num_layers, rms_norm_eps=v.rms_norm_eps,
        sliding_window=4096, rope_parameters={"rope_type": "default", "rope_theta": v.rope_theta})
    cfg.head_dim = v.head_dim
    cfg._attn_implementation = "eager"
    m = mm.Qwen3OmniMoeCode2WavTransformerModel(cfg).eval()
    pre = "decoder.pre_transformer."
    m.load_state_dict({k[len(pre):]: t for k, t in w.items() if k.startswith(pre + "layers.") or k == pre + "norm.weight"},
                      strict=True)
    voc = OV.Vocoder(v, w)
    for t in (1, 7, 100):
        x = torch.randn(2, t, v.hidden_size, generator=g)
        with torch.no_grad():
            y = m(inputs_embeds=x).last_hidden_state
        mine = voc._rms(voc._transformer(x), w[pre + "norm.weight"])
        assert float((y - mine).abs().max()) <= 2e-5 * float(y.abs().max()), t


def test_sampler_filters_match_hf_logits_processors():
    """The reference's sampler algebra (sampling.rs:140-285, 375-400) restated in oracle/sampling.py against transformers'
    independent logits processors on random rows: same repetition-penalty rule (positive logits divided, the rest
    multiplied), same top-k survivor set, same nucleus (a token survives iff the probability mass of the strictly more
    likely tokens is below p; HF states it from the other end of the sorted list).  Rows are continuous random values, so
    ties -- where the reference's `>=` rules keep extras, tested separately -- do not occur."""
    lp = pytest.importorskip("transformers.generation.logits_process")
    import numpy as np
    from oracle import sampling as osmp
    rng = np.random.default_rng(3)
    for trial in range(40):
        v = 3072 if trial % 2 else 2048
        row = (rng.standard_normal(v) * (1.0 + trial % 5)).astype(np.float32)
        t = torch.from_numpy(row)[None]
        # repetition penalty over a random seen-set
        seen = rng.choice(v, size=50, replace=False)
        mask = np.zeros((1, v), dtype=np.float32)
        mask[0, seen] = 1.0
        mine = osmp.apply_repetition_penalty_with_mask(row[None].copy(), mask, 1.05)[0]
        hf = lp.RepetitionPenaltyLogitsProcessor(1.05)(torch.from_numpy(seen)[None], t.clone())[0].numpy()
        assert np.allclose(mine, hf, rtol=3e-7, atol=0)            # x * (1/p) vs x / p: one rounding apart
        # top-k
        k = int(rng.integers(1, 100))
        mine_k = osmp.top_k_filter(row.copy(), k)
        hf_k = lp.TopKLogitsWarper(top_k=k)(None, t.clone())[0].numpy()
        assert np.array_equal(np.isfinite(mine_k), np.isfinite(hf_k)) and int(np.isfinite(mine_k).sum()) == k
        # nucleus, on the temperature-scaled row as the reference applies it (sampling.rs:148-171)
        p = float(rng.choice([0.3, 0.5, 0.9, 0.95]))
        scaled = (row * np.float32(1.0 / 0.9)).astype(np.float32)
        mine_p = np.isfinite(osmp.top_p_filter(scaled.copy(), p, "gpu"))
        hf_p = np.isfinite(lp.TopPLogitsWarper(top_p=p)(None, torch.from_numpy(scaled)[None].clone())[0].numpy())
        diff = np.nonzero(mine_p != hf_p)[0]
        if diff.size:                                              # only a token sitting on the p boundary may differ (f32 cumsum order)
            probs = osmp.softmax_f32(scaled)
            order = np.argsort(-probs)
            excl = np.concatenate([[0.0], np.cumsum(probs[order].astype(np.float64))[:-1]])
            pos = {int(tok): i for i, tok in enumerate(order)}
            assert diff.size == 1 and abs(excl[pos[int(diff[0])]] - p) < 1e-5, (trial, diff, p)
        assert mine_p.sum() >= 1 and mine_p[np.argmax(scaled)]


def test_rvq_dequantisation_matches_hf_mimi_split_rvq():
    """decoder_12hz.rs:199-225, 411-455: codebook = embedding_sum / clamp(cluster_usage), one semantic + 15 acoustic
    codebooks summed per group, a bias-free 1x1 output projection per group, the two groups added -- the Mimi split RVQ.
    transformers' `MimiSplitResidualVectorQuantizer.decode` is an independent implementation of it; cluster_usage is
    drawn away from 1 here so the division is exercised (the synthetic checkpoints use 1)."""
    mm = pytest.importorskip("transformers.models.mimi.modeling_mimi")
    from transformers.models.mimi.configuration_mimi import MimiConfig
    from oracle import generate as OG, vocoder as OV
    v = S.TINY_VOCODER
    g = torch.Generator().manual_seed(4)
    w = dict(W.make_vocoder_weights(v))
    for k in list(w):
        if k.endswith("cluster_usage"):
            w[k] = 0.5 + 2.5 * torch.rand(w[k].shape, generator=g)
    cfg = MimiConfig(codebook_size=v.codebook_size, codebook_dim=v.vq_dim, vector_quantization_hidden_dimension=v.vq_dim,
                     hidden_size=v.codebook_dim, num_quantizers=v.num_quantizers, num_semantic_quantizers=1)
    q = mm.MimiSplitResidualVectorQuantizer(cfg).eval()
    sd = {}
    for grp, hf in (("rvq_first", "semantic_residual_vector_quantizer"), ("rvq_rest", "acoustic_residual_vector_quantizer")):
        n = 1 if grp == "rvq_first" else v.num_quantizers - 1
        for i in range(n):
            sd[f"{hf}.layers.{i}.codebook.embed_sum"] = w[f"decoder.quantizer.{grp}.vq.layers.{i}._codebook.embedding_sum"]
            sd[f"{hf}.layers.{i}.codebook.cluster_usage"] = w[f"decoder.quantizer.{grp}.vq.layers.{i}._codebook.cluster_usage"]
        sd[f"{hf}.output_proj.weight"] = w[f"decoder.quantizer.{grp}.output_proj.weight"]
    missing = q.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all("input_proj" in k or "initialized" in k for k in missing.missing_keys), missing
    codes = torch.randint(0, v.codebook_size, (2, v.num_quantizers, 9), generator=g)
    with torch.no_grad():
        want = q.decode(codes)                                     # [B, codebook_dim, T]
    stages = {}
    OV.Vocoder(v, w).decode(codes.numpy(), stages)
    got = stages["quantized"]
    assert got.shape == want.shape and float((got - want).abs().max()) <= 2e-5 * float(want.abs().max())


def test_code_predictor_wiring_matches_hf_omni_code_predictor():
    """Which table feeds which pass (code_predictor.rs:320-416): pass 0 sees [talker hidden, semantic embedding] and reads
    lm_head[0] at position 1; pass g >= 1 embeds the previous code with codec_embedding[g-1] and reads lm_head[g].
    transformers' Qwen3-Omni talker code predictor (same model family) encodes that wiring inside its own forward
    (`generation_steps`): driven greedily on the oracle's weights it must emit the oracle's 15 codes."""
    mm = pytest.importorskip("transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe")
    cc = pytest.importorskip("transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe")
    spec = S.SPEC_TINY                                             # talker hidden == CP hidden: no projection, as in Omni
    w = W.make_talker_weights(spec, dtype=torch.float32)
    cfg = cc.Qwen3OmniMoeTalkerCodePredictorConfig(
        vocab_size=spec.cp_vocab, hidden_size=spec.cp_hidden, intermediate_size=spec.cp_inter, num_hidden_layers=spec.cp_layers,
        num_attention_heads=spec.cp_heads, num_key_value_heads=spec.cp_kv_heads, head_dim=spec.head_dim,
        rms_norm_eps=spec.rms_eps, rope_parameters={"rope_type": "default", "rope_theta": spec.rope_theta},
        attention_bias=False, num_code_groups=spec.groups, max_position_embeddings=1024, use_sliding_window=False)
    cfg._attn_implementation = "eager"
    m = mm.Qwen3OmniMoeTalkerCodePredictorModelForConditionalGeneration(cfg).eval().to(torch.float32)
    pre = "talker.code_predictor."
    sd = {k[len(pre):]: v for k, v in w.items() if k.startswith(pre)}
    m.load_state_dict(sd, strict=True)
    cp = OM.CodePredictor(spec, w, OM.F32P)
    g = torch.Generator().manual_seed(21)
    for trial in range(3):
        talker_hidden = torch.randn(1, 1, spec.hidden, generator=g)
        sem = w["talker.model.codec_embedding.weight"][100 + trial][None, None]
        codes, logits = cp.generate_acoustic_codes(talker_hidden, sem, cp.new_kv_caches(), return_logits=True)
        with torch.no_grad():
            o = m(inputs_embeds=torch.cat([talker_hidden, sem], 1), use_cache=True)
            hf_codes = [int(torch.argmax(o.logits[0, -1]))]
            assert float((o.logits[0, -1] - logits[0]).abs().max()) <= 5e-5 * float(logits[0].abs().max())
            for step in range(1, spec.groups - 1):
                o = m(input_ids=torch.tensor([[hf_codes[-1]]]), past_key_values=o.past_key_values, use_cache=True,
                      generation_steps=step, position_ids=torch.tensor([[step + 1]]))
                assert float((o.logits[0, -1] - logits[step]).abs().max()) <= 5e-5 * float(logits[step].abs().max()), step
                hf_codes.append(int(torch.argmax(o.logits[0, -1])))
        assert hf_codes == codes, trial


def test_vocoder_back_half_matches_hf_code2wav_stack():
    """The whole convolutional back half of the vocoder (decoder_12hz.rs:457-505: 2 x [transposed conv, ConvNeXt], init conv,
    4 decoder blocks of SnakeBeta + transposed conv + 3 residual units, final SnakeBeta + conv, clamp) against the
    `upsample` + `decoder` stacks of transformers' `Qwen3OmniMoeCode2Wav`, with all 140 tensors loaded BY NAME from the
    oracle's checkpoint layout.  The one known difference: transformers trims a transposed conv on both sides, the Rust
    reference on the right only (causal_trans_conv.rs:86-100) and keeps exactly T*1920 samples -- the reference wins.  Both
    stacks are causal and their right ends coincide, so the waveforms must agree sample for sample once the differing left
    edge (555 samples shorter in transformers) has left the 10-frame receptive field."""
    mm = pytest.importorskip("transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe")
    cc = pytest.importorskip("transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe")
    from oracle import vocoder as OV
    v = S.TINY_VOCODER
    w = W.make_vocoder_weights(v)
    cfg = cc.Qwen3OmniMoeCode2WavConfig(hidden_size=v.latent_dim, num_attention_heads=4, num_key_value_heads=4, intermediate_size=64,
                                        num_hidden_layers=1, decoder_dim=v.decoder_dim, upsample_rates=v.upsample_rates,
                                        upsampling_ratios=v.upsampling_ratios, codebook_size=64, num_quantizers=2)
    m = mm.Qwen3OmniMoeCode2Wav(cfg).eval()
    keys = [k for k in m.state_dict() if k.startswith(("upsample.", "decoder."))]
    assert len(keys) == 140
    res = m.load_state_dict({k: w["decoder." + k] for k in keys}, strict=False)
    assert not res.unexpected_keys and not any(k.startswith(("upsample.", "decoder.")) for k in res.missing_keys)
    g = torch.Generator().manual_seed(3)
    h = torch.randn(1, v.latent_dim, 30, generator=g)
    with torch.no_grad():
        x = h
        for blocks in m.upsample:
            for blk in blocks:
                x = blk(x)
        for blk in m.decoder:
            x = blk(x)
        hf = x.clamp(-1, 1)[0, 0]
    mine = OV.Vocoder(v, w).decode_back(h)[0, 0]
    assert mine.numel() == 30 * 1920 and mine.numel() - hf.numel() == 555
    tail = 18 * 1920                                              # the last 18 frames: 12 frames away from the left edge
    assert float((hf[-tail:] - mine[-tail:]).abs().max()) <= 5e-6
    assert float((hf[:1920] - mine[555:555 + 1920]).abs().max()) > 1e-3    # and the left edge really does differ


def test_text_projection_matches_hf_resize_mlp():
    """TextProjection (talker.rs:294-321): fc1 (bias) -> SiLU -> fc2 (bias), the tensors `talker.text_projection.linear_fc1/2`.
    transformers' Qwen3-Omni talker has the same module under the same tensor names (`Qwen3OmniMoeTalkerResizeMLP`)."""
    mm = pytest.importorskip("transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe")
    from types import SimpleNamespace
    spec = S.SPEC_TINY_PROJ
    w = W.make_talker_weights(spec, dtype=torch.float32)
    cfg = SimpleNamespace(thinker_hidden_size=spec.text_embed_dim,
                          text_config=SimpleNamespace(intermediate_size=spec.text_embed_dim, hidden_size=spec.hidden, hidden_act="silu"))
    m = mm.Qwen3OmniMoeTalkerResizeMLP(cfg).eval()
    m.load_state_dict({k[len("talker.text_projection."):]: v for k, v in w.items() if k.startswith("talker.text_projection.")}, strict=True)
    tk = OM.Talker(spec, w, OM.F32P)
    ids = [3, 77, 1500, 2047]
    x = w["talker.model.text_embedding.weight"][ids][None]
    with torch.no_grad():
        want = m(x)
    got = tk.projected_text(ids)
    assert got.shape == want.shape and float((got - want).abs().max()) <= 2e-6 * float(want.abs().max())
