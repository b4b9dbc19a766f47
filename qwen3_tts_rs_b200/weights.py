"""Synthetic weight generator (no real checkpoints exist in this environment).

Tensor names are the HuggingFace safetensors names the reference loads
(docs/QWEN3_TTS_ARCHITECTURE.md:431-459; src/models/codec/decoder_12hz.rs:191-381),
so a caller that has the real checkpoint can pass those tensors through the
same `q3_model_set_tensor` calls.

Distributions follow SURVEY.md §8(d): linear/embedding N(0,0.02^2); lm_heads and
codec_head N(0,0.05^2); norm weights 1+N(0,0.02^2); biases N(0,0.01^2); vocoder
conv weights N(0, 1/(C_in*k)); SnakeBeta alpha,beta N(0,0.1^2); layer_scale 0.01;
cluster_usage = 1.  Every tensor is generated from its own generator seeded by
(base_seed, crc32(name)) so the result does not depend on generation order and a
subset can be generated on demand.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterator, Tuple

import torch

from .spec import ModelSpec, VocoderSpec

BASE_SEED = 1234


def _gen(name: str, base_seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((base_seed << 32) ^ zlib.crc32(name.encode()))
    return g


def _normal(name, shape, std, base_seed, mean=0.0):
    t = torch.empty(shape, dtype=torch.float32)
    t.normal_(mean, std, generator=_gen(name, base_seed))
    return t


def talker_tensor_specs(spec: ModelSpec) -> Iterator[Tuple[str, tuple, str]]:
    """(name, shape, kind) for talker + code predictor. kind selects the distribution."""
    h, i = spec.hidden, spec.inter
    yield "talker.model.text_embedding.weight", (spec.text_vocab, spec.text_embed_dim), "linear"
    yield "talker.model.codec_embedding.weight", (spec.codec_vocab, h), "linear"
    yield "talker.text_projection.linear_fc1.weight", (spec.text_embed_dim, spec.text_embed_dim), "linear"
    yield "talker.text_projection.linear_fc1.bias", (spec.text_embed_dim,), "bias"
    yield "talker.text_projection.linear_fc2.weight", (h, spec.text_embed_dim), "linear"
    yield "talker.text_projection.linear_fc2.bias", (h,), "bias"
    for l in range(spec.layers):
        p = f"talker.model.layers.{l}"
        yield f"{p}.input_layernorm.weight", (h,), "norm"
        yield f"{p}.self_attn.q_proj.weight", (spec.q_dim, h), "linear"
        yield f"{p}.self_attn.k_proj.weight", (spec.kv_dim, h), "linear"
        yield f"{p}.self_attn.v_proj.weight", (spec.kv_dim, h), "linear"
        yield f"{p}.self_attn.o_proj.weight", (h, spec.q_dim), "linear"
        yield f"{p}.self_attn.q_norm.weight", (spec.head_dim,), "norm"
        yield f"{p}.self_attn.k_norm.weight", (spec.head_dim,), "norm"
        yield f"{p}.post_attention_layernorm.weight", (h,), "norm"
        yield f"{p}.mlp.gate_proj.weight", (i, h), "linear"
        yield f"{p}.mlp.up_proj.weight", (i, h), "linear"
        yield f"{p}.mlp.down_proj.weight", (h, i), "linear"
    yield "talker.model.norm.weight", (h,), "norm"
    yield "talker.codec_head.weight", (spec.codec_vocab, h), "head"
    # code predictor
    ch, ci = spec.cp_hidden, spec.cp_inter
    qd, kd = spec.cp_heads * spec.head_dim, spec.cp_kv_heads * spec.head_dim
    cp = "talker.code_predictor"
    if spec.has_cp_proj:
        yield f"{cp}.small_to_mtp_projection.weight", (ch, h), "linear"
        yield f"{cp}.small_to_mtp_projection.bias", (ch,), "bias"
    for g in range(spec.groups - 1):
        yield f"{cp}.model.codec_embedding.{g}.weight", (spec.cp_vocab, h), "linear"
    for l in range(spec.cp_layers):
        p = f"{cp}.model.layers.{l}"
        yield f"{p}.input_layernorm.weight", (ch,), "norm"
        yield f"{p}.self_attn.q_proj.weight", (qd, ch), "linear"
        yield f"{p}.self_attn.k_proj.weight", (kd, ch), "linear"
        yield f"{p}.self_attn.v_proj.weight", (kd, ch), "linear"
        yield f"{p}.self_attn.o_proj.weight", (ch, qd), "linear"
        yield f"{p}.self_attn.q_norm.weight", (spec.head_dim,), "norm"
        yield f"{p}.self_attn.k_norm.weight", (spec.head_dim,), "norm"
        yield f"{p}.post_attention_layernorm.weight", (ch,), "norm"
        yield f"{p}.mlp.gate_proj.weight", (ci, ch), "linear"
        yield f"{p}.mlp.up_proj.weight", (ci, ch), "linear"
        yield f"{p}.mlp.down_proj.weight", (ch, ci), "linear"
    yield f"{cp}.model.norm.weight", (ch,), "norm"
    for g in range(spec.groups - 1):
        yield f"{cp}.lm_head.{g}.weight", (spec.cp_vocab, ch), "head"


def vocoder_tensor_specs(v: VocoderSpec) -> Iterator[Tuple[str, tuple, str]]:
    """Names from decoder_12hz.rs:191-381."""
    q = "decoder.quantizer"
    yield f"{q}.rvq_first.vq.layers.0._codebook.embedding_sum", (v.codebook_size, v.vq_dim), "codebook"
    yield f"{q}.rvq_first.vq.layers.0._codebook.cluster_usage", (v.codebook_size,), "ones"
    for i in range(v.num_quantizers - 1):
        yield f"{q}.rvq_rest.vq.layers.{i}._codebook.embedding_sum", (v.codebook_size, v.vq_dim), "codebook"
        yield f"{q}.rvq_rest.vq.layers.{i}._codebook.cluster_usage", (v.codebook_size,), "ones"
    yield f"{q}.rvq_first.output_proj.weight", (v.codebook_dim, v.vq_dim, 1), "conv"
    yield f"{q}.rvq_rest.output_proj.weight", (v.codebook_dim, v.vq_dim, 1), "conv"
    yield "decoder.pre_conv.conv.weight", (v.latent_dim, v.codebook_dim, 3), "conv"
    yield "decoder.pre_conv.conv.bias", (v.latent_dim,), "bias"
    t = "decoder.pre_transformer"
    ad = v.num_heads * v.head_dim
    yield f"{t}.input_proj.weight", (v.hidden_size, v.latent_dim), "fc"
    yield f"{t}.input_proj.bias", (v.hidden_size,), "bias"
    yield f"{t}.output_proj.weight", (v.latent_dim, v.hidden_size), "fc"
    yield f"{t}.output_proj.bias", (v.latent_dim,), "bias"
    for l in range(v.num_layers):
        p = f"{t}.layers.{l}"
        yield f"{p}.input_layernorm.weight", (v.hidden_size,), "norm"
        yield f"{p}.self_attn.q_proj.weight", (ad, v.hidden_size), "fc"
        yield f"{p}.self_attn.k_proj.weight", (ad, v.hidden_size), "fc"
        yield f"{p}.self_attn.v_proj.weight", (ad, v.hidden_size), "fc"
        yield f"{p}.self_attn.o_proj.weight", (v.hidden_size, ad), "fc"
        yield f"{p}.self_attn_layer_scale.scale", (v.hidden_size,), "layer_scale"
        yield f"{p}.post_attention_layernorm.weight", (v.hidden_size,), "norm"
        yield f"{p}.mlp.gate_proj.weight", (v.intermediate_size, v.hidden_size), "fc"
        yield f"{p}.mlp.up_proj.weight", (v.intermediate_size, v.hidden_size), "fc"
        yield f"{p}.mlp.down_proj.weight", (v.hidden_size, v.intermediate_size), "fc"
        yield f"{p}.mlp_layer_scale.scale", (v.hidden_size,), "layer_scale"
    yield f"{t}.norm.weight", (v.hidden_size,), "norm"
    c = v.latent_dim
    for s, ratio in enumerate(v.upsampling_ratios):
        p = f"decoder.upsample.{s}"
        yield f"{p}.0.conv.weight", (c, c, ratio), "tconv"
        yield f"{p}.0.conv.bias", (c,), "bias"
        yield f"{p}.1.dwconv.conv.weight", (c, 1, 7), "conv"
        yield f"{p}.1.dwconv.conv.bias", (c,), "bias"
        yield f"{p}.1.norm.weight", (c,), "norm"
        yield f"{p}.1.norm.bias", (c,), "bias"
        yield f"{p}.1.pwconv1.weight", (4 * c, c), "fc"
        yield f"{p}.1.pwconv1.bias", (4 * c,), "bias"
        yield f"{p}.1.pwconv2.weight", (c, 4 * c), "fc"
        yield f"{p}.1.pwconv2.bias", (c,), "bias"
        yield f"{p}.1.gamma", (c,), "gamma"
    yield "decoder.decoder.0.conv.weight", (v.decoder_dim, v.latent_dim, 7), "conv"
    yield "decoder.decoder.0.conv.bias", (v.decoder_dim,), "bias"
    cin = v.decoder_dim
    for b, rate in enumerate(v.upsample_rates):
        cout = cin // 2
        bp = f"decoder.decoder.{b + 1}.block"
        yield f"{bp}.0.alpha", (cin,), "snake"
        yield f"{bp}.0.beta", (cin,), "snake"
        yield f"{bp}.1.conv.weight", (cin, cout, 2 * rate), "tconv"
        yield f"{bp}.1.conv.bias", (cout,), "bias"
        for u in (2, 3, 4):
            yield f"{bp}.{u}.act1.alpha", (cout,), "snake"
            yield f"{bp}.{u}.act1.beta", (cout,), "snake"
            yield f"{bp}.{u}.conv1.conv.weight", (cout, cout, 7), "conv"
            yield f"{bp}.{u}.conv1.conv.bias", (cout,), "bias"
            yield f"{bp}.{u}.act2.alpha", (cout,), "snake"
            yield f"{bp}.{u}.act2.beta", (cout,), "snake"
            yield f"{bp}.{u}.conv2.conv.weight", (cout, cout, 1), "conv"
            yield f"{bp}.{u}.conv2.conv.bias", (cout,), "bias"
        cin = cout
    yield "decoder.decoder.5.alpha", (cin,), "snake"
    yield "decoder.decoder.5.beta", (cin,), "snake"
    yield "decoder.decoder.6.conv.weight", (1, cin, 7), "conv"
    yield "decoder.decoder.6.conv.bias", (1,), "bias"


def make_tensor(name: str, shape: tuple, kind: str, base_seed: int = BASE_SEED) -> torch.Tensor:
    """f32 tensor for one named weight."""
    if kind == "linear":
        return _normal(name, shape, 0.02, base_seed)
    if kind == "head":
        return _normal(name, shape, 0.05, base_seed)
    if kind == "norm":
        return _normal(name, shape, 0.02, base_seed, mean=1.0)
    if kind == "bias":
        return _normal(name, shape, 0.01, base_seed)
    if kind == "snake":
        return _normal(name, shape, 0.1, base_seed)
    if kind == "ones":
        return torch.ones(shape, dtype=torch.float32)
    if kind == "layer_scale":
        return torch.full(shape, 0.01, dtype=torch.float32)
    if kind == "gamma":
        return _normal(name, shape, 0.02, base_seed, mean=0.1)
    if kind == "codebook":
        return _normal(name, shape, 1.0, base_seed)
    if kind == "conv":      # [C_out, C_in/groups, k]
        fan = shape[1] * shape[2]
        gain = 1.0
        if name.endswith("conv2.conv.weight"):
            gain = 0.25     # residual branch: keeps the activation scale flat across the 12 residual units
        if shape[0] == 1:
            gain = 0.1      # final conv: PCM rms ~0.25 so the [-1,1] clamp is rarely active
        return _normal(name, shape, gain / math.sqrt(fan), base_seed)
    if kind == "tconv":     # [C_in, C_out, k]; each output sample sees C_in * ceil(k/stride) taps
        fan = shape[0] * 2 if shape[2] > 2 else shape[0]
        return _normal(name, shape, 1.0 / math.sqrt(fan), base_seed)
    if kind == "fc":        # [out, in]
        return _normal(name, shape, 1.0 / math.sqrt(shape[1]), base_seed)
    raise ValueError(kind)


def make_talker_weights(spec: ModelSpec, base_seed: int = BASE_SEED,
                        dtype: torch.dtype = torch.bfloat16,
                        skip_text_embedding_rows: bool = False) -> Dict[str, torch.Tensor]:
    """All talker + code-predictor tensors, cast to `dtype` (bf16 is what the CUDA path of the
    reference stores, src/lib.rs:1436-1442)."""
    out = {}
    for name, shape, kind in talker_tensor_specs(spec):
        out[name] = make_tensor(name, shape, kind, base_seed).to(dtype)
    return out


def speaker_tensor_specs(c) -> Iterator[Tuple[str, tuple, str]]:
    """Names from speaker.rs:362-434 (prefix speaker_encoder.*)."""
    p = "speaker_encoder"
    ch, ks = c.enc_channels, c.enc_kernel_sizes
    def conv(name, cout, cin, k):
        yield f"{p}.{name}.weight", (cout, cin, k), "conv"
        yield f"{p}.{name}.bias", (cout,), "bias"
    yield from conv("blocks.0.conv", ch[0], c.mel_dim, ks[0])
    for i in range(1, 4):
        C, cs = ch[i], ch[i] // c.enc_res2net_scale
        yield from conv(f"blocks.{i}.tdnn1.conv", C, C, 1)
        for j in range(c.enc_res2net_scale - 1):
            yield from conv(f"blocks.{i}.res2net_block.blocks.{j}.conv", cs, cs, ks[i])
        yield from conv(f"blocks.{i}.tdnn2.conv", C, C, 1)
        yield from conv(f"blocks.{i}.se_block.conv1", c.enc_se_channels, C, 1)
        yield from conv(f"blocks.{i}.se_block.conv2", C, c.enc_se_channels, 1)
    yield from conv("mfa.conv", ch[4], sum(ch[1:4]), ks[4])
    yield from conv("asp.tdnn.conv", c.enc_attention_channels, ch[4] * 3, 1)
    yield from conv("asp.conv", ch[4], c.enc_attention_channels, 1)
    yield from conv("fc", c.enc_dim, ch[4] * 2, 1)


def make_speaker_weights(c, base_seed: int = BASE_SEED) -> Dict[str, torch.Tensor]:
    """All speaker-encoder tensors, F32."""
    return {name: make_tensor(name, shape, kind, base_seed) for name, shape, kind in speaker_tensor_specs(c)}


def make_vocoder_weights(v: VocoderSpec, base_seed: int = BASE_SEED) -> Dict[str, torch.Tensor]:
    """All vocoder tensors, F32 (the vocoder is F32 on every device, src/lib.rs:344-345)."""
    return {name: make_tensor(name, shape, kind, base_seed)
            for name, shape, kind in vocoder_tensor_specs(v)}


def synthetic_prompt(i: int, spec: ModelSpec, n_text: int | None = None):
    """Utterance i of the synthetic prompt set: text-token ids drawn from
    [0, text_vocab-300) with seed 42+i; n_text_i = 8 + (i*7 mod 57)."""
    if n_text is None:
        n_text = 8 + (i * 7) % 57
    g = torch.Generator(device="cpu")
    g.manual_seed(42 + i)
    hi = min(151643, spec.text_vocab - 300) if spec.text_vocab > 4096 else spec.text_vocab
    return torch.randint(0, hi, (n_text,), generator=g).tolist()
