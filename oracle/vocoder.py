"""Oracle restatement of the 12 Hz acoustic-codec decoder (TEST INFRASTRUCTURE ONLY).

Follows:
  Decoder12Hz::{from_weights, decode, conv1d_1x1, linear_3d, run_transformer, run_layer,
                rms_norm, apply_rope}          src/models/codec/decoder_12hz.rs:185-699
  CausalConv1d::forward                        src/models/codec/causal_conv.rs:94-103
  CausalTransConv1d::forward                   src/models/codec/causal_trans_conv.rs:63-100
  ConvNeXtBlock::forward                       src/models/codec/convnext_block.rs:110-141
  SnakeBeta::forward                           src/models/codec/snake_beta.rs:58-77
  ResidualUnit / DecoderBlock ::forward        src/models/codec/decoder_block.rs:81-92, 240-247

F32 everywhere (src/lib.rs:344-345).  `stages` optionally collects every intermediate the
reference's debug_decoder_stages test names.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from qwen3_tts_rs_b200.spec import VocoderSpec


def causal_conv1d(x, w, b, dilation=1, groups=1):
    """causal_conv.rs:94-103: left zero-pad dilation*(k-1), stride 1."""
    pad = dilation * (w.shape[2] - 1)
    if pad > 0:
        x = F.pad(x, (pad, 0))
    return F.conv1d(x, w, b, stride=1, padding=0, dilation=dilation, groups=groups)


def causal_trans_conv1d(x, w, b, stride):
    """causal_trans_conv.rs:63-100: ConvTranspose1d(pad 0) then right-trim k - stride."""
    out = F.conv_transpose1d(x, w, b, stride=stride)
    trim = max(w.shape[2] - stride, 0)
    if trim > 0:
        out = out[:, :, : out.shape[2] - trim]
    return out


def snake_beta(x, alpha, beta, eps=1e-9):
    """snake_beta.rs:58-77: x + sin^2(x*e^alpha) * recip(e^beta + 1e-9) -- a multiply by the
    reciprocal, not a divide."""
    a = torch.exp(alpha)[None, :, None]
    bb = torch.exp(beta)[None, :, None]
    inv = 1.0 / (bb + eps)
    s = torch.sin(x * a)
    return x + (s * s) * inv


def convnext_block(x, w: Dict[str, torch.Tensor], p: str):
    """convnext_block.rs:110-141."""
    c = x.shape[1]
    h = causal_conv1d(x, w[f"{p}.dwconv.conv.weight"], w[f"{p}.dwconv.conv.bias"], 1, groups=c)
    h = h.transpose(1, 2)
    h = F.layer_norm(h, (c,), w[f"{p}.norm.weight"], w[f"{p}.norm.bias"], eps=1e-6)
    h = h @ w[f"{p}.pwconv1.weight"].t() + w[f"{p}.pwconv1.bias"]
    h = F.gelu(h)                                    # erf GELU (convnext_block.rs:126)
    h = h @ w[f"{p}.pwconv2.weight"].t() + w[f"{p}.pwconv2.bias"]
    h = h * w[f"{p}.gamma"]
    return x + h.transpose(1, 2)


def residual_unit(x, w, p, dilation):
    """decoder_block.rs:81-92."""
    h = snake_beta(x, w[f"{p}.act1.alpha"], w[f"{p}.act1.beta"])
    h = causal_conv1d(h, w[f"{p}.conv1.conv.weight"], w[f"{p}.conv1.conv.bias"], dilation)
    h = snake_beta(h, w[f"{p}.act2.alpha"], w[f"{p}.act2.beta"])
    h = causal_conv1d(h, w[f"{p}.conv2.conv.weight"], w[f"{p}.conv2.conv.bias"], 1)
    return h + x


def decoder_block(x, w, bp, rate):
    """decoder_block.rs:240-247."""
    h = snake_beta(x, w[f"{bp}.0.alpha"], w[f"{bp}.0.beta"])
    h = causal_trans_conv1d(h, w[f"{bp}.1.conv.weight"], w[f"{bp}.1.conv.bias"], rate)
    h = residual_unit(h, w, f"{bp}.2", 1)
    h = residual_unit(h, w, f"{bp}.3", 3)
    h = residual_unit(h, w, f"{bp}.4", 9)
    return h


class Vocoder:
    def __init__(self, spec: VocoderSpec, w: Dict[str, torch.Tensor]):
        self.v = spec
        self.w = {k: t.to(torch.float32) for k, t in w.items()}
        q = "decoder.quantizer"
        eps = 1e-7   # decoder_12hz.rs:199-225: embedding_sum / clamp(cluster_usage, 1e-7)
        cb = lambda pre: self.w[f"{pre}._codebook.embedding_sum"] / self.w[f"{pre}._codebook.cluster_usage"].clamp(min=eps)[:, None]
        self.first_codebook = cb(f"{q}.rvq_first.vq.layers.0")
        self.rest_codebooks = [cb(f"{q}.rvq_rest.vq.layers.{i}") for i in range(spec.num_quantizers - 1)]

    def _rms(self, x, weight):
        """decoder_12hz.rs:675-679: x / sqrt(mean(x^2) + eps) * w."""
        var = (x * x).mean(-1, keepdim=True)
        return (x / torch.sqrt(var + self.v.rms_norm_eps)) * weight

    def _rope(self, x, cos, sin):
        """decoder_12hz.rs:682-691 (rotate-half; cos/sin tiled to head_dim)."""
        h = self.v.head_dim // 2
        rot = torch.cat([-x[..., h:], x[..., :h]], -1)
        return x * cos + rot * sin

    def _transformer(self, hidden, stages=None):
        v, w = self.v, self.w
        b, t, _ = hidden.shape
        i = np.arange(0, v.head_dim, 2, dtype=np.float32)
        inv = torch.from_numpy((np.float32(1.0) / np.power(np.float32(v.rope_theta), (i / np.float32(v.head_dim)).astype(np.float32), dtype=np.float32)).astype(np.float32))
        freqs = torch.arange(t, dtype=torch.float32)[:, None] * inv[None, :]
        cos = freqs.cos().repeat(1, 2)[None, None]
        sin = freqs.sin().repeat(1, 2)[None, None]
        mask = torch.full((t, t), float("-inf")).triu(1)[None, None]
        scale = v.head_dim ** -0.5
        for l in range(v.num_layers):
            p = f"decoder.pre_transformer.layers.{l}"
            n = self._rms(hidden, w[f"{p}.input_layernorm.weight"])
            sh = lambda y: y.reshape(b, t, v.num_heads, v.head_dim).transpose(1, 2)
            q = self._rope(sh(n @ w[f"{p}.self_attn.q_proj.weight"].t()), cos, sin)
            k = self._rope(sh(n @ w[f"{p}.self_attn.k_proj.weight"].t()), cos, sin)
            vv = sh(n @ w[f"{p}.self_attn.v_proj.weight"].t())
            a = (q @ k.transpose(-1, -2)) * scale + mask
            a = torch.softmax(a, -1) @ vv
            a = a.transpose(1, 2).reshape(b, t, v.num_heads * v.head_dim)
            a = (a @ w[f"{p}.self_attn.o_proj.weight"].t()) * w[f"{p}.self_attn_layer_scale.scale"]
            hidden = hidden + a
            n = self._rms(hidden, w[f"{p}.post_attention_layernorm.weight"])
            m = F.silu(n @ w[f"{p}.mlp.gate_proj.weight"].t()) * (n @ w[f"{p}.mlp.up_proj.weight"].t())
            m = (m @ w[f"{p}.mlp.down_proj.weight"].t()) * w[f"{p}.mlp_layer_scale.scale"]
            hidden = hidden + m
            if stages is not None:
                stages[f"transformer_layer_{l}"] = hidden
        return hidden

    def decode(self, codes, stages: Optional[dict] = None) -> torch.Tensor:
        """codes: i64 [B,16,T] -> f32 [B,1,T*1920] in [-1,1] (decoder_12hz.rs:411-505)."""
        return self.decode_back(self.decode_front(codes, stages), stages)

    def decode_front(self, codes, stages: Optional[dict] = None) -> torch.Tensor:
        """First half of `decode`, everything at the frame rate: RVQ de-quantisation, pre_conv, pre-transformer, output
        projection -> [B, latent, T].  (The split into front / back is not in the reference; it exists so that tests can
        measure how much left context the convolutional back half needs -- see test_vocoder_back_half_receptive_field.)"""
        v, w = self.v, self.w
        codes = torch.as_tensor(np.asarray(codes), dtype=torch.long)
        bsz, nq, t = codes.shape
        first = self.first_codebook[codes[:, 0] % v.codebook_size]           # :423-431 (mod 2048)
        first_proj = first @ w["decoder.quantizer.rvq_first.output_proj.weight"][:, :, 0].t()
        rest = torch.zeros(bsz, t, v.vq_dim)
        for i in range(v.num_quantizers - 1):                                   # :438-446
            rest = rest + self.rest_codebooks[i][codes[:, i + 1]]
        rest_proj = rest @ w["decoder.quantizer.rvq_rest.output_proj.weight"][:, :, 0].t()
        quantized = (first_proj + rest_proj).transpose(1, 2)                    # [B,512,T]
        st = stages if stages is not None else {}
        st["quantized"] = quantized
        h = causal_conv1d(quantized, w["decoder.pre_conv.conv.weight"], w["decoder.pre_conv.conv.bias"])
        st["pre_conv"] = h
        h = h.transpose(1, 2)
        h = h @ w["decoder.pre_transformer.input_proj.weight"].t() + w["decoder.pre_transformer.input_proj.bias"]
        st["pre_transformer"] = h
        h = self._transformer(h, stages)
        h = self._rms(h, w["decoder.pre_transformer.norm.weight"])
        h = h @ w["decoder.pre_transformer.output_proj.weight"].t() + w["decoder.pre_transformer.output_proj.bias"]
        st["output_proj"] = h
        return h.transpose(1, 2)

    def decode_back(self, h, stages: Optional[dict] = None) -> torch.Tensor:
        """Second half: the causal convolutional upsampler, [B, latent, T] -> [B, 1, T*1920]."""
        v, w = self.v, self.w
        st = stages if stages is not None else {}
        for s, ratio in enumerate(v.upsampling_ratios):
            p = f"decoder.upsample.{s}"
            h = causal_trans_conv1d(h, w[f"{p}.0.conv.weight"], w[f"{p}.0.conv.bias"], ratio)
            h = convnext_block(h, w, f"{p}.1")
            st[f"upsample_{s}"] = h
        h = causal_conv1d(h, w["decoder.decoder.0.conv.weight"], w["decoder.decoder.0.conv.bias"])
        st["decoder.0"] = h
        for bi, rate in enumerate(v.upsample_rates):
            h = decoder_block(h, w, f"decoder.decoder.{bi + 1}.block", rate)
            st[f"decoder.{bi + 1}"] = h
        h = snake_beta(h, w["decoder.decoder.5.alpha"], w["decoder.decoder.5.beta"])
        h = causal_conv1d(h, w["decoder.decoder.6.conv.weight"], w["decoder.decoder.6.conv.bias"])
        st["pre_clamp"] = h
        return h.clamp(-1.0, 1.0)
