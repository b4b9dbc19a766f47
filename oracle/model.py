"""Oracle restatement of the talker + code predictor (TEST INFRASTRUCTURE ONLY).

Follows:
  apply_rope_rotation / RotaryEmbedding / MRoPE   src/models/transformer.rs:42-69, 72-105, 112-182
  Attention::forward (matmul path)                src/models/transformer.rs:247-372 (:347-369)
  MLP::forward                                    src/models/transformer.rs:408-413
  DecoderLayer::forward                           src/models/transformer.rs:442-467
  FusedRmsNorm / fused_residual_rmsnorm.cu        src/models/fused_ops.rs:49-96, kernels/...cu:38-90
  KV cache append                                 src/models/kv_cache.rs:290-310
  TalkerModel (prefill builders, step, text proj) src/models/talker.rs:294-321, 437-491, 585-627, 716-841
  CodePredictor::generate_acoustic_codes          src/models/code_predictor.rs:320-416
  get_acoustic_embeddings_sum_from_tensor         src/models/code_predictor.rs:497-519

Precision policy `Prec`: in BF16 mode every value is kept as an f32 tensor holding a
bf16-representable number and is re-rounded after each op the reference executes as its
own candle op.  Inside matmul / rms_norm / softmax the arithmetic is f32 (cuBLAS bf16 GEMM
accumulates in f32; candle's rmsnorm/softmax kernels compute in float).  Assumptions about
candle 0.9 internals (not in /root/reference): see DESIGN.md §Oracle.
Attention uses the reference's matmul path (the `cuda` feature without `flash-attn`), the
only one whose rounding points are fully visible in the reference source.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from qwen3_tts_rs_b200 import spec as S
from qwen3_tts_rs_b200.spec import ModelSpec


class Prec:
    def __init__(self, bf16: bool):
        self.bf16 = bf16

    def r(self, x: torch.Tensor) -> torch.Tensor:
        """Round to the activation dtype (bf16 on the CUDA path, identity on the F32 CPU path)."""
        if self.bf16:
            return x.to(torch.bfloat16).to(torch.float32)
        return x


F32P = Prec(False)
BF16P = Prec(True)


# -- primitive ops -------------------------------------------------------------------------

def linear(p: Prec, x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None):
    """candle_nn::Linear: matmul (f32 accumulate, output rounded) then broadcast_add(bias)."""
    y = p.r(x @ w.t())
    if b is not None:
        y = p.r(y + b)
    return y


def rms_norm(p: Prec, x: torch.Tensor, w: torch.Tensor, eps: float):
    """candle rms_norm (transformer.rs:229-231,429-433): f32 internally, one rounding."""
    ms = (x * x).sum(-1, keepdim=True) / x.shape[-1]
    scale = torch.rsqrt(ms + torch.tensor(eps, dtype=torch.float32))
    return p.r((x * scale) * w)


def fused_residual_rmsnorm(p: Prec, x: torch.Tensor, res: torch.Tensor, w: torch.Tensor, eps: float,
                           cuda_kernel_semantics: bool = True):
    """FusedRmsNorm::forward_residual.  CUDA kernel (fused_residual_rmsnorm.cu:57-89): the sum
    of squares uses the UN-rounded f32 `si = x + r`, while pass 2 re-reads the ROUNDED stored
    sum.  Sequential CPU path (fused_ops.rs:60-69): add (rounded) then rms_norm of that."""
    si = x + res
    s = p.r(si)
    if cuda_kernel_semantics and p.bf16:
        ms = (si * si).sum(-1, keepdim=True) / x.shape[-1]
    else:
        ms = (s * s).sum(-1, keepdim=True) / x.shape[-1]
    scale = torch.rsqrt(ms + torch.tensor(eps, dtype=torch.float32))
    return p.r((s * scale) * w), s


def silu(p: Prec, x: torch.Tensor):
    return p.r(x / (1.0 + torch.exp(-x)))


def inv_freq(head_dim: int, theta: float) -> torch.Tensor:
    """1.0 / (theta as f32).powf(i as f32 / dim as f32), i = 0,2,..  (transformer.rs:79-82,133-136)."""
    i = np.arange(0, head_dim, 2, dtype=np.float32)
    e = (i / np.float32(head_dim)).astype(np.float32)
    v = (np.float32(1.0) / np.power(np.float32(theta), e, dtype=np.float32)).astype(np.float32)
    return torch.from_numpy(v)


def rope_cos_sin(positions: Sequence[int], head_dim: int, theta: float):
    """freqs = pos[:,None] @ inv_freq[None,:] in f32, then cos/sin in f32 (transformer.rs:88-90,169-175)."""
    pos = torch.tensor(list(positions), dtype=torch.float32)
    freqs = pos[:, None] * inv_freq(head_dim, theta)[None, :]
    return freqs.cos(), freqs.sin()


def apply_rope_rotation(p: Prec, x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor):
    """x [B,H,S,D]; cos/sin [S,D/2] cast to the activation dtype BEFORE multiplying
    (transformer.rs:48-57); four products and two sums, each its own (rounded) op (:60-66)."""
    d = x.shape[-1]
    x1, x2 = x[..., : d // 2], x[..., d // 2:]
    c, s = p.r(cos)[None, None], p.r(sin)[None, None]
    a = p.r(p.r(x1 * c) - p.r(x2 * s))
    b = p.r(p.r(x2 * c) + p.r(x1 * s))
    return torch.cat([a, b], dim=-1)


def softmax_last(p: Prec, x: torch.Tensor):
    return p.r(torch.softmax(x, dim=-1))


class KVCache:
    """Append-only cache; PreAllocKVCache semantics (kv_cache.rs:290-310): append at
    current_len, overflow is an error, reset() just rewinds."""

    def __init__(self, max_seq: Optional[int] = None):
        self.k: Optional[torch.Tensor] = None
        self.v: Optional[torch.Tensor] = None
        self.max_seq = max_seq

    def update(self, k: torch.Tensor, v: torch.Tensor):
        new_len = (0 if self.k is None else self.k.shape[2]) + k.shape[2]
        if self.max_seq is not None and new_len > self.max_seq:
            raise RuntimeError(f"KV cache overflow: {new_len} > max={self.max_seq}")
        self.k = k if self.k is None else torch.cat([self.k, k], 2)
        self.v = v if self.v is None else torch.cat([self.v, v], 2)
        return self.k, self.v

    def reset(self):
        self.k = None
        self.v = None

    def __len__(self):
        return 0 if self.k is None else self.k.shape[2]


def causal_mask(seq_len: int, offset: int) -> torch.Tensor:
    """create_causal_mask (transformer.rs:21-36)."""
    total = offset + seq_len
    i = torch.arange(seq_len)[:, None]
    j = torch.arange(total)[None, :]
    m = torch.zeros(seq_len, total)
    m[j > offset + i] = float("-inf")
    return m[None, None]


class LayerWeights:
    def __init__(self, w: Dict[str, torch.Tensor], prefix: str):
        g = lambda n: w[f"{prefix}.{n}"].to(torch.float32)
        self.in_ln = g("input_layernorm.weight")
        self.q = g("self_attn.q_proj.weight")
        self.k = g("self_attn.k_proj.weight")
        self.v = g("self_attn.v_proj.weight")
        self.o = g("self_attn.o_proj.weight")
        self.q_norm = g("self_attn.q_norm.weight")
        self.k_norm = g("self_attn.k_norm.weight")
        self.post_ln = g("post_attention_layernorm.weight")
        self.gate = g("mlp.gate_proj.weight")
        self.up = g("mlp.up_proj.weight")
        self.down = g("mlp.down_proj.weight")


def attention(p: Prec, lw: LayerWeights, x: torch.Tensor, cos, sin, mask, cache: Optional[KVCache],
              heads: int, kv_heads: int, head_dim: int, eps: float):
    """Attention::forward, matmul path (transformer.rs:247-284, 347-371)."""
    b, s, _ = x.shape
    q = linear(p, x, lw.q).reshape(b, s, heads, head_dim)
    k = linear(p, x, lw.k).reshape(b, s, kv_heads, head_dim)
    v = linear(p, x, lw.v).reshape(b, s, kv_heads, head_dim)
    q = rms_norm(p, q, lw.q_norm, eps).transpose(1, 2)          # per-head QK norm, :268-269
    k = rms_norm(p, k, lw.k_norm, eps).transpose(1, 2)
    v = v.transpose(1, 2)
    q = apply_rope_rotation(p, q, cos, sin)
    k = apply_rope_rotation(p, k, cos, sin)
    if cache is not None:
        k, v = cache.update(k, v)
    n_rep = heads // kv_heads
    k = k.repeat_interleave(n_rep, dim=1)                        # repeat_kv, :374-386
    v = v.repeat_interleave(n_rep, dim=1)
    aw = p.r(q @ k.transpose(-1, -2))
    # `* self.scale` is candle's affine op evaluated in the tensor dtype, so on the bf16
    # path the scalar itself is rounded to bf16 (candle-kernels affine.cu; assumption).
    scale = p.r(torch.tensor(1.0 / math.sqrt(head_dim), dtype=torch.float32))
    aw = p.r(aw * scale)
    if mask is not None:
        aw = p.r(aw + mask)
    aw = softmax_last(p, aw)
    out = p.r(aw @ v)
    out = out.transpose(1, 2).reshape(b, s, heads * head_dim)
    return linear(p, out, lw.o)


def mlp(p: Prec, lw: LayerWeights, x: torch.Tensor):
    """MLP::forward (transformer.rs:408-413)."""
    gate = silu(p, linear(p, x, lw.gate))
    up = linear(p, x, lw.up)
    return linear(p, p.r(gate * up), lw.down)


def decoder_layer(p: Prec, lw: LayerWeights, x, cos, sin, mask, cache, heads, kv_heads, head_dim, eps,
                  fused_cuda: bool = True):
    """DecoderLayer::forward (transformer.rs:442-467)."""
    h = rms_norm(p, x, lw.in_ln, eps)
    a = attention(p, lw, h, cos, sin, mask, cache, heads, kv_heads, head_dim, eps)
    normed, hs = fused_residual_rmsnorm(p, a, x, lw.post_ln, eps, cuda_kernel_semantics=fused_cuda)
    m = mlp(p, lw, normed)
    return p.r(hs + m)


def special_id(spec: ModelSpec, tok: int) -> int:
    """Special text-token ids sit at a fixed distance from the end of the text vocab; scaled-down
    test specs keep that distance (real vocab: identity)."""
    if spec.text_vocab == 151936:
        return tok
    return tok - 151936 + spec.text_vocab if tok >= 151643 else tok % (spec.text_vocab - 300)


class Talker:
    """TalkerModel (talker.rs:324-955)."""

    def __init__(self, spec: ModelSpec, w: Dict[str, torch.Tensor], prec: Prec):
        self.spec, self.p = spec, prec
        f = lambda n: w[n].to(torch.float32)
        self.text_embedding = f("talker.model.text_embedding.weight")
        self.codec_embedding = f("talker.model.codec_embedding.weight")
        self.fc1_w, self.fc1_b = f("talker.text_projection.linear_fc1.weight"), f("talker.text_projection.linear_fc1.bias")
        self.fc2_w, self.fc2_b = f("talker.text_projection.linear_fc2.weight"), f("talker.text_projection.linear_fc2.bias")
        self.layers = [LayerWeights(w, f"talker.model.layers.{l}") for l in range(spec.layers)]
        self.norm = f("talker.model.norm.weight")
        self.codec_head = f("talker.codec_head.weight")

    def new_kv_caches(self, max_seq: Optional[int] = None) -> List[KVCache]:
        return [KVCache(max_seq) for _ in range(self.spec.layers)]

    # -- text side (talker.rs:294-321, 851-890) --
    def text_projection(self, x: torch.Tensor):
        h = silu(self.p, linear(self.p, x, self.fc1_w, self.fc1_b))
        return linear(self.p, h, self.fc2_w, self.fc2_b)

    def projected_text(self, ids: Sequence[int]) -> torch.Tensor:
        if len(ids) == 0:
            return torch.zeros(1, 0, self.spec.hidden)
        e = self.text_embedding[torch.tensor(list(ids), dtype=torch.long)][None]
        return self.text_projection(e)

    def tts_pad_embed(self):
        return self.projected_text([special_id(self.spec, S.TTS_PAD)])

    def tts_eos_embed(self):
        return self.projected_text([special_id(self.spec, S.TTS_EOS)])

    def codec_embed(self, ids: Sequence[int]) -> torch.Tensor:
        return self.codec_embedding[torch.tensor(list(ids), dtype=torch.long)][None]

    def build_trailing_text(self, input_ids: Sequence[int]):
        """Qwen3TTS::build_trailing_text (lib.rs:508-519)."""
        if len(input_ids) > 1:
            t = torch.cat([self.projected_text(input_ids[1:]), self.tts_eos_embed()], 1)
        else:
            t = self.tts_eos_embed()
        return t, t.shape[1], self.tts_pad_embed()

    # -- prefill builders --
    def _role_prefix(self):
        return self.projected_text([special_id(self.spec, t) for t in (S.IM_START, S.ASSISTANT, S.NEWLINE)])

    def _tts_pad_bos(self, pad_count: int):
        pad = self.tts_pad_embed()
        bos = self.projected_text([special_id(self.spec, S.TTS_BOS)])
        return torch.cat([pad.expand(1, pad_count, -1), bos], 1)

    def custom_voice_embeds(self, text_tokens: Sequence[int], speaker_id: int, language_id: int):
        """prefill_custom_voice input assembly (talker.rs:451-488)."""
        p = self.p
        codec = self.codec_embed([S.CODEC_THINK, S.CODEC_THINK_BOS, language_id, S.CODEC_THINK_EOS,
                                  speaker_id, S.CODEC_PAD, S.CODEC_BOS])
        hidden = torch.cat([self._role_prefix(), p.r(self._tts_pad_bos(5) + codec[:, :6])], 1)
        if len(text_tokens) > 0:
            first = p.r(self.projected_text(text_tokens[:1]) + codec[:, 6:7])
            hidden = torch.cat([hidden, first], 1)
        return hidden

    def voice_design_embeds(self, text_tokens: Sequence[int], instruct_tokens: Sequence[int], language_id: int):
        """prefill_voice_design input assembly (talker.rs:585-624)."""
        p = self.p
        codec = self.codec_embed([S.CODEC_THINK, S.CODEC_THINK_BOS, language_id, S.CODEC_THINK_EOS,
                                  S.CODEC_PAD, S.CODEC_BOS])
        hidden = torch.cat([self.projected_text(instruct_tokens), self._role_prefix(),
                            p.r(self._tts_pad_bos(4) + codec[:, :5])], 1)
        if len(text_tokens) > 0:
            first = p.r(self.projected_text(text_tokens[:1]) + codec[:, 5:6])
            hidden = torch.cat([hidden, first], 1)
        return hidden

    def voice_clone_embeds(self, text_tokens: Sequence[int], speaker_embed: torch.Tensor, language_id: int, icl_mode: bool):
        """prefill_voice_clone input assembly (talker.rs:511-564): prefill_custom_voice with the discrete speaker token
        replaced by a continuous speaker embedding [hidden]; in ICL mode the (first text + codec_bos) position is omitted
        (9 positions instead of 10)."""
        p = self.p
        pre = self.codec_embed([S.CODEC_THINK, S.CODEC_THINK_BOS, language_id, S.CODEC_THINK_EOS])
        spk = p.r(speaker_embed.to(torch.float32)).reshape(1, 1, self.spec.hidden)     # cast to the compute dtype (lib.rs:930)
        suf = self.codec_embed([S.CODEC_PAD, S.CODEC_BOS])
        codec = torch.cat([pre, spk, suf], 1)
        hidden = torch.cat([self._role_prefix(), p.r(self._tts_pad_bos(5) + codec[:, :6])], 1)
        if not icl_mode and len(text_tokens) > 0:
            first = p.r(self.projected_text(text_tokens[:1]) + codec[:, 6:7])
            hidden = torch.cat([hidden, first], 1)
        return hidden

    def build_icl_prompt(self, target_text_ids: Sequence[int], ref_text_ids: Sequence[int], ref_codec_embeds: torch.Tensor,
                         non_streaming: bool = False):
        """TalkerModel::build_icl_prompt (talker.rs:646-705) -> (icl_embed [1, L, H], trailing [1, Lt, H]).
        Text side: text_proj([ref_text ++ target_text ++ tts_eos]); codec side: [codec_bos ++ ref_codec_embeds].
        Streaming form (the one lib.rs:960-963 calls): element-wise overlay over the codec length; the text that does not
        fit becomes the trailing text, otherwise the text is padded with tts_pad and the trailing text is tts_pad alone."""
        p = self.p
        all_text = list(ref_text_ids) + list(target_text_ids) + [special_id(self.spec, S.TTS_EOS)]
        text = self.projected_text(all_text)
        n_text = text.shape[1]
        codec = torch.cat([self.codec_embed([S.CODEC_BOS]), ref_codec_embeds.to(torch.float32)], 1)
        n_codec = codec.shape[1]
        pad = self.tts_pad_embed()
        if non_streaming:
            text_cp = p.r(text + self.codec_embed([S.CODEC_PAD]).expand(1, n_text, -1))
            codec_tp = p.r(codec + pad.expand(1, n_codec, -1))
            return torch.cat([text_cp, codec_tp], 1), pad
        if n_text > n_codec:
            return p.r(text[:, :n_codec] + codec), text[:, n_codec:]
        padded = torch.cat([text, pad.expand(1, n_codec - n_text, -1)], 1) if n_codec > n_text else text
        return p.r(padded + codec), pad

    def _rope(self, positions):
        return rope_cos_sin(positions, self.spec.head_dim, self.spec.rope_theta)

    def run_prefill_layers(self, hidden: torch.Tensor, caches: List[KVCache]):
        """talker.rs:823-841: causal mask, offset 0; returns (normed hidden [1,S,H], last logits)."""
        sp = self.spec
        s = hidden.shape[1]
        cos, sin = self._rope(range(s))
        mask = causal_mask(s, 0)
        for lw, c in zip(self.layers, caches):
            hidden = decoder_layer(self.p, lw, hidden, cos, sin, mask, c, sp.heads, sp.kv_heads, sp.head_dim, sp.rms_eps)
        hidden = rms_norm(self.p, hidden, self.norm, sp.rms_eps)
        logits = linear(self.p, hidden[:, s - 1: s], self.codec_head)
        return hidden, logits

    def generate_step_with_embed(self, x: torch.Tensor, caches: List[KVCache], offset: int):
        """talker.rs:716-736: one position, no mask; returns (post-norm hidden, logits)."""
        sp = self.spec
        cos, sin = self._rope([offset])
        h = x
        for lw, c in zip(self.layers, caches):
            h = decoder_layer(self.p, lw, h, cos, sin, None, c, sp.heads, sp.kv_heads, sp.head_dim, sp.rms_eps)
        h = rms_norm(self.p, h, self.norm, sp.rms_eps)
        return h, linear(self.p, h, self.codec_head)


class CodePredictor:
    """CodePredictor (code_predictor.rs:133-519)."""

    def __init__(self, spec: ModelSpec, w: Dict[str, torch.Tensor], prec: Prec):
        self.spec, self.p = spec, prec
        cp = "talker.code_predictor"
        f = lambda n: w[f"{cp}.{n}"].to(torch.float32)
        n_ac = spec.groups - 1
        self.codec_embeddings = [f(f"model.codec_embedding.{g}.weight") for g in range(n_ac)]
        self.proj = (f("small_to_mtp_projection.weight"), f("small_to_mtp_projection.bias")) if spec.has_cp_proj else None
        self.layers = [LayerWeights(w, f"{cp}.model.layers.{l}") for l in range(spec.cp_layers)]
        self.norm = f("model.norm.weight")
        self.lm_heads = [f(f"lm_head.{g}.weight") for g in range(n_ac)]
        # RotaryEmbedding table, 1024 positions (code_predictor.rs:208-213)
        self.cos, self.sin = rope_cos_sin(range(spec.cp_rope_positions), spec.head_dim, spec.rope_theta)

    def new_kv_caches(self):
        return [KVCache(self.spec.cp_max_seq) for _ in range(self.spec.cp_layers)]

    def _project(self, x):
        return linear(self.p, x, *self.proj) if self.proj is not None else x

    def _layers(self, h, caches, offset, mask):
        sp = self.spec
        s = h.shape[1]
        cos, sin = self.cos[offset: offset + s], self.sin[offset: offset + s]
        for lw, c in zip(self.layers, caches):
            h = decoder_layer(self.p, lw, h, cos, sin, mask, c, sp.cp_heads, sp.cp_kv_heads, sp.head_dim, sp.rms_eps)
        return rms_norm(self.p, h, self.norm, sp.rms_eps)

    def generate_acoustic_codes(self, talker_hidden, semantic_embed, caches, return_logits: bool = False,
                                forced_codes: Optional[Sequence[int]] = None):
        """code_predictor.rs:320-416: greedy argmax for each of the 15 codebooks.
        `forced_codes` (test aid, no reference counterpart): the embedding fed to pass g+1 is that of forced_codes[g]
        instead of this pass's own arg-max, so the oracle can FOLLOW a trajectory produced elsewhere (the CUDA path) and
        its per-pass logits stay comparable after the other side took a near-tie the other way.  The returned codes are
        still this oracle's own arg-max of every pass."""
        for c in caches:
            c.reset()
        n_ac = self.spec.groups - 1
        x = self._project(torch.cat([talker_hidden, semantic_embed], 1))
        h = self._layers(x, caches, 0, causal_mask(2, 0))
        logits = linear(self.p, h[:, 1:2], self.lm_heads[0])
        all_logits = [logits]
        codes = [int(torch.argmax(logits.flatten()))]
        offset = 2
        for g in range(1, n_ac):
            prev = codes[-1] if forced_codes is None else int(forced_codes[g - 1])
            e = self.codec_embeddings[g - 1][prev][None, None]
            h = self._layers(self._project(e), caches, offset, None)
            logits = linear(self.p, h, self.lm_heads[g])
            all_logits.append(logits)
            codes.append(int(torch.argmax(logits.flatten())))
            offset += 1
        if return_logits:
            return codes, torch.cat(all_logits, 1)[0]
        return codes

    def embed_codes_for_group(self, group: int, codes: Sequence[int]) -> torch.Tensor:
        """CodePredictor::embed_codes_for_group: rows of codec_embeddings[group] -> [1, T, H_talker]."""
        return self.codec_embeddings[group][torch.tensor(list(codes), dtype=torch.long)][None]

    def acoustic_embeddings_sum(self, codes: Sequence[int]):
        """code_predictor.rs:497-519: acc = E0[c0]; acc += Ei[ci] in order (each add rounded)."""
        acc = self.codec_embeddings[0][codes[0]][None, None]
        for i in range(1, len(codes)):
            acc = self.p.r(acc + self.codec_embeddings[i][codes[i]][None, None])
        return acc


def sum_ref_codec_embeddings(talker: Talker, cp: "CodePredictor", ref_codes: Sequence[Sequence[int]]) -> torch.Tensor:
    """Qwen3TTS::sum_ref_codec_embeddings (lib.rs:1239-1257): per reference frame, talker.codec_embedding[c0] plus the
    code predictor's embedding of group g for c_g, g = 1..15, added in that order (each add rounded) -> [1, T, H]."""
    p = talker.p
    codes = [list(map(int, fr)) for fr in ref_codes]
    summed = talker.codec_embed([fr[0] for fr in codes])
    for g in range(1, 16):
        summed = p.r(summed + cp.embed_codes_for_group(g - 1, [fr[g] for fr in codes]))
    return summed


# lib.rs:1472-1478
ICL_MIN_FRAMES = 75
ICL_FRAMES_PER_TOKEN = 6
ICL_MIN_REPETITION_PENALTY = 1.5


def voice_clone_prompt(talker: Talker, cp: "CodePredictor", text_ids: Sequence[int], speaker_embed: torch.Tensor, language_id: int,
                       ref_codes=None, ref_text_ids=None):
    """The prompt side of synthesize_voice_clone (lib.rs:895-1003): -> (prefill_embeds [1, L, H], trailing [1, Lt, H]).
    The reference prefills 9 (ICL) or 10 positions and, in ICL mode, runs the ICL block as a second causal chunk at offset 9
    (lib.rs:953-987); causal attention makes that identical to one prefill over the concatenation, which is what is
    returned here.  Trailing text: the ICL remainder (or tts_pad alone) in ICL mode, build_trailing_text otherwise."""
    is_icl = ref_codes is not None and ref_text_ids is not None
    hidden = talker.voice_clone_embeds(text_ids, speaker_embed, language_id, is_icl)
    if not is_icl:
        return hidden, talker.build_trailing_text(text_ids)[0]
    ref = sum_ref_codec_embeddings(talker, cp, ref_codes)
    icl, trailing = talker.build_icl_prompt(text_ids, ref_text_ids, ref, non_streaming=False)
    return torch.cat([hidden, icl], 1), trailing
