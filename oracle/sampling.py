"""Oracle restatement of the reference sampler (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows, op for op:
  SamplingContext / PCG-XSH-RR 64/32      src/generation/sampling.rs:18-96
  GenerationConfig                        src/generation/sampling.rs:100-129
  sample                                  src/generation/sampling.rs:140-178
  top_k_filter (GPU tensor path)          src/generation/sampling.rs:203-211
  top_p_filter (GPU tensor path / CPU)    src/generation/sampling.rs:263-286 / 221-262
  multinomial_sample                      src/generation/sampling.rs:290-319
  apply_repetition_penalty(_with_mask)    src/generation/sampling.rs:325-400
  greedy_sample                           src/generation/sampling.rs:403-405
  build_suppression_mask / apply          src/generation/tts.rs:21-68
  apply_generation_penalties_gpu          src/lib.rs:1271-1322
  update_penalty_mask                     src/lib.rs:662-673

All arithmetic is numpy float32 with sequential (index-order) reductions, which is the
order the candle CPU kernels use.  Assumptions about candle internals (not in tree):
  * `Tensor / f64` lowers to affine(mul = 1/T, add = 0) evaluated in the tensor dtype (f32);
  * softmax_last_dim = exp(x - max) / sum, f32;
  * cumsum = inclusive running sum, f32;
  * argmin/argmax return the lowest index among ties.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

F32 = np.float32
MASK64 = (1 << 64) - 1
PCG_MULT = 6364136223846793005
PCG_INC = 1442695040888963407
SEED_MIX = 2685821657736338717


class SamplingContext:
    """sampling.rs:18-96 (seeded mode only; the unseeded time-based LCG is not reproducible)."""

    def __init__(self, seed: int):
        self.reset(seed)

    def reset(self, seed: int):
        self.state = (seed * SEED_MIX + PCG_INC) & MASK64        # sampling.rs:36-38

    def next_u32(self) -> int:
        old = self.state
        self.state = (old * PCG_MULT + PCG_INC) & MASK64          # sampling.rs:86-88
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF      # sampling.rs:90
        rot = old >> 59                                            # sampling.rs:91
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def rand_f32(self) -> np.float32:
        # (output as f32) / (u32::MAX as f32): both conversions round to nearest, so the
        # divisor is exactly 2^32 and the result can be 1.0 (sampling.rs:94).
        return F32(F32(self.next_u32()) / F32(4294967295))


@dataclass
class GenerationConfig:
    """sampling.rs:100-129; defaults here are SynthesisOptions::default (lib.rs:1822-1836)."""
    max_new_tokens: int = 2048
    temperature: float = 0.9
    top_k: int = 50
    top_p: float = 0.9
    repetition_penalty: float = 1.05
    eos_token_id: Optional[int] = 2150
    min_new_tokens: int = 2


def softmax_f32(x: np.ndarray) -> np.ndarray:
    """candle softmax_last_dim on one row: exp(x-max)/sum with a sequential f32 sum."""
    x = x.astype(F32)
    m = x.max()
    e = np.exp((x - m).astype(F32)).astype(F32)
    s = np.cumsum(e, dtype=F32)[-1]   # np.cumsum accumulates sequentially in index order
    return (e / s).astype(F32)


def cumsum_f32(x: np.ndarray) -> np.ndarray:
    return np.cumsum(x.astype(F32), dtype=F32)   # numpy's cumsum is sequential


def greedy_sample(logits: np.ndarray) -> np.ndarray:
    return np.argmax(logits, axis=-1).astype(np.uint32)          # sampling.rs:403-405


def top_k_filter(row: np.ndarray, k: int) -> np.ndarray:
    """sampling.rs:183-211: threshold = k-th largest value, keep `>=` (ties keep extras)."""
    k = min(k, row.shape[0])
    thr = np.sort(row)[::-1][k - 1]
    return np.where(row >= thr, row, F32(-np.inf)).astype(F32)


def top_p_filter(row: np.ndarray, p: float, mode: str = "gpu") -> np.ndarray:
    """mode='gpu': sampling.rs:263-286 (sort desc, softmax, exclusive cumsum >= p removed,
    keep originals >= min kept value).  mode='cpu': sampling.rs:221-262 (keep by index up to
    and including the first inclusive cumsum > p)."""
    sorted_desc = np.sort(row)[::-1].astype(F32)
    probs = softmax_f32(sorted_desc)
    cum = cumsum_f32(probs)
    if mode == "gpu":
        shifted = np.concatenate([np.zeros(1, F32), cum[:-1]])
        remove = shifted >= F32(p)
        kept = np.where(remove, F32(np.inf), sorted_desc)
        min_kept = kept.min()
        return np.where(row >= min_kept, row, F32(-np.inf)).astype(F32)
    order = np.argsort(-row, kind="stable")
    cutoff = row.shape[0]
    c = F32(0.0)
    for i, pr in enumerate(probs):
        c = F32(c + pr)
        if c > F32(p):
            cutoff = i + 1
            break
    out = np.full_like(row, F32(-np.inf))
    out[order[:cutoff]] = row[order[:cutoff]]
    return out


def multinomial_sample(probs: np.ndarray, u: np.float32) -> int:
    """sampling.rs:290-319: first index whose inclusive cumsum >= u, else 0."""
    cum = cumsum_f32(probs)
    hit = np.nonzero(cum >= F32(u))[0]
    return int(hit[0]) if hit.size else 0


def sample_row(logits_row: np.ndarray, cfg: GenerationConfig, u: Optional[np.float32],
               top_p_mode: str = "gpu", return_debug: bool = False):
    """One row of `sample` (sampling.rs:140-178). `u` is the uniform draw for this row."""
    x = logits_row.astype(F32)
    if cfg.temperature != 1.0 and cfg.temperature > 0.0:
        x = (x * F32(1.0 / cfg.temperature)).astype(F32)         # affine(1/T, 0)
    if cfg.temperature < 0.01:
        return int(np.argmax(x))
    if cfg.top_k > 0:
        x = top_k_filter(x, cfg.top_k)
    if 0.0 < cfg.top_p < 1.0:
        x = top_p_filter(x, cfg.top_p, top_p_mode)
    probs = softmax_f32(x)
    tok = multinomial_sample(probs, u)
    if return_debug:
        cum = cumsum_f32(probs)
        # distance from u to the nearest CDF boundary: the exemption measure used when a
        # GPU/CPU exp() ulp difference could flip a sample (tests state the epsilon).
        margin = float(np.min(np.abs(cum[probs > 0] - F32(u)))) if (probs > 0).any() else 0.0
        return tok, dict(probs=probs, margin=margin, kept=int((probs > 0).sum()))
    return tok


def sample(logits: np.ndarray, cfg: GenerationConfig, ctx: SamplingContext,
           top_p_mode: str = "gpu") -> np.ndarray:
    """[batch, vocab] -> [batch] u32; one PCG draw per row, drawn in row order (sampling.rs:297)."""
    logits = np.asarray(logits, dtype=F32)
    if cfg.temperature < 0.01:
        x = logits
        if cfg.temperature != 1.0 and cfg.temperature > 0.0:
            x = (x * F32(1.0 / cfg.temperature)).astype(F32)
        return greedy_sample(x)
    us = [ctx.rand_f32() for _ in range(logits.shape[0])]
    return np.array([sample_row(logits[b], cfg, us[b], top_p_mode) for b in range(logits.shape[0])],
                    dtype=np.uint32)


# -- penalties ---------------------------------------------------------------------------

def build_suppression_mask(vocab_size: int = 3072, eos_token_id: int = 2150) -> np.ndarray:
    """tts.rs:21-43: true for ids in [vocab-1024, vocab) except EOS."""
    m = np.zeros(vocab_size, dtype=bool)
    m[vocab_size - 1024:] = True
    if 0 <= eos_token_id < vocab_size:
        m[eos_token_id] = False
    return m


def apply_token_suppression(logits: np.ndarray, vocab_size: int = 3072, eos_token_id: int = 2150):
    """tts.rs:46-68."""
    mask = build_suppression_mask(vocab_size, eos_token_id)
    return np.where(mask[None, :], F32(-np.inf), logits).astype(F32)


def apply_repetition_penalty_with_mask(logits: np.ndarray, penalty_mask: np.ndarray, penalty: float):
    """sampling.rs:375-400: seen & x>0 -> x*(1/p); seen & x<=0 -> x*p (1/p formed in f32)."""
    if abs(penalty - 1.0) < 1e-9:
        return logits.astype(F32)
    p32 = F32(penalty)
    pos_factor = F32(F32(1.0) / p32)
    factor = np.where(logits > 0, pos_factor, p32).astype(F32)
    final = np.where(penalty_mask > 0, factor, F32(1.0)).astype(F32)
    return (logits * final).astype(F32)


def apply_repetition_penalty(logits: np.ndarray, input_ids, penalty: float):
    """sampling.rs:325-367 (list form; builds the mask then same algebra)."""
    mask = np.zeros(logits.shape[-1], dtype=F32)
    for t in input_ids:
        if t < logits.shape[-1]:
            mask[t] = 1.0
    return apply_repetition_penalty_with_mask(logits, mask[None, :], penalty)


def update_penalty_mask(mask: np.ndarray, token_id: int):
    """lib.rs:662-673 (out-of-range ids are a no-op)."""
    if token_id < mask.shape[-1]:
        mask[..., token_id] = 1.0


def apply_generation_penalties(logits: np.ndarray, penalty_mask: np.ndarray, cfg: GenerationConfig,
                               token_count: int, suppression_mask: np.ndarray) -> np.ndarray:
    """lib.rs:1271-1322: to f32 -> repetition penalty -> suppression -> min_new_tokens EOS."""
    x = np.asarray(logits, dtype=F32)
    if cfg.repetition_penalty != 1.0:
        x = apply_repetition_penalty_with_mask(x, penalty_mask, cfg.repetition_penalty)
    x = np.where(suppression_mask[None, :], F32(-np.inf), x).astype(F32)
    if token_count < cfg.min_new_tokens and cfg.eos_token_id is not None:
        x = x.copy()
        x[:, cfg.eos_token_id] = F32(-np.inf)
    return x
