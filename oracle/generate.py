"""Oracle restatement of the generation loop (TEST INFRASTRUCTURE ONLY).

Follows:
  Qwen3TTS::generate_codes             src/lib.rs:530-656
  StreamingSession::{from_prefill,next_chunk}  src/lib.rs:1584-1759
  codes_to_tensor                      src/lib.rs:1417-1431
  decode_codes                         src/lib.rs:881-890
  synthesize_with_voice / _voice_design  src/lib.rs:718-784, 802-870

Batch semantics: the reference has no batching (SURVEY.md "where the north star and the
reference disagree" #2); a batch of B utterances is B independent runs of this loop.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import sampling as smp
from .model import CodePredictor, Talker, Prec


def codes_to_tensor(codes: Sequence[Sequence[int]]) -> np.ndarray:
    """lib.rs:1417-1431: [n_frames][16] -> i64 [1,16,T], data[q*T + f]."""
    n = len(codes)
    out = np.zeros((1, 16, n), dtype=np.int64)
    for f, frame in enumerate(codes):
        for q, c in enumerate(frame):
            out[0, q, f] = c
    return out


class Trace:
    """Per-frame intermediates for teacher-forced comparison with the CUDA path."""

    def __init__(self):
        self.frames: List[dict] = []
        self.first: Optional[dict] = None      # penalised prefill logits + RNG state of the first draw (lib.rs:557-571)


def sample_first(talker: Talker, logits: torch.Tensor, cfg: smp.GenerationConfig, ctx: smp.SamplingContext,
                 penalty_mask: np.ndarray, suppression: np.ndarray, logit_hook=None, trace: Optional["Trace"] = None):
    """lib.rs:557-571: sample token 0 with token_count = 0."""
    l2 = logits[:, 0].numpy().astype(np.float32)
    if logit_hook is not None:
        l2 = logit_hook(-1, l2)
    l2 = smp.apply_generation_penalties(l2, penalty_mask, cfg, 0, suppression)
    if trace is not None:
        trace.first = dict(penalised=l2.copy(), rng_state=ctx.state)
    tok = int(smp.sample(l2, cfg, ctx)[0])
    smp.update_penalty_mask(penalty_mask, tok)
    return tok


def generate_codes(talker: Talker, cp: CodePredictor, cfg: smp.GenerationConfig, ctx: smp.SamplingContext,
                   kv_caches, offset: int, last_hidden: torch.Tensor, initial_logits: torch.Tensor,
                   trailing_text_hidden: torch.Tensor, trailing_text_len: int, tts_pad_embed: torch.Tensor,
                   trace: Optional[Trace] = None,
                   logit_hook: Optional[Callable[[int, np.ndarray], np.ndarray]] = None) -> List[List[int]]:
    """lib.rs:530-656.  `logit_hook(frame_idx, logits_f32[1,V])` lets a test force EOS at a chosen
    frame by editing the raw talker logits before penalties."""
    p: Prec = talker.p
    vocab = talker.spec.codec_vocab
    suppression = smp.build_suppression_mask(vocab, 2150)
    penalty_mask = np.zeros((1, vocab), dtype=np.float32)
    cp_caches = cp.new_kv_caches()

    tok = sample_first(talker, initial_logits, cfg, ctx, penalty_mask, suppression, logit_hook, trace)
    token_count = 1
    frames: List[List[int]] = []
    for frame_idx in range(cfg.max_new_tokens):
        if cfg.eos_token_id is not None and tok == cfg.eos_token_id:      # lib.rs:581-585
            break
        sem = talker.codec_embed([tok])                                   # lib.rs:588-590
        if trace is not None:
            codes, cp_logits = cp.generate_acoustic_codes(last_hidden, sem, cp_caches, return_logits=True)
        else:
            codes = cp.generate_acoustic_codes(last_hidden, sem, cp_caches)
        frames.append([tok] + codes)                                      # lib.rs:605-609
        summed = p.r(sem + cp.acoustic_embeddings_sum(codes))             # lib.rs:612-615
        if frame_idx < trailing_text_len:                                 # lib.rs:617-621
            text_add = trailing_text_hidden[:, frame_idx: frame_idx + 1]
        else:
            text_add = tts_pad_embed
        step_input = p.r(summed + text_add)
        h, logits = talker.generate_step_with_embed(step_input, kv_caches, offset)   # lib.rs:627-631
        offset += 1
        raw = logits[:, 0].numpy().astype(np.float32)
        if logit_hook is not None:
            raw = logit_hook(frame_idx, raw)
        l2 = smp.apply_generation_penalties(raw, penalty_mask, cfg, token_count, suppression)
        rng_state_before = ctx.state
        next_tok = int(smp.sample(l2, cfg, ctx)[0])
        sample_margin = None
        if trace is not None and cfg.temperature >= 0.01:
            # distance of this draw from the nearest CDF boundary (exemption measure for bf16 logit noise)
            probe = smp.SamplingContext(0)
            probe.state = rng_state_before
            _, dbg = smp.sample_row(l2[0], cfg, probe.rand_f32(), return_debug=True)
            sample_margin = dbg["margin"]
        if trace is not None:
            trace.frames.append(dict(frame=frame_idx, tok=tok, codes=codes, cp_in_hidden=last_hidden.clone(),
                                     cp_logits=cp_logits, step_input=step_input.clone(), hidden=h.clone(),
                                     logits=raw.copy(), penalised=l2.copy(), rng_state=rng_state_before,
                                     next_tok=next_tok, sample_margin=sample_margin))
        last_hidden = h
        tok = next_tok
        smp.update_penalty_mask(penalty_mask, tok)
        token_count += 1
    return frames


def prefill_and_generate(talker: Talker, cp: CodePredictor, prefill_embeds: torch.Tensor,
                         text_ids: Sequence[int], cfg: smp.GenerationConfig, seed: int,
                         trace: Optional[Trace] = None, logit_hook=None, kv_max: Optional[int] = None):
    """synthesize_with_voice minus tokenizer and vocoder (lib.rs:743-777)."""
    ctx = smp.SamplingContext(seed)
    trailing, tlen, pad = talker.build_trailing_text(text_ids)
    caches = talker.new_kv_caches(kv_max if kv_max is not None else cfg.max_new_tokens + 256)
    hidden, logits = talker.run_prefill_layers(prefill_embeds, caches)
    plen = hidden.shape[1]
    last_hidden = hidden[:, plen - 1: plen]
    return generate_codes(talker, cp, cfg, ctx, caches, plen, last_hidden, logits, trailing, tlen, pad,
                          trace=trace, logit_hook=logit_hook)


def follow(talker: Talker, cp: CodePredictor, prefill_embeds: torch.Tensor, text_ids: Sequence[int],
           cfg: smp.GenerationConfig, seed: int, frames: Sequence[Sequence[int]],
           first_logits: Optional[np.ndarray] = None, frame_logits: Optional[Sequence[np.ndarray]] = None,
           kv_max: Optional[int] = None, text_rows=None, trailing_override: Optional[torch.Tensor] = None) -> dict:
    """Test aid (no reference counterpart): the loop of generate_codes (lib.rs:530-656) FOLLOWING a trajectory produced
    elsewhere.  `frames[f] = [tok, c0..c14]` are the codes the CUDA path emitted; every decision input (semantic token,
    acoustic codes fed to the next code-predictor pass, penalty mask) is taken from them, every tensor is this oracle's
    own.  Returns per frame the oracle's code-predictor logits, its own arg-max codes, the talker input built from the
    followed codes (lib.rs:612-622), the talker hidden state and raw logits -- so the other side can be compared
    element-wise at EVERY frame of a free-running run, not only up to its first near-tie fork.

    Sampler replay: when `first_logits` ([V] f32, the prefill logits the other side sampled token 0 from) and
    `frame_logits[f]` ([V] f32, the talker logits it sampled frame f+1's token from) are given, the oracle's penalty /
    top-k / top-p / multinomial pipeline (sampling.rs:140-319, lib.rs:1271-1322) is run on THOSE logits with this
    oracle's RNG stream and penalty mask; `replayed[f]` is the token it draws (index 0 = the first token), `margins[f]`
    the draw's distance from the nearest CDF boundary and `rng_states[f]` the PCG state before the draw."""
    p: Prec = talker.p
    vocab = talker.spec.codec_vocab
    suppression = smp.build_suppression_mask(vocab, 2150)
    penalty_mask = np.zeros((1, vocab), dtype=np.float32)
    ctx = smp.SamplingContext(seed)
    trailing, tlen, pad = talker.build_trailing_text(text_ids)
    if trailing_override is not None:
        # voice-clone ICL mode (lib.rs:953-987): the trailing text is what build_icl_prompt left over, [1, Lt, H]
        trailing, tlen = trailing_override.to(torch.float32), int(trailing_override.shape[1])
    out_text = dict(trailing=trailing.clone(), pad=pad.clone())
    if text_rows is not None:
        # (trailing [1, tlen, H], pad [1, 1, H]) as the OTHER side projected them: the talker-input add (lib.rs:617-621) is then
        # checked bit-exactly on its own, without inheriting the rounding noise of the text-projection GEMM
        tr_o, pad_o = text_rows
        assert tr_o.shape[1] == tlen, (tr_o.shape, tlen)
        trailing, pad = tr_o.to(torch.float32), pad_o.to(torch.float32)
    caches = talker.new_kv_caches(kv_max if kv_max is not None else cfg.max_new_tokens + 256)
    hidden, logits = talker.run_prefill_layers(prefill_embeds, caches)
    offset = hidden.shape[1]
    last_hidden = hidden[:, offset - 1: offset]
    cp_caches = cp.new_kv_caches()
    out = dict(text=out_text, prefill_logits=logits[:, 0].numpy().astype(np.float32)[0], frames=[], replayed=[], margins=[], rng_states=[])

    def replay(raw, token_count):
        l2 = smp.apply_generation_penalties(np.asarray(raw, dtype=np.float32)[None], penalty_mask, cfg, token_count, suppression)
        out["rng_states"].append(ctx.state)
        probe = smp.SamplingContext(0)
        probe.state = ctx.state
        tok = int(smp.sample(l2, cfg, ctx)[0])
        margin = None
        if cfg.temperature >= 0.01:
            _, dbg = smp.sample_row(l2[0], cfg, probe.rand_f32(), return_debug=True)
            margin = dbg["margin"]
        out["replayed"].append(tok)
        out["margins"].append(margin)

    if first_logits is not None or frame_logits is not None:
        replay(first_logits if first_logits is not None else out["prefill_logits"], 0)    # one draw per sampled token
    if len(frames):
        smp.update_penalty_mask(penalty_mask, int(frames[0][0]))
    token_count = 1
    for f, fr in enumerate(frames):
        tok, forced = int(fr[0]), [int(c) for c in fr[1:]]
        sem = talker.codec_embed([tok])
        own_codes, cp_logits = cp.generate_acoustic_codes(last_hidden, sem, cp_caches, return_logits=True, forced_codes=forced)
        summed = p.r(sem + cp.acoustic_embeddings_sum(forced))                 # lib.rs:612-615
        text_add = trailing[:, f: f + 1] if f < tlen else pad                  # lib.rs:617-621
        step_input = p.r(summed + text_add)
        h, lg = talker.generate_step_with_embed(step_input, caches, offset)
        offset += 1
        out["frames"].append(dict(frame=f, cp_logits=cp_logits, own_codes=own_codes, step_input=step_input.clone(),
                                  hidden=h.clone(), logits=lg[:, 0].numpy().astype(np.float32)[0]))
        if frame_logits is not None:
            replay(frame_logits[f], token_count)
        nxt = int(frames[f + 1][0]) if f + 1 < len(frames) else (out["replayed"][-1] if frame_logits is not None else None)
        if nxt is not None:
            smp.update_penalty_mask(penalty_mask, nxt)
        token_count += 1
        last_hidden = h
    return out


class StreamingSession:
    """StreamingSession (lib.rs:1484-1782): same per-frame body; every `chunk_frames` frames the
    buffered codes are decoded INDEPENDENTLY (no vocoder state crosses chunks, lib.rs:1755-1758)."""

    def __init__(self, talker: Talker, cp: CodePredictor, decode_fn, prefill_embeds: torch.Tensor,
                 text_ids: Sequence[int], cfg: smp.GenerationConfig, seed: int, chunk_frames: int = 10,
                 left_context: int = 0):
        self.talker, self.cp, self.decode_fn, self.cfg = talker, cp, decode_fn, cfg
        # left_context != 0 is the product's opt-in extension (q3_session_set_stream_context), restated here only so the
        # tests can check it; 0 is the reference.  -1 = the whole history.
        self.left_context = left_context
        self.ctx = smp.SamplingContext(seed)
        self.trailing, self.tlen, self.pad = talker.build_trailing_text(text_ids)
        self.caches = talker.new_kv_caches(cfg.max_new_tokens + 256)
        hidden, logits = talker.run_prefill_layers(prefill_embeds, self.caches)
        plen = hidden.shape[1]
        self.offset = plen
        self.last_hidden = hidden[:, plen - 1: plen]
        vocab = talker.spec.codec_vocab
        self.suppression = smp.build_suppression_mask(vocab, 2150)
        self.penalty_mask = np.zeros((1, vocab), dtype=np.float32)
        first = sample_first(talker, logits, cfg, self.ctx, self.penalty_mask, self.suppression)
        self.done = cfg.eos_token_id == first                          # lib.rs:1621
        self.current_token = None if self.done else first
        self.frames_generated = 0
        self.frame_buffer: List[List[int]] = []
        self.chunk_frames = chunk_frames
        self.token_count = 1
        self.cp_caches = cp.new_kv_caches()
        self.all_frames: List[List[int]] = []

    def next_chunk(self):
        """lib.rs:1650-1759. Returns PCM ndarray or None."""
        p = self.talker.p
        if self.done:
            if self.frame_buffer:
                buf, self.frame_buffer = self.frame_buffer, []
                return self._decode_chunk(buf)
            return None
        while len(self.frame_buffer) < self.chunk_frames and self.frames_generated < self.cfg.max_new_tokens:
            if self.current_token is None:
                self.done = True
                break
            tok = self.current_token
            sem = self.talker.codec_embed([tok])
            codes = self.cp.generate_acoustic_codes(self.last_hidden, sem, self.cp_caches)
            self.frame_buffer.append([tok] + codes)
            self.all_frames.append([tok] + codes)
            frame_idx = self.frames_generated
            self.frames_generated += 1
            summed = p.r(sem + self.cp.acoustic_embeddings_sum(codes))
            text_add = self.trailing[:, frame_idx: frame_idx + 1] if frame_idx < self.tlen else self.pad
            step_input = p.r(summed + text_add)
            h, logits = self.talker.generate_step_with_embed(step_input, self.caches, self.offset)
            self.offset += 1
            self.last_hidden = h
            l2 = smp.apply_generation_penalties(logits[:, 0].numpy().astype(np.float32), self.penalty_mask,
                                                self.cfg, self.token_count, self.suppression)
            nxt = int(smp.sample(l2, self.cfg, self.ctx)[0])
            smp.update_penalty_mask(self.penalty_mask, nxt)
            self.token_count += 1
            if self.cfg.eos_token_id == nxt:                            # lib.rs:1740-1747
                self.current_token = None
                self.done = True
            else:
                self.current_token = nxt
        if not self.frame_buffer:
            return None
        buf, self.frame_buffer = self.frame_buffer, []
        return self._decode_chunk(buf)

    def _decode_chunk(self, buf):
        """lib.rs:1755-1758 decodes `buf` alone.  With left context c > 0 the c frames before it are decoded again in
        front and their samples dropped (all vocoder ops are causal)."""
        f0 = len(self.all_frames) - len(buf)
        c = f0 if self.left_context < 0 else min(self.left_context, f0)
        if c == 0:
            return self.decode_fn(codes_to_tensor(buf))
        pcm = self.decode_fn(codes_to_tensor(self.all_frames[f0 - c:]))
        return pcm[c * 1920:]

    def is_done(self):
        return self.done and not self.frame_buffer

    def __iter__(self):
        while True:
            c = self.next_chunk()
            if c is None:
                return
            yield c
