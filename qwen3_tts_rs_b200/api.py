"""Host-side mirror of the reference's generation / session API over the C ABI.

Names, argument meaning and error behaviour follow the reference so that the parity tests read
like the reference's own tests:

  Qwen3TTS.from_weights            src/lib.rs:267-274
  Qwen3TTS.synthesize_with_voice   src/lib.rs:718-784     (tokenisation is out of scope: takes token ids)
  Qwen3TTS.synthesize_voice_design src/lib.rs:802-870
  Qwen3TTS.generate_codes          src/lib.rs:530-656
  Qwen3TTS.decode_codes            src/lib.rs:881-890
  Qwen3TTS.synthesize_streaming    src/lib.rs:1070-1093 -> StreamingSession (src/lib.rs:1484-1782)
  codes_to_tensor                  src/lib.rs:1417-1431
  SynthesisOptions                 src/lib.rs:1786-1836
  SynthesisTiming                  src/lib.rs:136-147

The reference has no batching; here every call takes a LIST of utterances and row i of the batch is
an independent reference run with utterance i's prompt and seed (disagreement #2).
torch is used for host tensors (bf16 storage) only; all compute happens inside libq3tts_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np
import torch

from . import lib as L
from . import spec as S
from .spec import ModelSpec

CODEC_EOS_TOKEN_ID = S.CODEC_EOS
SAMPLES_PER_FRAME = S.SAMPLES_PER_FRAME


@dataclass
class SynthesisOptions:
    """src/lib.rs:1786-1836 (defaults identical)."""
    max_length: int = 2048
    temperature: float = 0.9
    top_k: int = 50
    top_p: float = 0.9
    repetition_penalty: float = 1.05
    eos_token_id: Optional[int] = CODEC_EOS_TOKEN_ID
    chunk_frames: int = 10
    min_new_tokens: int = 2
    seed: Optional[int] = None
    # Not in the reference (opt-in): 0 = the reference's stateless chunks; -1 = stateful streaming (the
    # session carries the vocoder's cross-chunk state: streamed PCM == non-streamed PCM at a per-chunk cost independent of the
    # utterance length); c > 0 = c frames of left context decoded again in front of every streamed chunk and dropped.
    stream_left_context: int = 0
    # Not in the reference (opt-in): frames of the FIRST streamed chunk (0 = chunk_frames).  With stream_left_context = -1 the
    # waveform does not depend on where chunks are cut, so first chunk 2 + chunk_frames 10 = TTFA of 2-frame chunks at the
    # throughput of 10-frame chunks.
    stream_first_chunk: int = 0

    def to_gen_config(self) -> L.GenConfig:
        g = L.GenConfig()
        g.max_new_tokens = self.max_length
        g.temperature = self.temperature
        g.top_k = self.top_k
        g.top_p = self.top_p
        g.repetition_penalty = self.repetition_penalty
        g.eos_token_id = -1 if self.eos_token_id is None else self.eos_token_id
        g.min_new_tokens = self.min_new_tokens
        g.chunk_frames = self.chunk_frames
        return g


@dataclass
class SynthesisTiming:
    prefill_ms: float
    generation_ms: float
    generation_frames: int
    decode_ms: float


@dataclass
class AudioBuffer:
    """src/audio/io.rs:27-33: mono f32 samples + sample rate."""
    samples: np.ndarray
    sample_rate: int = 24000

    def __len__(self):
        return int(self.samples.shape[-1])

    def duration(self) -> float:
        return len(self) / self.sample_rate

    def is_empty(self) -> bool:
        return len(self) == 0

    def save(self, path: str) -> None:
        """AudioBuffer::save (io.rs:71-73): PCM16 mono WAV."""
        from . import formats
        formats.save_wav(path, self.samples, self.sample_rate)

    @classmethod
    def load(cls, path: str) -> "AudioBuffer":
        """AudioBuffer::load (io.rs:76-78)."""
        from . import formats
        samples, rate = formats.load_wav(path)
        return cls(samples, rate)

    def normalize(self) -> None:
        from . import formats
        self.samples = formats.normalize(self.samples)

    def normalize_db(self, target_db: float) -> None:
        from . import formats
        self.samples = formats.normalize_db(self.samples, target_db)


@dataclass
class VoiceClonePrompt:
    """VoiceClonePrompt (lib.rs:127-134).  The encoders that produce it (ECAPA-TDNN speaker encoder, Mimi speech encoder) are
    out of scope here: the speaker embedding and the reference codes enter as data."""
    speaker_embedding: torch.Tensor                     # [hidden]
    ref_codes: Optional[np.ndarray] = None              # [T_ref, 16] codec codes of the reference audio (ICL mode)
    ref_text_ids: Optional[Sequence[int]] = None        # tokenized reference transcript (ICL mode)

    @property
    def is_icl(self) -> bool:
        return self.ref_codes is not None and self.ref_text_ids is not None


# lib.rs:1472-1478
ICL_MIN_FRAMES = 75
ICL_FRAMES_PER_TOKEN = 6
ICL_MIN_REPETITION_PENALTY = 1.5
POS_SPEAKER = -2                                        # Q3_POS_SPEAKER


def pos_ref_frame(t: int) -> int:                       # Q3_POS_REF_FRAME(t)
    return -16 - t


def codes_to_tensor(codes: Sequence[Sequence[int]]) -> np.ndarray:
    """src/lib.rs:1417-1431: [n_frames][16] u32 -> i64 [1,16,T], data[q*T + f] = codes[f][q]."""
    n = len(codes)
    out = np.zeros((1, 16, n), dtype=np.int64)
    if n:
        out[0] = np.asarray(codes, dtype=np.int64).T
    return out


def _ptr(a) -> C.c_void_p:
    if isinstance(a, torch.Tensor):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def _bf16_bits(t: torch.Tensor) -> torch.Tensor:
    """contiguous bf16 host tensor (rounds f32 inputs to nearest-even, like candle's to_dtype)."""
    return t.detach().to(device="cpu", dtype=torch.bfloat16).contiguous()


class Model:
    """Owns a q3_model handle (weights resident on one GPU)."""

    def __init__(self, spec: ModelSpec, device: int = 0):
        self.spec, self.device = spec, device
        self.lib = L.load()
        self.handle = C.c_void_p()
        desc = L.model_desc(spec, device)
        L.check(self.lib.q3_model_create(C.byref(desc), C.byref(self.handle)))
        self.finalized = False

    def set_tensor(self, name: str, t: torch.Tensor):
        if name.startswith(("decoder.", "speaker_encoder.")):       # the vocoder and the speaker encoder are F32 (lib.rs:344-345)
            h = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
            dt = L.Q3_F32
        else:
            h = _bf16_bits(t)
            dt = L.Q3_BF16
        shape = (C.c_int64 * h.dim())(*h.shape)
        L.check(self.lib.q3_model_set_tensor(self.handle, name.encode(), _ptr(h), dt, shape, h.dim(), 0))

    def load(self, weights: Dict[str, torch.Tensor]):
        for k, v in weights.items():
            self.set_tensor(k, v)
        return self

    def finalize(self):
        L.check(self.lib.q3_model_finalize(self.handle))
        self.finalized = True
        return self

    def close(self):
        if self.handle:
            self.lib.q3_model_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Session:
    """Owns a q3_session (KV caches, RNG, penalty masks, frame graph) for `batch` utterances."""

    def __init__(self, model: Model, batch: int, options: SynthesisOptions, seeds: Sequence[int],
                 max_seq: Optional[int] = None):
        assert len(seeds) == batch
        self.model, self.B, self.options = model, batch, options
        self.lib = model.lib
        self.max_seq = max_seq if max_seq is not None else options.max_length + 256   # lib.rs:756
        self.cfg = options.to_gen_config()
        self.handle = C.c_void_p()
        sd = (C.c_uint64 * batch)(*[int(s) & ((1 << 64) - 1) for s in seeds])
        L.check(self.lib.q3_session_create(model.handle, batch, self.max_seq, C.byref(self.cfg), sd, C.byref(self.handle)))
        if options.stream_left_context:
            L.check(self.lib.q3_session_set_stream_context(self.handle, int(options.stream_left_context)))
        if options.stream_first_chunk:
            L.check(self.lib.q3_session_set_first_chunk(self.handle, int(options.stream_first_chunk)))

    def reset(self, seeds: Sequence[int]):
        sd = (C.c_uint64 * self.B)(*[int(s) & ((1 << 64) - 1) for s in seeds])
        L.check(self.lib.q3_session_reset(self.handle, sd))

    # -- prompt ------------------------------------------------------------------------------
    def prefill_embeds(self, embeds: Sequence[torch.Tensor]):
        """embeds[b]: [L_b, hidden] (any float dtype; stored bf16)."""
        H = self.model.spec.hidden
        lens = np.array([e.shape[0] for e in embeds], dtype=np.int32)
        lmax = int(lens.max())
        buf = torch.zeros(self.B, lmax, H, dtype=torch.bfloat16)
        for b, e in enumerate(embeds):
            buf[b, : e.shape[0]] = e.to(torch.bfloat16)
        L.check(self.lib.q3_prefill_embeds(self.handle, _ptr(buf), _ptr(lens), lmax))

    def prefill_ids(self, text_ids: Sequence[Sequence[int]], codec_ids: Sequence[Sequence[int]]):
        lens = np.array([len(t) for t in text_ids], dtype=np.int32)
        lmax = int(lens.max())
        ti = np.full((self.B, lmax), -1, dtype=np.int32)
        ci = np.full((self.B, lmax), -1, dtype=np.int32)
        for b in range(self.B):
            ti[b, : lens[b]] = text_ids[b]
            ci[b, : lens[b]] = codec_ids[b]
        L.check(self.lib.q3_prefill_ids(self.handle, _ptr(ti), _ptr(ci), _ptr(lens), lmax))

    def prefill_voice_clone(self, text_ids: Sequence[Sequence[int]], codec_ids: Sequence[Sequence[int]],
                            speaker_embeds: Sequence[torch.Tensor], ref_codes: Sequence[Optional[np.ndarray]]):
        """q3_prefill_voice_clone: as prefill_ids with Q3_POS_SPEAKER / Q3_POS_REF_FRAME(t) codec parts; speaker_embeds[b]
        [hidden] (stored bf16), ref_codes[b] u32 [T_ref, 16] or None."""
        H = self.model.spec.hidden
        lens = np.array([len(t) for t in text_ids], dtype=np.int32)
        lmax = int(lens.max())
        ti = np.full((self.B, lmax), -1, dtype=np.int32)
        ci = np.full((self.B, lmax), -1, dtype=np.int32)
        for b in range(self.B):
            ti[b, : lens[b]] = text_ids[b]
            ci[b, : lens[b]] = codec_ids[b]
        spk = torch.zeros(self.B, H, dtype=torch.bfloat16)
        for b, e in enumerate(speaker_embeds):
            spk[b] = e.reshape(H).to(torch.bfloat16)
        t_ref = np.array([0 if r is None else len(r) for r in ref_codes], dtype=np.int32)
        tmax = int(t_ref.max())
        rc = np.zeros((self.B, max(1, tmax), 16), dtype=np.uint32)
        for b, r in enumerate(ref_codes):
            if r is not None and len(r):
                rc[b, : len(r)] = np.asarray(r, dtype=np.uint32)
        L.check(self.lib.q3_prefill_voice_clone(self.handle, _ptr(ti), _ptr(ci), _ptr(lens), lmax, _ptr(spk), _ptr(rc), _ptr(t_ref), tmax))

    def set_trailing_text(self, trailing: Sequence[torch.Tensor], tts_pad: torch.Tensor):
        H = self.model.spec.hidden
        lt = np.array([t.shape[0] for t in trailing], dtype=np.int32)
        lmax = max(1, int(lt.max()))
        buf = torch.zeros(self.B, lmax, H, dtype=torch.bfloat16)
        for b, t in enumerate(trailing):
            buf[b, : t.shape[0]] = t.to(torch.bfloat16)
        pad = _bf16_bits(tts_pad.reshape(-1))
        L.check(self.lib.q3_set_trailing_text(self.handle, _ptr(buf), _ptr(lt), lmax, _ptr(pad)))

    def set_trailing_ids(self, ids: Sequence[Sequence[int]]):
        """Rows = text_proj(ids[b]) ++ tts_eos; pad = text_proj(tts_pad)  (lib.rs:508-519)."""
        sp = self.model.spec
        # ids[b] is None: no trailing rows at all (every frame adds tts_pad) -- an ICL prompt that consumed the text
        n = np.array([-1 if t is None else len(t) for t in ids], dtype=np.int32)
        nmax = max(1, int(n.max()))
        buf = np.zeros((self.B, nmax), dtype=np.int32)
        for b, t in enumerate(ids):
            if t is not None:
                buf[b, : len(t)] = t
        L.check(self.lib.q3_set_trailing_ids(self.handle, _ptr(buf), _ptr(n), nmax,
                                             S.special_text_id(sp, S.TTS_EOS), S.special_text_id(sp, S.TTS_PAD)))

    # -- generation ---------------------------------------------------------------------------
    def generate(self, max_frames: int):
        codes = np.zeros((self.B, max_frames, 16), dtype=np.uint32)
        n = np.zeros(self.B, dtype=np.int32)
        L.check(self.lib.q3_generate(self.handle, max_frames, _ptr(codes), _ptr(n)))
        return codes, n

    def trailing_rows(self, cap: int = 256):
        """q3_debug_get_trailing (test aid): (trailing bf16 [B, cap, H], lt [B], tts_pad bf16 [H]) as built on the device."""
        H = self.model.spec.hidden
        tr = torch.zeros(self.B, cap, H, dtype=torch.bfloat16)
        lt = np.zeros(self.B, dtype=np.int32)
        pad = torch.zeros(H, dtype=torch.bfloat16)
        fn = self.lib.q3_debug_get_trailing
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
        L.check(fn(self.handle, _ptr(tr), cap, _ptr(lt), _ptr(pad)))
        return tr, lt, pad

    def decode_generation(self) -> int:
        """0 = multi-kernel CUDA graph, 1..4 = generation of the persistent frame kernel this session runs on (debug aid)."""
        fn = self.lib.q3_debug_decode_generation
        fn.argtypes = [C.c_void_p]
        fn.restype = C.c_int
        return int(fn(self.handle))

    def generate_tapped(self, max_frames: int):
        """q3_debug_generate_tapped (test aid): q3_generate with the loop's decision inputs read back every frame.
        Returns (codes, n_frames, taps) with taps = dict(first_logits [B,V], logits [F,B,V], cp_logits [F,15,B,cpV],
        rng [F+2,B] (PCG state before the first draw, then after every draw), step_input bf16 [F,B,H])."""
        sp = self.model.spec
        B, F, n_ac = self.B, max_frames, sp.groups - 1
        codes = np.zeros((B, F, 16), dtype=np.uint32)
        n = np.zeros(B, dtype=np.int32)
        taps = dict(first_logits=np.zeros((B, sp.codec_vocab), dtype=np.float32),
                    logits=np.zeros((F, B, sp.codec_vocab), dtype=np.float32),
                    cp_logits=np.zeros((F, n_ac, B, sp.cp_vocab), dtype=np.float32),
                    rng=np.zeros((F + 2, B), dtype=np.uint64),
                    step_input=torch.zeros(F, B, sp.hidden, dtype=torch.bfloat16))
        fn = self.lib.q3_debug_generate_tapped
        fn.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 7
        fn.restype = C.c_int
        L.check(fn(self.handle, F, _ptr(codes), _ptr(n), _ptr(taps["first_logits"]), _ptr(taps["logits"]),
                   _ptr(taps["cp_logits"]), _ptr(taps["rng"]), _ptr(taps["step_input"])))
        return codes, n, taps

    def generate_async(self, max_frames: int):
        L.check(self.lib.q3_generate_async(self.handle, max_frames))

    def get_codes(self, max_frames: int):
        codes = np.zeros((self.B, max_frames, 16), dtype=np.uint32)
        n = np.zeros(self.B, dtype=np.int32)
        L.check(self.lib.q3_get_codes(self.handle, max_frames, _ptr(codes), _ptr(n)))
        return codes, n

    def vocode(self, max_frames: int, to_host: bool = True):
        up = self.model.spec.vocoder.total_upsample
        if not to_host:
            L.check(self.lib.q3_vocode_session(self.handle, max_frames, None))
            return None
        pcm = np.zeros((self.B, max_frames * up), dtype=np.float32)
        L.check(self.lib.q3_vocode_session(self.handle, max_frames, _ptr(pcm)))
        return pcm

    def stream_next(self):
        chunk = max(1, self.options.chunk_frames)
        up = self.model.spec.vocoder.total_upsample
        codes = np.zeros((self.B, chunk, 16), dtype=np.uint32)
        pcm = np.zeros((self.B, chunk * up), dtype=np.float32)
        n = np.zeros(self.B, dtype=np.int32)
        done = C.c_int32(0)
        L.check(self.lib.q3_stream_next(self.handle, _ptr(codes), _ptr(pcm), _ptr(n), C.byref(done)))
        return codes, pcm, n, bool(done.value)

    # -- per-op entry points ------------------------------------------------------------------------
    def talker_step(self, step_input: torch.Tensor):
        sp = self.model.spec
        x = _bf16_bits(step_input.reshape(self.B, sp.hidden))
        hid = torch.zeros(self.B, sp.hidden, dtype=torch.bfloat16)
        logits = np.zeros((self.B, sp.codec_vocab), dtype=np.float32)
        L.check(self.lib.q3_talker_step(self.handle, _ptr(x), _ptr(hid), _ptr(logits)))
        return hid, logits

    def code_predictor_frame(self, last_hidden: torch.Tensor, sem_tokens: Sequence[int], want_logits: bool = False):
        sp = self.model.spec
        h = _bf16_bits(last_hidden.reshape(self.B, sp.hidden))
        tk = np.asarray(sem_tokens, dtype=np.uint32)
        codes = np.zeros((self.B, sp.groups - 1), dtype=np.uint32)
        lg = np.zeros((self.B, sp.groups - 1, sp.cp_vocab), dtype=np.float32) if want_logits else None
        L.check(self.lib.q3_code_predictor_frame(self.handle, _ptr(h), _ptr(tk), _ptr(codes),
                                                 _ptr(lg) if want_logits else None))
        return (codes, lg) if want_logits else codes

    def timing(self) -> SynthesisTiming:
        t = L.Timing()
        L.check(self.lib.q3_session_timing(self.handle, C.byref(t)))
        return SynthesisTiming(t.prefill_ms, t.generation_ms, t.generation_frames, t.decode_ms)

    def synchronize(self):
        L.check(self.lib.q3_session_synchronize(self.handle))

    def close(self):
        if self.handle:
            self.lib.q3_session_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sample(logits: np.ndarray, cfg: SynthesisOptions, rng_states: np.ndarray, seen_mask: np.ndarray,
           token_count: int, model: Optional[Model] = None) -> np.ndarray:
    """Per-op sampler entry point (q3_sample): penalties + sample + mask update, in place on
    rng_states (uint64 [B]) and seen_mask (uint8 [B,V])."""
    lib = L.load()
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    B, V = lg.shape
    assert rng_states.dtype == np.uint64 and seen_mask.dtype == np.uint8
    g = cfg.to_gen_config()
    out = np.zeros(B, dtype=np.uint32)
    L.check(lib.q3_sample(model.handle if model else None, _ptr(lg), B, V, C.byref(g), _ptr(rng_states), _ptr(seen_mask),
                          token_count, _ptr(out)))
    return out


def fused_residual_rmsnorm(x: torch.Tensor, r: torch.Tensor, w: torch.Tensor, eps: float, device: int = 0):
    """FusedRmsNorm::forward_residual through the C ABI with host buffers -> (normed, sum)."""
    lib = L.load()
    assert x.dtype == r.dtype == w.dtype and x.dtype in (torch.bfloat16, torch.float32)
    cols = x.shape[-1]
    rows = x.numel() // cols
    xc, rc, wc = x.contiguous(), r.contiguous(), w.contiguous()
    normed, total = torch.empty_like(xc), torch.empty_like(xc)
    dt = L.Q3_BF16 if x.dtype == torch.bfloat16 else L.Q3_F32
    L.check(lib.q3_fused_residual_rmsnorm_host(_ptr(xc), _ptr(rc), _ptr(wc), _ptr(normed), _ptr(total), rows, cols,
                                               float(eps), dt, device))
    return normed, total


class StreamingSession:
    """StreamingSession (src/lib.rs:1484-1782) for one utterance."""

    def __init__(self, tts: "Qwen3TTS", sess: Session):
        self.tts, self.sess = tts, sess
        self._frames = 0
        self._done = False

    def next_chunk(self) -> Optional[AudioBuffer]:
        if self._done:
            return None
        codes, pcm, n, done = self.sess.stream_next()
        self._frames += int(n[0])
        self._done = done
        if n[0] == 0:
            return None
        return AudioBuffer(pcm[0, : int(n[0]) * SAMPLES_PER_FRAME].copy(), 24000)

    def frames_generated(self) -> int:
        return self._frames

    def is_done(self) -> bool:
        return self._done

    def __iter__(self) -> Iterator[AudioBuffer]:
        while True:
            c = self.next_chunk()
            if c is None:
                return
            yield c


class Qwen3TTS:
    """Facade mirroring `Qwen3TTS` (src/lib.rs:150-1322) for the decode hot path."""

    def __init__(self, model: Model):
        self.model = model
        self.spec = model.spec
        self.model_type: Optional[str] = None

    @classmethod
    def from_weights(cls, spec: ModelSpec, talker_weights: Dict[str, torch.Tensor],
                     vocoder_weights: Optional[Dict[str, torch.Tensor]] = None, device: int = 0) -> "Qwen3TTS":
        m = Model(spec, device)
        m.load(talker_weights)
        if vocoder_weights:
            m.load(vocoder_weights)
        m.finalize()
        return cls(m)

    @classmethod
    def from_pretrained(cls, model_id: str, device: int = 0) -> "Qwen3TTS":
        """Qwen3TTS::from_pretrained (lib.rs:183-262) for the decode path: a local checkpoint directory with
        config.json (optional), model.safetensors and speech_tokenizer/model.safetensors.  Tokenizer loading and hub
        download are outside the hot path (callers pass token ids).  The variant detected from config.json is kept
        in `model_type` (None when the dimensions came from weight inspection, lib.rs:383-386)."""
        from . import formats
        ck = formats.load_checkpoint(model_id)
        voc = dict(ck.vocoder_weights)
        voc.update(ck.speaker_weights or {})          # F32 parts: vocoder + (Base checkpoints) the ECAPA speaker encoder
        tts = cls.from_weights(ck.spec, ck.talker_weights, voc, device)
        tts.model_type = ck.config.model_type if ck.config is not None else None
        return tts

    def supports_voice_cloning(self) -> bool:
        """lib.rs:389-391: a speaker encoder was loaded."""
        return int(self.model.lib.q3_speaker_embed_dim(self.model.handle)) > 0

    def supports_preset_speakers(self) -> bool:
        """lib.rs:393-404: CustomVoice only; permissive when the variant is unknown."""
        return self.model_type in (None, "custom_voice")

    def supports_voice_design(self) -> bool:
        """lib.rs:406-411: VoiceDesign only (an unknown variant is NOT accepted here, unlike the preset speakers)."""
        return self.model_type == "voice_design"

    # -- prompt assembly (host logic only: id lists; embedding math runs on the device) -----------------
    def custom_voice_prompt(self, text_ids: Sequence[int], speaker: str, language: str):
        """Position-wise (text id, codec id) pairs of prefill_custom_voice (talker.rs:451-488); -1 = absent."""
        sp = self.spec
        sid = lambda t: S.special_text_id(sp, t)
        text = [sid(S.IM_START), sid(S.ASSISTANT), sid(S.NEWLINE)] + [sid(S.TTS_PAD)] * 5 + [sid(S.TTS_BOS)]
        codec = [-1, -1, -1, S.CODEC_THINK, S.CODEC_THINK_BOS, S.LANGUAGE_IDS[language], S.CODEC_THINK_EOS,
                 S.SPEAKER_IDS[speaker], S.CODEC_PAD]
        if len(text_ids) > 0:
            text.append(int(text_ids[0]))
            codec.append(S.CODEC_BOS)
        return text, codec

    def voice_design_prompt(self, text_ids: Sequence[int], instruct_ids: Sequence[int], language: str):
        """prefill_voice_design (talker.rs:585-624)."""
        sp = self.spec
        sid = lambda t: S.special_text_id(sp, t)
        text = [int(t) for t in instruct_ids] + [sid(S.IM_START), sid(S.ASSISTANT), sid(S.NEWLINE)] + \
               [sid(S.TTS_PAD)] * 4 + [sid(S.TTS_BOS)]
        codec = [-1] * (len(instruct_ids) + 3) + [S.CODEC_THINK, S.CODEC_THINK_BOS, S.LANGUAGE_IDS[language],
                                                  S.CODEC_THINK_EOS, S.CODEC_PAD]
        if len(text_ids) > 0:
            text.append(int(text_ids[0]))
            codec.append(S.CODEC_BOS)
        return text, codec

    def voice_clone_prompt(self, text_ids: Sequence[int], prompt: VoiceClonePrompt, language: str):
        """Position-wise (text id, codec part) pairs of prefill_voice_clone (talker.rs:511-564) followed, in ICL mode, by the
        streaming overlay of build_icl_prompt (talker.rs:646-705; the reference runs it as a second causal chunk, lib.rs:953-987
        -- one causal prefill over the concatenation is the same computation).  -> (text, codec, trailing ids or None):
        trailing ids are what set_trailing_ids takes (it appends tts_eos); None = no trailing rows, every frame adds tts_pad."""
        sp = self.spec
        sid = lambda t: S.special_text_id(sp, t)
        text = [sid(S.IM_START), sid(S.ASSISTANT), sid(S.NEWLINE)] + [sid(S.TTS_PAD)] * 5 + [sid(S.TTS_BOS)]
        codec = [-1, -1, -1, S.CODEC_THINK, S.CODEC_THINK_BOS, S.LANGUAGE_IDS[language], S.CODEC_THINK_EOS, POS_SPEAKER, S.CODEC_PAD]
        text_ids = [int(t) for t in text_ids]
        if not prompt.is_icl:
            if len(text_ids) > 0:
                text.append(text_ids[0])
                codec.append(S.CODEC_BOS)
            return text, codec, text_ids[1:]
        all_text = [int(t) for t in prompt.ref_text_ids] + text_ids + [sid(S.TTS_EOS)]
        n_text, n_codec = len(all_text), len(prompt.ref_codes) + 1
        for i in range(n_codec):
            text.append(all_text[i] if i < n_text else sid(S.TTS_PAD))
            codec.append(S.CODEC_BOS if i == 0 else pos_ref_frame(i - 1))
        trailing = all_text[n_codec:-1] if n_text > n_codec else None      # the remainder ends in tts_eos, which set_trailing_ids appends
        return text, codec, trailing

    def generate_codes_voice_clone(self, batch_text_ids, prompts: Sequence[VoiceClonePrompt], language: str = "english",
                                   options: Optional[SynthesisOptions] = None, seeds=None):
        """The code-generation half of synthesize_voice_clone_debug (lib.rs:895-1017) for a batch: ICL adjustments of the
        generation config (repetition penalty >= 1.5, frame budget min(max_length, max(75, 6 x text tokens)), lib.rs:913-927),
        voice-clone prefill (+ ICL block), trailing text, frame loop.  -> list of FrameCodes."""
        import dataclasses
        options = options or SynthesisOptions()
        icl = [p.is_icl for p in prompts]
        if any(icl) != all(icl):
            raise ValueError("a batch must be all-ICL or all x-vector-only (the repetition penalty differs)")
        budgets = [min(options.max_length, max(ICL_MIN_FRAMES, len(t) * ICL_FRAMES_PER_TOKEN)) if p.is_icl else options.max_length
                   for t, p in zip(batch_text_ids, prompts)]
        if all(icl):
            options = dataclasses.replace(options, repetition_penalty=max(options.repetition_penalty, ICL_MIN_REPETITION_PENALTY))
        pp = [self.voice_clone_prompt(t, p, language) for t, p in zip(batch_text_ids, prompts)]
        mf = max(budgets)
        B = len(prompts)
        sess = Session(self.model, B, dataclasses.replace(options, max_length=mf), self._seeds(options, B, seeds),
                       max_seq=max(len(p[0]) for p in pp) + mf + 8)
        try:
            sess.prefill_voice_clone([p[0] for p in pp], [p[1] for p in pp], [p.speaker_embedding for p in prompts],
                                     [p.ref_codes if p.is_icl else None for p in prompts])
            sess.set_trailing_ids([p[2] for p in pp])
            codes, n = sess.generate(mf)
            # rows are independent: a row whose own budget is below the batch's is cut there, as its own run would have been
            return [codes[b, : min(int(n[b]), budgets[b])].tolist() for b in range(B)]
        finally:
            sess.close()

    def synthesize_voice_clone(self, batch_text_ids, prompts: Sequence[VoiceClonePrompt], language: str = "english",
                               options: Optional[SynthesisOptions] = None, seeds=None):
        """synthesize_voice_clone (lib.rs:895-1060): in ICL mode the reference frames are decoded in front of the generated
        ones and ref_len / total_len of the waveform is cut from its start (lib.rs:1021-1040)."""
        all_codes = self.generate_codes_voice_clone(batch_text_ids, prompts, language, options, seeds)
        out = []
        for codes, p in zip(all_codes, prompts):
            if p.is_icl:
                ref = [[int(c) for c in fr] for fr in np.asarray(p.ref_codes)]
                combined = ref + codes
                audio = self.decode_codes(combined)
                cut = len(ref) * len(audio) // max(1, len(combined))
                audio.samples = audio.samples[min(cut, len(audio)):]
                out.append(audio)
            else:
                out.append(self.decode_codes(codes))
        return out

    def _new_session(self, batch_text_ids, prompts, options: SynthesisOptions, seeds, max_seq=None) -> Session:
        sess = Session(self.model, len(batch_text_ids), options, seeds, max_seq)
        sess.prefill_ids([p[0] for p in prompts], [p[1] for p in prompts])
        sess.set_trailing_ids([list(t[1:]) for t in batch_text_ids])
        return sess

    def _seeds(self, options: SynthesisOptions, n: int, seeds):
        if seeds is not None:
            return list(seeds)
        if options.seed is None:
            raise ValueError("a seed is required (the reference's unseeded mode is time-based and not reproducible)")
        return [options.seed + i for i in range(n)]

    def generate_codes(self, batch_text_ids: Sequence[Sequence[int]], speaker: str = "ryan", language: str = "english",
                       options: Optional[SynthesisOptions] = None, seeds: Optional[Sequence[int]] = None,
                       max_frames: Optional[int] = None):
        """generate_codes for a batch of CustomVoice utterances -> list of FrameCodes ([n_frames][16])."""
        options = options or SynthesisOptions()
        prompts = [self.custom_voice_prompt(t, speaker, language) for t in batch_text_ids]
        sess = self._new_session(batch_text_ids, prompts, options, self._seeds(options, len(prompts), seeds))
        try:
            mf = max_frames if max_frames is not None else options.max_length
            codes, n = sess.generate(mf)
            return [codes[b, : n[b]].tolist() for b in range(len(prompts))]
        finally:
            sess.close()

    def synthesize_with_voice(self, batch_text_ids, speaker="ryan", language="english",
                              options: Optional[SynthesisOptions] = None, seeds=None, max_frames=None,
                              with_timing: bool = False):
        """synthesize_with_voice / synthesize_with_timing: prefill -> generate_codes -> decode_codes."""
        options = options or SynthesisOptions()
        prompts = [self.custom_voice_prompt(t, speaker, language) for t in batch_text_ids]
        return self._synthesize(batch_text_ids, prompts, options, seeds, max_frames, with_timing)

    def synthesize_voice_design(self, batch_text_ids, batch_instruct_ids, language="english",
                                options: Optional[SynthesisOptions] = None, seeds=None, max_frames=None,
                                with_timing: bool = False):
        options = options or SynthesisOptions()
        prompts = [self.voice_design_prompt(t, i, language) for t, i in zip(batch_text_ids, batch_instruct_ids)]
        return self._synthesize(batch_text_ids, prompts, options, seeds, max_frames, with_timing)

    def _synthesize(self, batch_text_ids, prompts, options, seeds, max_frames, with_timing):
        lmax = max(len(p[0]) for p in prompts)
        mf = max_frames if max_frames is not None else options.max_length
        sess = self._new_session(batch_text_ids, prompts, options, self._seeds(options, len(prompts), seeds),
                                 max_seq=max(options.max_length + 256, lmax + mf))
        try:
            codes, n = sess.generate(mf)
            pcm = sess.vocode(mf)
            audio = [AudioBuffer(pcm[b, : n[b] * SAMPLES_PER_FRAME].copy()) for b in range(len(prompts))]
            if with_timing:
                return audio, sess.timing()
            return audio
        finally:
            sess.close()

    def synthesize_streaming(self, text_ids: Sequence[int], speaker="ryan", language="english",
                             options: Optional[SynthesisOptions] = None) -> StreamingSession:
        options = options or SynthesisOptions()
        prompts = [self.custom_voice_prompt(text_ids, speaker, language)]
        sess = self._new_session([text_ids], prompts, options, self._seeds(options, 1, None))
        return StreamingSession(self, sess)

    def synthesize_voice_design_streaming(self, text_ids, instruct_ids, language="english",
                                          options: Optional[SynthesisOptions] = None) -> StreamingSession:
        options = options or SynthesisOptions()
        prompts = [self.voice_design_prompt(text_ids, instruct_ids, language)]
        sess = self._new_session([text_ids], prompts, options, self._seeds(options, 1, None))
        return StreamingSession(self, sess)

    def speaker_encode(self, mel: np.ndarray) -> torch.Tensor:
        """SpeakerEncoder::forward (speaker.rs:448-476): mel spectrogram f32 [B, mel_dim, T] -> raw speaker embedding
        f32 [B, enc_dim] (the VoiceClonePrompt.speaker_embedding).  Needs the speaker_encoder.* weights in the model; the
        mel front end (src/audio/mel.rs) is the caller's."""
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        B, _, T = mel.shape
        dim = int(self.model.lib.q3_speaker_embed_dim(self.model.handle))
        if dim <= 0:
            raise L.Q3Error(5, "model has no finalized speaker-encoder weights")
        out = np.zeros((B, dim), dtype=np.float32)
        L.check(self.model.lib.q3_speaker_encode(self.model.handle, _ptr(mel), B, T, _ptr(out)))
        return torch.from_numpy(out)

    def codes_to_tensor(self, codes):
        return codes_to_tensor(codes)

    def decode_codes(self, codes: Sequence[Sequence[int]]) -> AudioBuffer:
        """decode_codes (lib.rs:881-890): [n_frames][16] -> 24 kHz audio."""
        t = codes_to_tensor(codes)
        return AudioBuffer(self.decode_tensor(t)[0])

    def decode_tensor(self, codes_i64: np.ndarray) -> np.ndarray:
        """Decoder12Hz::decode: i64 [B,16,T] -> f32 [B, T*1920]."""
        codes_i64 = np.ascontiguousarray(codes_i64, dtype=np.int64)
        B, nq, T = codes_i64.shape
        up = self.spec.vocoder.total_upsample
        pcm = np.zeros((B, T * up), dtype=np.float32)
        L.check(self.model.lib.q3_vocoder_decode(self.model.handle, _ptr(codes_i64), B, T, _ptr(pcm)))
        return pcm
