"""Model dimension tables for the Qwen3-TTS decode hot path.

Mirrors the dimension sources in the reference:
  TalkerConfig::default / custom_voice   (src/models/talker.rs:208-274)
  CodePredictorConfig::default / custom_voice (src/models/code_predictor.rs:48-113)
  Decoder12HzConfig::default             (src/models/codec/decoder_12hz.rs:47-67)
  codec / tts special token ids          (src/models/talker.rs:31-54, 96-105, 147-156)

Only dimensions live here; `formats.ParsedModelConfig` turns a checkpoint's config.json
into one of these tables.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Tuple

# -- token id tables (src/models/talker.rs:31-54) ---------------------------
IM_START = 151644
IM_END = 151645
ASSISTANT = 77091
NEWLINE = 198
TTS_PAD = 151671
TTS_BOS = 151672
TTS_EOS = 151673
CODEC_PAD = 2148
CODEC_BOS = 2149
CODEC_EOS = 2150
CODEC_THINK = 2154
CODEC_NOTHINK = 2155
CODEC_THINK_BOS = 2156
CODEC_THINK_EOS = 2157
CODEC_VOCAB_SIZE = 3072
SAMPLES_PER_FRAME = 1920  # src/lib.rs:1469

LANGUAGE_IDS = {  # src/models/talker.rs:96-105
    "chinese": 2055, "english": 2050, "japanese": 2058, "korean": 2064,
    "german": 2053, "french": 2061, "russian": 2069, "portuguese": 2071,
    "spanish": 2054, "italian": 2070,
}
SPEAKER_IDS = {  # src/models/talker.rs:147-156
    "serena": 3066, "vivian": 3065, "uncle_fu": 3010, "ryan": 3061,
    "aiden": 2861, "ono_anna": 2873, "sohee": 2864, "eric": 2875, "dylan": 2878,
}


@dataclass(frozen=True)
class VocoderSpec:
    """Decoder12HzConfig (decoder_12hz.rs:47-67)."""
    codebook_dim: int = 512        # output of the two 1x1 projections
    vq_dim: int = 256              # width of one codebook row
    latent_dim: int = 1024
    hidden_size: int = 512
    num_layers: int = 8
    num_heads: int = 16
    head_dim: int = 64
    intermediate_size: int = 1024
    num_quantizers: int = 16
    codebook_size: int = 2048
    upsampling_ratios: Tuple[int, ...] = (2, 2)
    decoder_dim: int = 1536
    upsample_rates: Tuple[int, ...] = (8, 5, 4, 3)
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    layer_scale: float = 0.01

    @property
    def total_upsample(self) -> int:
        n = 1
        for r in self.upsampling_ratios + self.upsample_rates:
            n *= r
        return n


@dataclass(frozen=True)
class SpeakerSpec:
    """SpeakerEncoderConfig (config.rs:102-175): ECAPA-TDNN on a 128-band mel spectrogram."""
    mel_dim: int = 128
    enc_dim: int = 1024
    enc_channels: Tuple[int, ...] = (512, 512, 512, 512, 1536)     # initial, 3 x SE-Res2Net, MFA
    enc_kernel_sizes: Tuple[int, ...] = (5, 3, 3, 3, 1)
    enc_dilations: Tuple[int, ...] = (1, 2, 3, 4, 1)
    enc_attention_channels: int = 128
    enc_res2net_scale: int = 8
    enc_se_channels: int = 128


TINY_SPEAKER = SpeakerSpec(mel_dim=32, enc_dim=256, enc_channels=(64, 64, 64, 64, 192), enc_attention_channels=32,
                           enc_res2net_scale=4, enc_se_channels=32)


@dataclass(frozen=True)
class ModelSpec:
    name: str
    # talker (talker.rs:208-274)
    hidden: int
    inter: int
    layers: int = 28
    heads: int = 16
    kv_heads: int = 8
    head_dim: int = 128
    codec_vocab: int = CODEC_VOCAB_SIZE
    text_vocab: int = 151936
    text_embed_dim: int = 2048
    rope_theta: float = 1000000.0
    rms_eps: float = 1e-6
    # code predictor (code_predictor.rs:48-64)
    cp_hidden: int = 1024
    cp_inter: int = 3072
    cp_layers: int = 5
    cp_heads: int = 16
    cp_kv_heads: int = 8
    cp_vocab: int = 2048
    groups: int = 16
    cp_rope_positions: int = 1024   # code_predictor.rs:208-213
    cp_max_seq: int = 17            # code_predictor.rs:282-284
    vocoder: VocoderSpec = field(default_factory=VocoderSpec)

    @property
    def has_cp_proj(self) -> bool:
        """small_to_mtp_projection exists iff talker hidden != CP hidden (code_predictor.rs:175-183)."""
        return self.hidden != self.cp_hidden

    @property
    def q_dim(self) -> int:
        return self.heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.kv_heads * self.head_dim

    def to_dict(self):
        return asdict(self)


SPEC_0_6B = ModelSpec(name="0.6b", hidden=1024, inter=3072)
SPEC_1_7B = ModelSpec(name="1.7b", hidden=2048, inter=6144)

# Scaled-down variants for CPU-speed tests.  Same structure, same codec vocab
# (the suppression rule is defined on 3072 ids), head_dim stays 128 because the
# CUDA kernels are specialised for it.
TINY_VOCODER = VocoderSpec(codebook_dim=64, vq_dim=32, latent_dim=96, hidden_size=64,
                           num_layers=2, num_heads=4, head_dim=16, intermediate_size=128,
                           codebook_size=2048, decoder_dim=128)
SPEC_TINY = ModelSpec(name="tiny", hidden=256, inter=512, layers=3, heads=4, kv_heads=2,
                      text_vocab=2048, text_embed_dim=256,
                      cp_hidden=256, cp_inter=512, cp_layers=2, cp_heads=4, cp_kv_heads=2,
                      vocoder=TINY_VOCODER)
# tiny with a small_to_mtp projection (talker hidden != cp hidden), like the 1.7B
SPEC_TINY_PROJ = ModelSpec(name="tiny_proj", hidden=512, inter=512, layers=2, heads=4, kv_heads=2,
                           text_vocab=2048, text_embed_dim=256,
                           cp_hidden=256, cp_inter=512, cp_layers=2, cp_heads=4, cp_kv_heads=2,
                           vocoder=TINY_VOCODER)

SPECS = {s.name: s for s in (SPEC_0_6B, SPEC_1_7B, SPEC_TINY, SPEC_TINY_PROJ)}


def talker_weight_bytes(spec: ModelSpec) -> int:
    """bf16 bytes streamed by one talker step (SURVEY.md §8d W_talker)."""
    h, i = spec.hidden, spec.inter
    layer = h * spec.q_dim + 2 * h * spec.kv_dim + spec.q_dim * h + 3 * h * i + 2 * h + 2 * spec.head_dim
    return 2 * (spec.layers * layer + h + spec.codec_vocab * h)


def cp_weight_bytes_per_frame(spec: ModelSpec) -> int:
    """bf16 bytes streamed by the 15 code-predictor passes of one frame."""
    h, i = spec.cp_hidden, spec.cp_inter
    qd, kd = spec.cp_heads * spec.head_dim, spec.cp_kv_heads * spec.head_dim
    layer = h * qd + 2 * h * kd + qd * h + 3 * h * i + 2 * h + 2 * spec.head_dim
    layers = spec.cp_layers * layer + h
    proj = (spec.hidden * h + h) if spec.has_cp_proj else 0
    n_ac = spec.groups - 1
    return 2 * (n_ac * (layers + proj) + n_ac * spec.cp_vocab * h)


def kv_bytes_per_position(spec: ModelSpec) -> int:
    return spec.layers * 2 * spec.kv_heads * spec.head_dim * 2


def step_bytes(spec: ModelSpec, batch: int, ctx_len: float) -> float:
    """Algorithmic HBM bytes of one decode step for `batch` rows (SURVEY.md §8d formula)."""
    return (talker_weight_bytes(spec) + cp_weight_bytes_per_frame(spec)
            + batch * 30 * spec.hidden * 2 + batch * ctx_len * kv_bytes_per_position(spec))


def special_text_id(spec: ModelSpec, tok: int) -> int:
    """Special text-token ids (>= 151643) sit at a fixed distance from the end of the text
    vocab; scaled-down test specs keep that distance.  Real vocab: identity."""
    if spec.text_vocab == 151936:
        return tok
    return tok - 151936 + spec.text_vocab if tok >= 151643 else tok % (spec.text_vocab - 300)


# mid-size spec: hidden >= 1024 exercises the 128-thread reference-order RMSNorm path and the
# small_to_mtp projection, still cheap enough for the CPU oracle.
SPEC_MID = ModelSpec(name="mid", hidden=2048, inter=1024, layers=2, heads=4, kv_heads=2,
                     text_vocab=2048, text_embed_dim=256,
                     cp_hidden=1024, cp_inter=1024, cp_layers=2, cp_heads=4, cp_kv_heads=2,
                     vocoder=TINY_VOCODER)
SPECS["mid"] = SPEC_MID

# "ring": the 1.7B's matrix shapes (hidden 2048, 16/8 heads, inter 6144; code predictor 1024 / 3072; small_to_mtp projection)
# with 2 + 2 layers: every skinny-GEMM K is a multiple of 1024, so the TMA-ring kernel (mega4.cuh, Q3_MEGA=4) takes it --
# K = 3072 and 6144 down projections, 48-row gate/up tiles, 16-token pass 0 -- and the CPU oracle still finishes in seconds.
SPEC_RING = ModelSpec(name="ring", hidden=2048, inter=6144, layers=2, heads=16, kv_heads=8,
                      text_vocab=2048, text_embed_dim=256,
                      cp_hidden=1024, cp_inter=3072, cp_layers=2, cp_heads=16, cp_kv_heads=8,
                      vocoder=TINY_VOCODER)
SPECS["ring"] = SPEC_RING
