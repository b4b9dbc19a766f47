"""On-disk / wire formats either side of the decode hot path.

These are the formats the reference reads and writes around `generate_codes` / `Decoder12Hz::decode`,
restated so that somebody holding the real checkpoint and a Rust toolchain can exchange files with this
build (`generate_audio --compare`, reference_validation.rs golden vectors, HF checkpoints):

  save_codes_binary / load_codes_binary   src/bin/generate_audio.rs:788-801, 826-832   i64 LE, frame-major
  save_audio_binary / load_audio_binary   src/bin/generate_audio.rs:803-813, 878-884   f32 LE
  load_reference                          tests/reference_validation.rs:15-23          raw f32 LE + caller's shape
  compare_with_reference                  src/bin/generate_audio.rs:816-920            codes equality + audio diff stats
  save_wav / load_wav                     src/audio/io.rs:110-165                      PCM16 mono, `(clamp(x)*32767) as i16`
  normalize / normalize_db                src/audio/io.rs:83-103
  load_safetensors / save_safetensors     candle_core::safetensors::load (src/lib.rs:1390-1396); format = safetensors 0.4
  ParsedModelConfig                       src/models/config.rs:205-353                 config.json keys and defaults
  detect_spec_from_weights                src/lib.rs:370-381                           fallback when config.json is absent
  hot_path_tensor_names / load_checkpoint src/lib.rs:183-262, 305-366                  which files and tensors the decode path needs
  export_checkpoint                       (inverse of the above; gives the reference the synthetic model)

Host-side byte work only: nothing here touches the GPU, and nothing here imports `oracle/`.
"""
from __future__ import annotations

import json
import math
import os
import struct
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import spec as S
from .spec import ModelSpec, VocoderSpec
from . import weights as W

# ------------------------------------------------------------------------------------------------------------
# codes / audio dumps (generate_audio.rs)
# ------------------------------------------------------------------------------------------------------------


def save_codes_binary(codes: Sequence[Sequence[int]], path: str) -> None:
    """[n_frames][16] u32 -> i64 little-endian, frame0_q0, frame0_q1, ..., frame1_q0, ... (generate_audio.rs:788-801)."""
    flat = np.asarray([c for frame in codes for c in frame], dtype="<i8")
    with open(path, "wb") as f:
        f.write(flat.tobytes())


def load_codes_binary(path: str, groups: int = 16) -> List[List[int]]:
    """Inverse of save_codes_binary.  A trailing partial i64 is dropped the way `chunks(8)` + `try_into().unwrap()`
    would panic on it: here it is an error; a length that is not a whole number of frames is an error too."""
    raw = open(path, "rb").read()
    if len(raw) % 8:
        raise ValueError(f"{path}: {len(raw)} bytes is not a whole number of i64 values")
    flat = np.frombuffer(raw, dtype="<i8")
    if flat.size % groups:
        raise ValueError(f"{path}: {flat.size} values is not a whole number of {groups}-code frames")
    return flat.reshape(-1, groups).tolist()


def save_audio_binary(samples: np.ndarray, path: str) -> None:
    """f32 little-endian samples (generate_audio.rs:803-813)."""
    with open(path, "wb") as f:
        f.write(np.ascontiguousarray(samples, dtype="<f4").tobytes())


def load_audio_binary(path: str) -> np.ndarray:
    raw = open(path, "rb").read()
    return np.frombuffer(raw[: len(raw) // 4 * 4], dtype="<f4").astype(np.float32)   # chunks_exact(4) semantics


def load_reference(path: str, shape: Sequence[int]) -> np.ndarray:
    """Golden-vector file of tests/reference_validation.rs:15-23: raw f32 LE, shape supplied by the test."""
    a = load_audio_binary(path)
    n = int(np.prod(shape)) if len(shape) else 1
    if a.size != n:
        raise ValueError(f"{path}: {a.size} f32 values, shape {tuple(shape)} needs {n}")
    return a.reshape(tuple(shape))


@dataclass
class CompareReport:
    """What `compare_with_reference` prints (generate_audio.rs:816-920), as values."""
    codes_found: bool = False
    codes_match: bool = False
    n_ref_codes: int = 0
    n_our_codes: int = 0
    n_code_diffs: int = 0
    first_code_diffs: Tuple[Tuple[int, int, int], ...] = ()   # (flat index, reference, ours), first five
    audio_found: bool = False
    n_audio_compared: int = 0
    max_diff: float = 0.0
    mean_diff: float = 0.0
    rmse: float = 0.0


def reference_dump_paths(reference_dir: str, seed: int, num_frames: int) -> Tuple[str, str]:
    """File names of the Python exporter's dumps (generate_audio.rs:826, 876)."""
    return (os.path.join(reference_dir, f"codes_seed{seed}_frames{num_frames}.bin"),
            os.path.join(reference_dir, f"audio_seed{seed}_frames{num_frames}.bin"))


def compare_with_reference(reference_dir: str, seed: int, num_frames: int,
                           codes: Sequence[Sequence[int]], audio: np.ndarray) -> CompareReport:
    """Codes must be identical value for value (and in count); audio is compared over the common prefix with
    max / mean absolute difference and RMSE accumulated in f64 (generate_audio.rs:886-905)."""
    rep = CompareReport()
    cpath, apath = reference_dump_paths(reference_dir, seed, num_frames)
    if os.path.exists(cpath):
        rep.codes_found = True
        ref = np.frombuffer(open(cpath, "rb").read(), dtype="<i8")
        ours = np.asarray([c for frame in codes for c in frame], dtype=np.int64)
        rep.n_ref_codes, rep.n_our_codes = int(ref.size), int(ours.size)
        m = min(ref.size, ours.size)
        diff = np.nonzero(ref[:m] != ours[:m])[0]
        rep.n_code_diffs = int(diff.size)
        rep.first_code_diffs = tuple((int(i), int(ref[i]), int(ours[i])) for i in diff[:5])
        rep.codes_match = ref.size == ours.size and diff.size == 0
    if os.path.exists(apath):
        rep.audio_found = True
        ref = load_audio_binary(apath)
        ours = np.asarray(audio, dtype=np.float32).reshape(-1)
        m = min(ref.size, ours.size)
        rep.n_audio_compared = int(m)
        if m:
            d = np.abs(ref[:m] - ours[:m])                   # f32 subtraction, like the reference
            rep.max_diff = float(d.max())
            rep.mean_diff = float(d.astype(np.float64).sum() / m)
            rep.rmse = float(np.sqrt((d * d).astype(np.float64).sum() / m))
    return rep


# ------------------------------------------------------------------------------------------------------------
# WAV (src/audio/io.rs; the reference uses the `hound` crate)
# ------------------------------------------------------------------------------------------------------------


def pcm_f32_to_i16(samples: np.ndarray) -> np.ndarray:
    """`(sample.clamp(-1.0, 1.0) * 32767.0) as i16` (io.rs:155-160): f32 multiply, truncation toward zero.
    NaN clamps to NaN in Rust and `NaN as i16` is 0."""
    x = np.asarray(samples, dtype=np.float32)
    y = np.clip(x, np.float32(-1.0), np.float32(1.0)) * np.float32(32767.0)
    y = np.where(np.isnan(y), np.float32(0.0), y)
    return np.trunc(y).astype("<i2")


def save_wav(path: str, samples: np.ndarray, sample_rate: int = 24000) -> None:
    """PCM16, mono, canonical 44-byte RIFF header (what hound's WavWriter emits for this WavSpec)."""
    data = pcm_f32_to_i16(np.asarray(samples).reshape(-1)).tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE"
    hdr += b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, sample_rate, sample_rate * 2, 2, 16)
    hdr += b"data" + struct.pack("<I", len(data))
    with open(path, "wb") as f:
        f.write(hdr + data)


def load_wav(path: str) -> Tuple[np.ndarray, int]:
    """-> (mono f32 samples, sample_rate).  Integer PCM is scaled by 1/2^(bits-1), float is taken as is, and
    multi-channel files are averaged to mono (io.rs:110-141).  8-bit PCM is unsigned in the file and centred
    on 128, as hound presents it."""
    raw = open(path, "rb").read()
    if len(raw) < 12 or raw[:4] != b"RIFF" or raw[8:12] != b"WAVE":
        raise ValueError(f"Failed to open WAV file: {path}: not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(raw):
        cid, size = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        body = raw[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = body
        elif cid == b"data":
            data = body
            break
        pos += 8 + size + (size & 1)
    if fmt is None or data is None or len(fmt) < 16:
        raise ValueError(f"Failed to open WAV file: {path}: missing fmt or data chunk")
    tag, channels, rate, _, _, bits = struct.unpack("<HHIIHH", fmt[:16])
    if tag == 0xFFFE and len(fmt) >= 26:                      # WAVE_FORMAT_EXTENSIBLE: sub-format GUID's first two bytes
        tag = struct.unpack("<H", fmt[24:26])[0]
    if tag == 3 and bits == 32:
        x = np.frombuffer(data[: len(data) // 4 * 4], dtype="<f4").astype(np.float32)
    elif tag == 1 and bits in (8, 16, 24, 32):
        if bits == 8:
            v = np.frombuffer(data, dtype=np.uint8).astype(np.int32) - 128
        elif bits == 16:
            v = np.frombuffer(data[: len(data) // 2 * 2], dtype="<i2").astype(np.int32)
        elif bits == 24:
            b = np.frombuffer(data[: len(data) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v >= 1 << 23, v - (1 << 24), v)
        else:
            v = np.frombuffer(data[: len(data) // 4 * 4], dtype="<i4").astype(np.int64)
        x = v.astype(np.float32) / np.float32(1 << (bits - 1))
    else:
        raise ValueError(f"{path}: unsupported WAV format tag {tag} with {bits} bits per sample")
    if channels > 1:
        x = x[: x.size // channels * channels].reshape(-1, channels)
        x = (x.sum(axis=1, dtype=np.float32) / np.float32(channels)).astype(np.float32)
    return x, int(rate)


def normalize(samples: np.ndarray) -> np.ndarray:
    """AudioBuffer::normalize (io.rs:83-91): divide by the peak unless it is 0 or already 1."""
    x = np.asarray(samples, dtype=np.float32)
    m = np.float32(np.abs(x).max()) if x.size else np.float32(0)
    return x / m if (m > 0 and m != 1) else x.copy()


def normalize_db(samples: np.ndarray, target_db: float) -> np.ndarray:
    """AudioBuffer::normalize_db (io.rs:94-103): peak to 10^(dB/20)."""
    x = np.asarray(samples, dtype=np.float32)
    m = np.float32(np.abs(x).max()) if x.size else np.float32(0)
    if not m > 0:
        return x.copy()
    return x * (np.float32(10.0) ** np.float32(target_db / 20.0) / m)


# ------------------------------------------------------------------------------------------------------------
# safetensors (8-byte LE header length, JSON header {name: {dtype, shape, data_offsets}}, raw LE tensor bytes)
# ------------------------------------------------------------------------------------------------------------

_ST_DTYPES = {
    "F64": torch.float64, "F32": torch.float32, "F16": torch.float16, "BF16": torch.bfloat16,
    "I64": torch.int64, "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8, "BOOL": torch.bool,
}
_ST_NAMES = {v: k for k, v in _ST_DTYPES.items()}


def read_safetensors_header(path: str) -> Tuple[Dict[str, dict], int]:
    """-> (header dict without __metadata__, byte offset of the data section).  Validates what the format
    requires: header length in range, offsets inside the file, contiguous and matching dtype*shape."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        head = f.read(8)
        if len(head) < 8:
            raise ValueError(f"{path}: too short for a safetensors header")
        n = struct.unpack("<Q", head)[0]
        if n > size - 8 or n > 100_000_000:
            raise ValueError(f"{path}: header length {n} is out of range")
        hdr = json.loads(f.read(n).decode("utf-8"))
    if not isinstance(hdr, dict):
        raise ValueError(f"{path}: header is not a JSON object")
    hdr.pop("__metadata__", None)
    base = 8 + n
    for name, e in hdr.items():
        if not isinstance(e, dict) or not {"dtype", "shape", "data_offsets"} <= set(e):
            raise ValueError(f"{path}: tensor {name} lacks dtype / shape / data_offsets")
        dt = _ST_DTYPES.get(e["dtype"])
        if dt is None:
            raise ValueError(f"{path}: tensor {name} has unsupported dtype {e['dtype']}")
        b, end = e["data_offsets"]
        if not all(isinstance(d, int) and not isinstance(d, bool) and d >= 0 for d in e["shape"]):
            raise ValueError(f"{path}: tensor {name} has a bad shape {e['shape']}")
        numel = math.prod(e["shape"])                       # Python ints: a crafted shape cannot wrap around
        want = numel * torch.empty(0, dtype=dt).element_size()
        if not (0 <= b <= end <= size - base) or end - b != want:
            raise ValueError(f"{path}: tensor {name} has bad data_offsets {e['data_offsets']} for {e['dtype']}{e['shape']}")
    return hdr, base


def load_safetensors(path: str, names: Optional[Sequence[str]] = None) -> Dict[str, torch.Tensor]:
    """candle_core::safetensors::load: every tensor, in its stored dtype.  `names` restricts the read to the tensors
    the caller needs (the hot path uses 0.9-3.9 GB of a checkpoint that also holds the speaker / speech encoders).
    The file is memory-mapped; each returned tensor owns a private copy."""
    hdr, base = read_safetensors_header(path)
    want = hdr.keys() if names is None else [n for n in names if n in hdr]
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    out = {}
    for name in want:
        e = hdr[name]
        b, end = e["data_offsets"]
        dt = _ST_DTYPES[e["dtype"]]
        buf = np.array(mm[base + b: base + end])             # copy out of the mapping
        if buf.size == 0:
            out[name] = torch.empty(e["shape"], dtype=dt)
        else:
            out[name] = torch.frombuffer(buf, dtype=dt).reshape(e["shape"])
    del mm
    return out


def save_safetensors(tensors: Dict[str, torch.Tensor], path: str, metadata: Optional[Dict[str, str]] = None) -> None:
    """Writer used to export the synthetic checkpoints so that the reference can load the very same weights
    (names sorted, header padded with spaces to an 8-byte boundary, as the safetensors library writes them)."""
    hdr, off, blobs = {}, 0, []
    if metadata:
        hdr["__metadata__"] = dict(metadata)
    for name in sorted(tensors):
        t = tensors[name].detach().cpu().contiguous()
        if t.dtype not in _ST_NAMES:
            raise ValueError(f"{name}: dtype {t.dtype} has no safetensors name")
        raw = t.reshape(-1).view(torch.uint8).numpy().tobytes() if t.numel() else b""
        hdr[name] = {"dtype": _ST_NAMES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        off += len(raw)
        blobs.append(raw)
    js = json.dumps(hdr, separators=(",", ":")).encode("utf-8")
    js += b" " * (-len(js) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(js)))
        f.write(js)
        for raw in blobs:
            f.write(raw)


# ------------------------------------------------------------------------------------------------------------
# config.json (src/models/config.rs:205-353)
# ------------------------------------------------------------------------------------------------------------

MODEL_TYPES = ("base", "custom_voice", "voice_design")


def _u(node, key, default):
    """`v[key].as_u64().unwrap_or(default)`: anything that is not a non-negative JSON integer takes the default."""
    x = node.get(key) if isinstance(node, dict) else None
    return int(x) if isinstance(x, int) and not isinstance(x, bool) and x >= 0 else default


def _f(node, key, default):
    x = node.get(key) if isinstance(node, dict) else None
    return float(x) if isinstance(x, (int, float)) and not isinstance(x, bool) else default


@dataclass
class ParsedModelConfig:
    """Field names, JSON keys and defaults of `ParsedModelConfig::from_file` (config.rs:236-336)."""
    model_type: str = "base"
    model_size: str = "unknown"
    talker_hidden_size: int = 1024
    talker_intermediate_size: int = 3072
    talker_num_hidden_layers: int = 28
    talker_num_attention_heads: int = 16
    talker_num_key_value_heads: int = 8
    talker_head_dim: int = 128
    talker_vocab_size: int = 3072
    talker_text_vocab_size: int = 151936
    talker_text_hidden_size: int = 2048
    talker_rms_norm_eps: float = 1e-6
    talker_rope_theta: float = 1000000.0
    talker_max_position_embeddings: int = 32768
    mrope_section: Optional[Tuple[int, int, int]] = None
    cp_hidden_size: int = 1024
    cp_intermediate_size: int = 3072
    cp_num_hidden_layers: int = 5
    cp_num_attention_heads: int = 16
    cp_num_key_value_heads: int = 8
    cp_head_dim: int = 128
    cp_vocab_size: int = 2048
    cp_num_code_groups: int = 16
    cp_rms_norm_eps: float = 1e-6
    cp_rope_theta: float = 1000000.0
    speaker_enc_dim: Optional[int] = None          # Some(SpeakerEncoderConfig{enc_dim,..}) iff the object exists
    speaker_sample_rate: Optional[int] = None

    @classmethod
    def from_json(cls, text: str) -> "ParsedModelConfig":
        v = json.loads(text)
        if not isinstance(v, dict):
            v = {}
        mt = v.get("tts_model_type")
        model_type = mt if mt in ("custom_voice", "voice_design") else "base"
        ms = v.get("tts_model_size")
        t = v.get("talker_config") if isinstance(v.get("talker_config"), dict) else {}
        cp = t.get("code_predictor_config") if isinstance(t.get("code_predictor_config"), dict) else {}
        sec = None
        rs = t.get("rope_scaling")
        arr = rs.get("mrope_section") if isinstance(rs, dict) else None
        if isinstance(arr, list) and len(arr) == 3 and all(isinstance(a, int) and not isinstance(a, bool) and a >= 0
                                                           for a in arr):
            sec = (arr[0], arr[1], arr[2])
        se = v.get("speaker_encoder_config")
        return cls(
            model_type=model_type,
            model_size=ms if isinstance(ms, str) else "unknown",
            talker_hidden_size=_u(t, "hidden_size", 1024),
            talker_intermediate_size=_u(t, "intermediate_size", 3072),
            talker_num_hidden_layers=_u(t, "num_hidden_layers", 28),
            talker_num_attention_heads=_u(t, "num_attention_heads", 16),
            talker_num_key_value_heads=_u(t, "num_key_value_heads", 8),
            talker_head_dim=_u(t, "head_dim", 128),
            talker_vocab_size=_u(t, "vocab_size", 3072),
            talker_text_vocab_size=_u(t, "text_vocab_size", 151936),
            talker_text_hidden_size=_u(t, "text_hidden_size", 2048),
            talker_rms_norm_eps=_f(t, "rms_norm_eps", 1e-6),
            talker_rope_theta=_f(t, "rope_theta", 1000000.0),
            talker_max_position_embeddings=_u(t, "max_position_embeddings", 32768),
            mrope_section=sec,
            cp_hidden_size=_u(cp, "hidden_size", 1024),
            cp_intermediate_size=_u(cp, "intermediate_size", 3072),
            cp_num_hidden_layers=_u(cp, "num_hidden_layers", 5),
            cp_num_attention_heads=_u(cp, "num_attention_heads", 16),
            cp_num_key_value_heads=_u(cp, "num_key_value_heads", 8),
            cp_head_dim=_u(cp, "head_dim", 128),
            cp_vocab_size=_u(cp, "vocab_size", 2048),
            cp_num_code_groups=_u(cp, "num_code_groups", 16),
            cp_rms_norm_eps=_f(cp, "rms_norm_eps", 1e-6),
            cp_rope_theta=_f(cp, "rope_theta", 1000000.0),
            speaker_enc_dim=_u(se, "enc_dim", 1024) if isinstance(se, dict) else None,
            speaker_sample_rate=_u(se, "sample_rate", 24000) if isinstance(se, dict) else None,
        )

    @classmethod
    def from_file(cls, path: str) -> "ParsedModelConfig":
        try:
            text = open(path, "r", encoding="utf-8").read()
        except OSError as e:
            raise OSError(f"Failed to read config from {path}: {e}") from e
        try:
            return cls.from_json(text)
        except json.JSONDecodeError as e:
            raise ValueError(f"Failed to parse config from {path}: {e}") from e

    def label(self) -> str:
        """config.rs:339-351, e.g. "1.7B CustomVoice"."""
        size = {"0b6": "0.6B", "1b7": "1.7B"}.get(self.model_size, self.model_size)
        variant = {"base": "Base", "custom_voice": "CustomVoice", "voice_design": "VoiceDesign"}[self.model_type]
        return f"{size} {variant}"

    def to_spec(self, name: Optional[str] = None, vocoder: Optional[VocoderSpec] = None) -> ModelSpec:
        """TalkerConfig::from_parsed + CodePredictorConfig::from_parsed (talker.rs:237-254, code_predictor.rs:72-91)
        as the dimension table the C ABI takes.  The CUDA kernels are specialised for head_dim 128 and plain RoPE
        (the MRoPE sections are degenerate for TTS: all three position streams are equal); anything else is refused
        here rather than at the first launch."""
        if self.talker_head_dim != 128 or self.cp_head_dim != 128:
            raise ValueError(f"head_dim {self.talker_head_dim}/{self.cp_head_dim}: the decode kernels are built for 128")
        if self.mrope_section is not None and sum(self.mrope_section) != self.talker_head_dim // 2:
            raise ValueError(f"mrope_section {self.mrope_section} does not cover head_dim/2 = {self.talker_head_dim // 2}")
        if self.cp_rope_theta != self.talker_rope_theta or self.cp_rms_norm_eps != self.talker_rms_norm_eps:
            raise ValueError("code predictor rope_theta / rms_norm_eps differ from the talker's: not supported by the C ABI")
        return ModelSpec(
            name=name or self.label(), hidden=self.talker_hidden_size, inter=self.talker_intermediate_size,
            layers=self.talker_num_hidden_layers, heads=self.talker_num_attention_heads,
            kv_heads=self.talker_num_key_value_heads, head_dim=self.talker_head_dim,
            codec_vocab=self.talker_vocab_size, text_vocab=self.talker_text_vocab_size,
            text_embed_dim=self.talker_text_hidden_size, rope_theta=self.talker_rope_theta,
            rms_eps=self.talker_rms_norm_eps, cp_hidden=self.cp_hidden_size, cp_inter=self.cp_intermediate_size,
            cp_layers=self.cp_num_hidden_layers, cp_heads=self.cp_num_attention_heads,
            cp_kv_heads=self.cp_num_key_value_heads, cp_vocab=self.cp_vocab_size, groups=self.cp_num_code_groups,
            vocoder=vocoder or VocoderSpec())


def config_json_for_spec(spec: ModelSpec, model_type: str = "custom_voice") -> str:
    """A config.json with the keys the reference parses, for a given dimension table (used when exporting the
    synthetic checkpoints; `tts_model_size` follows the published names where the size is a published one)."""
    if model_type not in MODEL_TYPES:
        raise ValueError(model_type)
    size = {"0.6b": "0b6", "1.7b": "1b7"}.get(spec.name, spec.name)
    cfg = {
        "tts_model_type": model_type, "tts_model_size": size,
        "talker_config": {
            "hidden_size": spec.hidden, "intermediate_size": spec.inter, "num_hidden_layers": spec.layers,
            "num_attention_heads": spec.heads, "num_key_value_heads": spec.kv_heads, "head_dim": spec.head_dim,
            "vocab_size": spec.codec_vocab, "text_vocab_size": spec.text_vocab, "text_hidden_size": spec.text_embed_dim,
            "rms_norm_eps": spec.rms_eps, "rope_theta": spec.rope_theta, "max_position_embeddings": 32768,
            "rope_scaling": {"mrope_section": [24, 20, 20]},
            "code_predictor_config": {
                "hidden_size": spec.cp_hidden, "intermediate_size": spec.cp_inter, "num_hidden_layers": spec.cp_layers,
                "num_attention_heads": spec.cp_heads, "num_key_value_heads": spec.cp_kv_heads, "head_dim": spec.head_dim,
                "vocab_size": spec.cp_vocab, "num_code_groups": spec.groups, "rms_norm_eps": spec.rms_eps,
                "rope_theta": spec.rope_theta,
            },
        },
    }
    return json.dumps(cfg, indent=2)


_VOC_KEYS = {   # speech_tokenizer/config.json "decoder_config" key -> VocoderSpec field
    "codebook_dim": "codebook_dim", "vq_dim": "vq_dim", "latent_dim": "latent_dim", "hidden_size": "hidden_size",
    "num_hidden_layers": "num_layers", "num_attention_heads": "num_heads", "head_dim": "head_dim",
    "intermediate_size": "intermediate_size", "num_quantizers": "num_quantizers", "codebook_size": "codebook_size",
    "decoder_dim": "decoder_dim", "rms_norm_eps": "rms_norm_eps", "rope_theta": "rope_theta",
    "layer_scale_initial_scale": "layer_scale",
}


def vocoder_config_json(v: VocoderSpec) -> str:
    """speech_tokenizer/config.json for a vocoder dimension table.  The reference never reads this file
    (`Decoder12Hz::from_weights(.., Default::default())`, lib.rs:345): it exists so that the scaled-down synthetic
    checkpoints of the tests describe themselves; for the published dimensions it can be absent."""
    d = {k: getattr(v, f) for k, f in _VOC_KEYS.items()}
    d["upsample_rates"], d["upsampling_ratios"] = list(v.upsample_rates), list(v.upsampling_ratios)
    return json.dumps({"decoder_config": d}, indent=2)


def vocoder_spec_from_json(text: str) -> VocoderSpec:
    """Inverse of vocoder_config_json; absent keys keep Decoder12HzConfig::default (decoder_12hz.rs:47-67)."""
    v = json.loads(text)
    d = v.get("decoder_config") if isinstance(v, dict) and isinstance(v.get("decoder_config"), dict) else {}
    kw = {f: type(getattr(VocoderSpec(), f))(d[k]) for k, f in _VOC_KEYS.items() if k in d}
    for k in ("upsample_rates", "upsampling_ratios"):
        if isinstance(d.get(k), list):
            kw[k] = tuple(int(x) for x in d[k])
    return VocoderSpec(**kw)


def detect_spec_from_weights(shapes: Dict[str, Sequence[int]]) -> ModelSpec:
    """detect_talker_config (lib.rs:370-381): `talker.model.norm.weight` of length 2048 means the 1.7B dimension
    table, anything else the 0.6B one.  `shapes` maps tensor name -> shape (a safetensors header is enough)."""
    sh = shapes.get("talker.model.norm.weight")
    if sh is None:
        raise KeyError("Missing talker.model.norm.weight")
    return S.SPEC_1_7B if int(sh[0]) == 2048 else S.SPEC_0_6B


# ------------------------------------------------------------------------------------------------------------
# checkpoint directory -> the tensors of the decode path
# ------------------------------------------------------------------------------------------------------------


def hot_path_tensor_names(spec: ModelSpec) -> Tuple[List[str], List[str]]:
    """(names read from model.safetensors, names read from speech_tokenizer/model.safetensors).  Everything else in
    the two files belongs to the voice-clone front end: `speaker_encoder.*` is uploaded when present
    (load_checkpoint), `encoder.*` of the speech tokenizer (the Mimi speech encoder) is not."""
    return ([n for n, _, _ in W.talker_tensor_specs(spec)], [n for n, _, _ in W.vocoder_tensor_specs(spec.vocoder)])


def resolve_checkpoint_paths(model_dir: str) -> Tuple[Optional[str], str, str]:
    """(config.json or None, model.safetensors, speech tokenizer weights) with the reference's look-up rules and error
    texts (lib.rs:200-254): the speech tokenizer may sit in the model directory or beside it."""
    cfg = os.path.join(model_dir, "config.json")
    model = os.path.join(model_dir, "model.safetensors")
    if not os.path.exists(model):
        raise FileNotFoundError(f"Model weights not found at {model}. Please download the model first.")
    st = os.path.join(model_dir, "speech_tokenizer", "model.safetensors")
    if not os.path.exists(st):
        parent = os.path.dirname(os.path.normpath(model_dir))
        st = os.path.join(parent, "speech_tokenizer", "model.safetensors")
        if not os.path.exists(st):
            raise FileNotFoundError("Speech tokenizer weights not found")
    return (cfg if os.path.exists(cfg) else None), model, st


@dataclass
class Checkpoint:
    spec: ModelSpec
    config: Optional[ParsedModelConfig]
    talker_weights: Dict[str, torch.Tensor]      # stored dtype (bf16 in the published checkpoints); uploaded as bf16
    vocoder_weights: Dict[str, torch.Tensor]     # uploaded as f32 ("always F32", lib.rs:344-345)
    speaker_weights: Dict[str, torch.Tensor] = None   # speaker_encoder.* when the checkpoint has them (Base models), else {}


def load_checkpoint(model_dir: str) -> Checkpoint:
    """The file side of `Qwen3TTS::from_pretrained` (lib.rs:183-262) for the decode path: parse config.json when present
    (a config that fails to parse falls back to weight inspection, lib.rs:203-216), then read exactly the tensors
    the hot path needs.  A tensor the dimension table requires but the file lacks is reported by name."""
    cfg_path, model_path, st_path = resolve_checkpoint_paths(model_dir)
    cfg = None
    if cfg_path is not None:
        try:
            cfg = ParsedModelConfig.from_file(cfg_path)
        except (OSError, ValueError):
            cfg = None
    hdr, _ = read_safetensors_header(model_path)
    vcfg_path = os.path.join(os.path.dirname(st_path), "config.json")
    voc = vocoder_spec_from_json(open(vcfg_path, encoding="utf-8").read()) if os.path.exists(vcfg_path) else VocoderSpec()
    spec = cfg.to_spec(vocoder=voc) if cfg is not None else \
        detect_spec_from_weights({k: v["shape"] for k, v in hdr.items()})
    tnames, vnames = hot_path_tensor_names(spec)
    missing = [n for n in tnames if n not in hdr]
    if missing:
        raise KeyError(f"Missing weight: {missing[0]} (and {len(missing) - 1} more) in {model_path}")
    vhdr, _ = read_safetensors_header(st_path)
    vmissing = [n for n in vnames if n not in vhdr]
    if vmissing:
        raise KeyError(f"Missing weight: {vmissing[0]} (and {len(vmissing) - 1} more) in {st_path}")
    # try_load_speaker_encoder (lib.rs:1333-1360): present only in checkpoints that carry `speaker_encoder.*` keys
    snames = sorted(n for n in hdr if n.startswith("speaker_encoder."))
    return Checkpoint(spec, cfg, load_safetensors(model_path, tnames), load_safetensors(st_path, vnames),
                      load_safetensors(model_path, snames) if snames else {})


def export_checkpoint(model_dir: str, spec: ModelSpec, talker_weights: Dict[str, torch.Tensor],
                      vocoder_weights: Dict[str, torch.Tensor], model_type: str = "custom_voice") -> None:
    """Write a checkpoint directory in the layout `from_pretrained` reads: config.json, model.safetensors,
    speech_tokenizer/model.safetensors.  With the synthetic weights this gives the reference the very same model
    this build is measured on."""
    os.makedirs(os.path.join(model_dir, "speech_tokenizer"), exist_ok=True)
    with open(os.path.join(model_dir, "config.json"), "w", encoding="utf-8") as f:
        f.write(config_json_for_spec(spec, model_type))
    if spec.vocoder != VocoderSpec():
        with open(os.path.join(model_dir, "speech_tokenizer", "config.json"), "w", encoding="utf-8") as f:
            f.write(vocoder_config_json(spec.vocoder))
    save_safetensors(talker_weights, os.path.join(model_dir, "model.safetensors"), {"format": "pt"})
    save_safetensors(vocoder_weights, os.path.join(model_dir, "speech_tokenizer", "model.safetensors"), {"format": "pt"})


# ------------------------------------------------------------------------------------------------------------
# the file set `generate_audio` leaves in --output-dir (src/bin/generate_audio.rs:686-741)
# ------------------------------------------------------------------------------------------------------------


def max_frames_from_args(frames: int, duration: Optional[float]) -> int:
    """generate_audio.rs:137-144: `--duration` seconds override `--frames` at 12.5 frames per second, truncated."""
    return int(duration * 12.5) if duration is not None else frames


def write_generation_outputs(output_dir: str, seed: int, codes: Sequence[Sequence[int]], audio: np.ndarray,
                             text: str, input_ids: Sequence[int], temperature: float, top_k: int, top_p: float,
                             wav_path: Optional[str] = None) -> Dict[str, str]:
    """codes_seed{S}_frames{N}.bin, audio_seed{S}_frames{N}.wav / .bin and metadata_seed{S}_frames{N}.json with the
    `GenerationMetadata` fields (generate_audio.rs:123-135), N = frames actually generated.  -> the paths."""
    n = len(codes)
    audio = np.asarray(audio, dtype=np.float32).reshape(-1)
    os.makedirs(output_dir, exist_ok=True)
    stem = f"seed{seed}_frames{n}"
    paths = {"codes": os.path.join(output_dir, f"codes_{stem}.bin"),
             "wav": wav_path or os.path.join(output_dir, f"audio_{stem}.wav"),
             "audio": os.path.join(output_dir, f"audio_{stem}.bin"),
             "metadata": os.path.join(output_dir, f"metadata_{stem}.json")}
    if wav_path and os.path.dirname(wav_path):
        os.makedirs(os.path.dirname(wav_path), exist_ok=True)
    save_codes_binary(codes, paths["codes"])
    save_wav(paths["wav"], audio, 24000)
    save_audio_binary(audio, paths["audio"])
    meta = {"text": text, "seed": int(seed), "num_frames": n, "temperature": float(temperature), "top_k": int(top_k),
            "top_p": float(top_p), "input_ids": [int(i) for i in input_ids], "codes_shape": [1, 16, n],
            "audio_samples": int(audio.size), "sample_rate": 24000}
    with open(paths["metadata"], "w", encoding="utf-8") as f:
        json.dump(meta, f, indent=2)
    return paths
