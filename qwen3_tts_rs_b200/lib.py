"""ctypes binding of libq3tts_b200.so (include/q3tts.h).

This is the only way compute reaches the GPU from Python: there is no CPU or torch fallback,
and `load()` raises if the shared library is missing (build it with
`python -m qwen3_tts_rs_b200.build` or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from .spec import ModelSpec

HERE = os.path.dirname(os.path.abspath(__file__))
# Q3TTS_LIB=dev selects the development library (historical kernel generations + profiling hooks, build.py); read once, at
# the first load of the process
LIB_PATH = os.environ.get("Q3TTS_LIB_PATH") or os.path.join(HERE, "libq3tts_b200_dev.so" if os.environ.get("Q3TTS_LIB") == "dev" else "libq3tts_b200.so")
IS_DEV = os.environ.get("Q3TTS_LIB") == "dev"

Q3_BF16, Q3_F32 = 0, 1
STATUS = {0: "Q3_OK", 1: "Q3_ERR_INVALID", 2: "Q3_ERR_CUDA", 3: "Q3_ERR_KV_OVERFLOW",
          4: "Q3_ERR_MISSING_WEIGHT", 5: "Q3_ERR_STATE", 6: "Q3_ERR_UNSUPPORTED"}

# every symbol include/q3tts.h declares (tests check the library exports all of them)
EXPORTS = [
    "q3_last_error", "q3_abi_version", "q3_kernel_launch_count",
    "q3_model_create", "q3_model_set_tensor", "q3_model_finalize", "q3_model_destroy",
    "q3_session_create", "q3_session_reset", "q3_session_destroy", "q3_session_stream", "q3_session_synchronize",
    "q3_prefill_embeds", "q3_prefill_ids", "q3_prefill_voice_clone", "q3_set_trailing_text", "q3_set_trailing_ids",
    "q3_generate", "q3_generate_async", "q3_get_codes", "q3_stream_next", "q3_session_set_stream_context", "q3_session_set_first_chunk",
    "q3_vocoder_decode", "q3_vocode_session", "q3_speaker_encode", "q3_speaker_embed_dim",
    "q3_talker_step", "q3_code_predictor_frame", "q3_sample",
    "q3_fused_residual_rmsnorm", "q3_fused_residual_rmsnorm_host", "q3_session_timing",
]


class Q3Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code
        self.status = STATUS.get(code, str(code))


class ModelDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("hidden", "inter", "layers", "heads", "kv_heads", "head_dim",
                                         "codec_vocab", "text_vocab", "text_embed_dim")] + \
               [("rope_theta", C.c_float), ("rms_eps", C.c_float)] + \
               [(n, C.c_int32) for n in ("cp_hidden", "cp_inter", "cp_layers", "cp_heads", "cp_kv_heads", "cp_vocab",
                                         "groups", "cp_rope_positions", "cp_max_seq",
                                         "v_codebook_dim", "v_vq_dim", "v_latent_dim", "v_hidden", "v_layers", "v_heads",
                                         "v_head_dim", "v_inter", "v_quantizers", "v_codebook_size", "v_decoder_dim",
                                         "v_n_upsampling")] + \
               [("v_upsampling", C.c_int32 * 4), ("v_n_rates", C.c_int32), ("v_rates", C.c_int32 * 8),
                ("v_rms_eps", C.c_float), ("v_rope_theta", C.c_float), ("device", C.c_int32)]


class GenConfig(C.Structure):
    _fields_ = [("max_new_tokens", C.c_int32), ("temperature", C.c_double), ("top_k", C.c_int32),
                ("top_p", C.c_double), ("repetition_penalty", C.c_double), ("eos_token_id", C.c_int32),
                ("min_new_tokens", C.c_int32), ("chunk_frames", C.c_int32)]


class Timing(C.Structure):
    _fields_ = [("prefill_ms", C.c_float), ("generation_ms", C.c_float), ("decode_ms", C.c_float),
                ("generation_frames", C.c_int32)]


def model_desc(spec: ModelSpec, device: int = 0) -> ModelDesc:
    v = spec.vocoder
    d = ModelDesc()
    for n in ("hidden", "inter", "layers", "heads", "kv_heads", "head_dim", "codec_vocab", "text_vocab",
              "text_embed_dim", "cp_hidden", "cp_inter", "cp_layers", "cp_heads", "cp_kv_heads", "cp_vocab", "groups",
              "cp_rope_positions", "cp_max_seq"):
        setattr(d, n, getattr(spec, n))
    d.rope_theta, d.rms_eps = spec.rope_theta, spec.rms_eps
    d.v_codebook_dim, d.v_vq_dim, d.v_latent_dim, d.v_hidden = v.codebook_dim, v.vq_dim, v.latent_dim, v.hidden_size
    d.v_layers, d.v_heads, d.v_head_dim, d.v_inter = v.num_layers, v.num_heads, v.head_dim, v.intermediate_size
    d.v_quantizers, d.v_codebook_size, d.v_decoder_dim = v.num_quantizers, v.codebook_size, v.decoder_dim
    d.v_n_upsampling = len(v.upsampling_ratios)
    for i, r in enumerate(v.upsampling_ratios):
        d.v_upsampling[i] = r
    d.v_n_rates = len(v.upsample_rates)
    for i, r in enumerate(v.upsample_rates):
        d.v_rates[i] = r
    d.v_rms_eps, d.v_rope_theta = v.rms_norm_eps, v.rope_theta
    d.device = device
    return d


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the CUDA library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m qwen3_tts_rs_b200.build`. "
                           "There is no CPU fallback for the decode hot path.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
    lib.q3_last_error.restype = C.c_char_p
    lib.q3_abi_version.restype = C.c_int
    lib.q3_kernel_launch_count.restype = C.c_uint64
    lib.q3_model_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(vp)]
    lib.q3_model_set_tensor.argtypes = [vp, C.c_char_p, vp, C.c_int, C.POINTER(C.c_int64), i32, i32]
    lib.q3_model_finalize.argtypes = [vp]
    lib.q3_model_destroy.argtypes = [vp]
    lib.q3_model_destroy.restype = None
    lib.q3_session_create.argtypes = [vp, i32, i32, C.POINTER(GenConfig), C.POINTER(u64), C.POINTER(vp)]
    lib.q3_session_reset.argtypes = [vp, C.POINTER(u64)]
    lib.q3_session_destroy.argtypes = [vp]
    lib.q3_session_destroy.restype = None
    lib.q3_session_stream.argtypes = [vp]
    lib.q3_session_stream.restype = vp
    lib.q3_session_synchronize.argtypes = [vp]
    lib.q3_prefill_embeds.argtypes = [vp, vp, vp, i32]
    lib.q3_prefill_ids.argtypes = [vp, vp, vp, vp, i32]
    lib.q3_prefill_voice_clone.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i32]
    lib.q3_speaker_encode.argtypes = [vp, vp, i32, i32, vp]
    lib.q3_speaker_embed_dim.argtypes = [vp]
    lib.q3_speaker_embed_dim.restype = i32
    lib.q3_set_trailing_text.argtypes = [vp, vp, vp, i32, vp]
    lib.q3_set_trailing_ids.argtypes = [vp, vp, vp, i32, i32, i32]
    lib.q3_generate.argtypes = [vp, i32, vp, vp]
    lib.q3_generate_async.argtypes = [vp, i32]
    lib.q3_get_codes.argtypes = [vp, i32, vp, vp]
    lib.q3_stream_next.argtypes = [vp, vp, vp, vp, C.POINTER(i32)]
    lib.q3_vocoder_decode.argtypes = [vp, vp, i32, i32, vp]
    lib.q3_vocode_session.argtypes = [vp, i32, vp]
    lib.q3_talker_step.argtypes = [vp, vp, vp, vp]
    lib.q3_code_predictor_frame.argtypes = [vp, vp, vp, vp, vp]
    lib.q3_sample.argtypes = [vp, vp, i32, i32, C.POINTER(GenConfig), vp, vp, i32, vp]
    lib.q3_fused_residual_rmsnorm.argtypes = [vp, vp, vp, vp, vp, i32, i32, C.c_float, C.c_int, vp]
    lib.q3_fused_residual_rmsnorm_host.argtypes = [vp, vp, vp, vp, vp, i32, i32, C.c_float, C.c_int, i32]
    lib.q3_session_timing.argtypes = [vp, C.POINTER(Timing)]
    lib.q3_session_set_stream_context.argtypes = [vp, i32]
    lib.q3_session_set_first_chunk.argtypes = [vp, i32]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("q3_last_error", "q3_abi_version", "q3_kernel_launch_count", "q3_model_destroy",
                        "q3_session_destroy", "q3_session_stream"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(code: int):
    if code != 0:
        raise Q3Error(code, load().q3_last_error().decode(errors="replace"))
