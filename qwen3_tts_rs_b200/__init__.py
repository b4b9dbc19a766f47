"""B200-native (sm_100a) decode hot path of Qwen3-TTS behind the reference's session API.

The compute path is the CUDA library `libq3tts_b200.so` (csrc/, C ABI in include/q3tts.h);
importing this package does not load it -- `qwen3_tts_rs_b200.lib.load()` does and raises
if it is missing (there is no CPU fallback).
"""
from . import spec  # noqa: F401
