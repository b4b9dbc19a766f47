"""CPU tests for the wire / on-disk formats either side of the hot path.

Each test names the reference code or reference test it restates.  Independent witnesses used where one exists
offline: the stdlib `wave` module for WAV, the `safetensors` package (0.7) for checkpoints.
"""
import dataclasses
import json
import os
import struct
import wave

import numpy as np
import pytest
import torch

from qwen3_tts_rs_b200 import api, formats as F, spec as S, weights as W


# ---- codes / audio dumps ---------------------------------------------------------------------------------

def test_codes_binary_layout_and_round_trip(tmp_path):
    """generate_audio.rs:788-801: i64 LE, frame-major."""
    codes = [[f * 100 + q for q in range(16)] for f in range(3)]
    p = str(tmp_path / "codes.bin")
    F.save_codes_binary(codes, p)
    raw = open(p, "rb").read()
    assert len(raw) == 3 * 16 * 8
    assert raw[:8] == struct.pack("<q", 0) and raw[8:16] == struct.pack("<q", 1)
    assert raw[16 * 8:17 * 8] == struct.pack("<q", 100)           # frame 1, q0 follows frame 0's 16 values
    assert F.load_codes_binary(p) == codes
    # this is the transpose of codes_to_tensor's [1,16,T] layout (lib.rs:1417-1431)
    t = api.codes_to_tensor(codes)
    assert np.array_equal(np.frombuffer(raw, "<i8").reshape(3, 16).T, t[0])


def test_codes_binary_empty_and_ragged(tmp_path):
    p = str(tmp_path / "c.bin")
    F.save_codes_binary([], p)
    assert os.path.getsize(p) == 0 and F.load_codes_binary(p) == []
    open(p, "wb").write(b"\0" * 12)
    with pytest.raises(ValueError):
        F.load_codes_binary(p)
    open(p, "wb").write(b"\0" * 8 * 17)
    with pytest.raises(ValueError):
        F.load_codes_binary(p)


def test_audio_binary_and_golden_loader(tmp_path):
    """generate_audio.rs:803-813 and reference_validation.rs:15-23."""
    x = np.random.default_rng(0).standard_normal(1920 * 2).astype(np.float32)
    p = str(tmp_path / "a.bin")
    F.save_audio_binary(x, p)
    assert open(p, "rb").read()[:4] == struct.pack("<f", float(x[0]))
    assert np.array_equal(F.load_audio_binary(p), x)
    assert F.load_reference(p, (2, 1920)).shape == (2, 1920)
    with pytest.raises(ValueError):
        F.load_reference(p, (3, 1920))
    open(p, "ab").write(b"\x01\x02")                              # chunks_exact(4) ignores a ragged tail
    assert np.array_equal(F.load_audio_binary(p), x)


def test_compare_with_reference(tmp_path):
    """generate_audio.rs:816-920."""
    d = str(tmp_path)
    codes = [[(f * 7 + q) % 2048 for q in range(16)] for f in range(5)]
    audio = np.linspace(-1, 1, 5 * 1920, dtype=np.float32)
    cpath, apath = F.reference_dump_paths(d, 42, 5)
    assert os.path.basename(cpath) == "codes_seed42_frames5.bin" and os.path.basename(apath) == "audio_seed42_frames5.bin"
    rep = F.compare_with_reference(d, 42, 5, codes, audio)
    assert not rep.codes_found and not rep.audio_found
    F.save_codes_binary(codes, cpath)
    F.save_audio_binary(audio, apath)
    rep = F.compare_with_reference(d, 42, 5, codes, audio)
    assert rep.codes_match and rep.n_code_diffs == 0 and rep.n_ref_codes == 80
    assert rep.audio_found and rep.max_diff == 0.0 and rep.rmse == 0.0 and rep.n_audio_compared == audio.size
    bad = [list(fr) for fr in codes]
    bad[1][3] += 1
    bad[4][15] += 2
    a2 = audio.copy()
    a2[10] += 0.5
    rep = F.compare_with_reference(d, 42, 5, bad, a2[:-100])
    assert not rep.codes_match and rep.n_code_diffs == 2
    assert rep.first_code_diffs[0] == (1 * 16 + 3, codes[1][3], codes[1][3] + 1)
    assert rep.n_audio_compared == audio.size - 100
    assert abs(rep.max_diff - 0.5) < 1e-6
    assert abs(rep.mean_diff - 0.5 / (audio.size - 100)) < 1e-9
    assert abs(rep.rmse - np.sqrt(0.25 / (audio.size - 100))) < 1e-9
    rep = F.compare_with_reference(d, 42, 5, codes[:4], audio)   # same prefix, different length -> mismatch
    assert not rep.codes_match and rep.n_code_diffs == 0 and rep.n_our_codes == 64


# ---- WAV -------------------------------------------------------------------------------------------------

def test_pcm16_conversion_rule():
    """io.rs:155-160: clamp, f32 multiply by 32767, `as i16` truncates toward zero (no rounding, no dither)."""
    x = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 2.0, -3.0, 1e-5, -1e-5, 0.99999, np.nan], dtype=np.float32)
    got = F.pcm_f32_to_i16(x).tolist()
    assert got == [0, 16383, -16383, 32767, -32767, 32767, -32767, 0, 0, 32766, 0]
    # scalar restatement over random data
    r = np.random.default_rng(1).uniform(-1.2, 1.2, 5000).astype(np.float32)
    want = [int(np.float32(min(max(v, np.float32(-1)), np.float32(1))) * np.float32(32767.0)) for v in r]
    assert F.pcm_f32_to_i16(r).tolist() == want


def test_save_and_load_wav(tmp_path):
    """io.rs test_save_and_load_wav (:276-291): round trip within 1e-4; header is PCM16 mono at the given rate."""
    p = str(tmp_path / "t.wav")
    orig = np.array([0.1, 0.2, -0.3, 0.4, -0.5], dtype=np.float32)
    api.AudioBuffer(orig, 24000).save(p)
    raw = open(p, "rb").read()
    assert len(raw) == 44 + 10 and raw[:4] == b"RIFF" and struct.unpack("<I", raw[4:8])[0] == len(raw) - 8
    assert struct.unpack("<HHIIHH", raw[20:36]) == (1, 1, 24000, 48000, 2, 16)
    back = api.AudioBuffer.load(p)
    assert back.sample_rate == 24000 and len(back) == 5 and not back.is_empty()
    assert np.abs(back.samples - orig).max() < 1e-4
    with wave.open(p, "rb") as w:                                 # independent reader
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 24000, 5)
        assert np.array_equal(np.frombuffer(w.readframes(5), "<i2"), F.pcm_f32_to_i16(orig))
    F.save_wav(p, np.array([0.0, 0.5, 1.0, -0.5, -1.0], dtype=np.float32), 16000)   # io.rs test_save_wav_function
    assert F.load_wav(p)[1] == 16000
    with pytest.raises(OSError):                                  # io.rs test_load_nonexistent_file
        F.load_wav("/nonexistent/path/to/file.wav")
    open(p, "wb").write(b"not a wav file at all")
    with pytest.raises(ValueError):
        F.load_wav(p)


def test_load_wav_variants(tmp_path):
    """io.rs:110-141: int PCM scaled by 2^(bits-1), float passthrough, channels averaged."""
    p = str(tmp_path / "s.wav")
    with wave.open(p, "wb") as w:                                 # stereo PCM16 written by the stdlib
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(48000)
        w.writeframes(np.array([[16384, 0], [-32768, -32768], [100, 300]], dtype="<i2").tobytes())
    x, r = F.load_wav(p)
    assert r == 48000 and np.allclose(x, [0.25, -1.0, 200 / 32768.0])
    with wave.open(p, "wb") as w:                                 # 24-bit mono
        w.setnchannels(1); w.setsampwidth(3); w.setframerate(24000)
        w.writeframes((1 << 22).to_bytes(3, "little", signed=True) + (-(1 << 23)).to_bytes(3, "little", signed=True))
    x, _ = F.load_wav(p)
    assert np.allclose(x, [0.5, -1.0])
    data = np.array([0.125, -0.75], dtype="<f4").tobytes()        # IEEE float, with an extra chunk before `data`
    fmt = struct.pack("<HHIIHH", 3, 1, 24000, 96000, 4, 32)
    body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + b"LIST" + struct.pack("<I", 3) + b"abc\0" + \
        b"data" + struct.pack("<I", len(data)) + data
    open(p, "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    x, _ = F.load_wav(p)
    assert x.tolist() == [0.125, -0.75]


def test_normalize_rules():
    """io.rs tests :200-236."""
    b = api.AudioBuffer(np.array([0.5, -0.25, 0.1], dtype=np.float32))
    b.normalize()
    assert np.allclose(b.samples, [1.0, -0.5, 0.2], atol=1e-6)
    b = api.AudioBuffer(np.array([1.0, -1.0, 0.5], dtype=np.float32))
    b.normalize()
    assert np.allclose(b.samples, [1.0, -1.0, 0.5])
    b = api.AudioBuffer(np.zeros(3, dtype=np.float32))
    b.normalize()
    assert not b.samples.any()
    b = api.AudioBuffer(np.array([0.5, -0.5, 0.25], dtype=np.float32))
    b.normalize_db(-6.0)
    assert abs(np.abs(b.samples).max() - 0.501187) < 0.01
    assert abs(api.AudioBuffer(np.zeros(48000, np.float32), 24000).duration() - 2.0) < 1e-6


# ---- safetensors -----------------------------------------------------------------------------------------

def _sample_tensors():
    g = torch.Generator().manual_seed(3)
    return {
        "talker.model.norm.weight": torch.randn(64, generator=g).to(torch.bfloat16),
        "decoder.pre_conv.conv.weight": torch.randn(8, 4, 3, generator=g),
        "ids": torch.arange(-3, 9, dtype=torch.int64).reshape(3, 4),
        "half": torch.randn(5, generator=g).to(torch.float16),
        "scalar": torch.tensor(1.5),
        "empty": torch.zeros(0, 7),
    }


def test_safetensors_round_trip_and_cross_check(tmp_path):
    st = pytest.importorskip("safetensors.torch")
    ts = _sample_tensors()
    ours, theirs = str(tmp_path / "ours.safetensors"), str(tmp_path / "theirs.safetensors")
    F.save_safetensors(ts, ours, {"format": "pt"})
    st.save_file(ts, theirs, {"format": "pt"})
    a = st.load_file(ours)                 # their reader on our file
    b = F.load_safetensors(theirs)         # our reader on their file
    c = F.load_safetensors(ours)
    for k, v in ts.items():
        for got in (a[k], b[k], c[k]):
            assert got.dtype == v.dtype and tuple(got.shape) == tuple(v.shape), k
            assert torch.equal(got.reshape(-1).view(torch.uint8) if v.numel() else got, v.reshape(-1).view(torch.uint8) if v.numel() else v), k
    n = struct.unpack("<Q", open(ours, "rb").read(8))[0]
    assert n % 8 == 0                      # header padded to 8 bytes like the library's
    sub = F.load_safetensors(theirs, ["ids", "not_there"])
    assert list(sub) == ["ids"]


def test_safetensors_rejects_corrupt_files(tmp_path):
    p = str(tmp_path / "x.safetensors")
    F.save_safetensors({"w": torch.ones(4, 4)}, p)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-8])                                  # data section truncated
    with pytest.raises(ValueError):
        F.load_safetensors(p)
    open(p, "wb").write(struct.pack("<Q", 1 << 40) + raw[8:])      # absurd header length
    with pytest.raises(ValueError):
        F.load_safetensors(p)
    open(p, "wb").write(b"\x01\x02")
    with pytest.raises(ValueError):
        F.load_safetensors(p)
    hdr = json.dumps({"w": {"dtype": "F32", "shape": [4, 4], "data_offsets": [0, 60]}}).encode()
    open(p, "wb").write(struct.pack("<Q", len(hdr)) + hdr + b"\0" * 64)   # offsets disagree with dtype*shape
    with pytest.raises(ValueError):
        F.load_safetensors(p)


# ---- config.json -----------------------------------------------------------------------------------------

def _hf_config(size, variant):
    """The keys the published checkpoints carry that the reference reads (config.rs:244-306)."""
    h, i = (1024, 3072) if size == "0b6" else (2048, 6144)
    cfg = {
        "architectures": ["Qwen3TTSForConditionalGeneration"], "tts_model_type": variant, "tts_model_size": size,
        "talker_config": {
            "hidden_size": h, "intermediate_size": i, "num_hidden_layers": 28, "num_attention_heads": 16,
            "num_key_value_heads": 8, "head_dim": 128, "vocab_size": 3072, "text_vocab_size": 151936,
            "text_hidden_size": 2048, "rms_norm_eps": 1e-6, "rope_theta": 1000000, "max_position_embeddings": 32768,
            "rope_scaling": {"mrope_section": [24, 20, 20], "interleaved": True, "rope_type": "default"},
            "code_predictor_config": {"hidden_size": 1024, "intermediate_size": 3072, "num_hidden_layers": 5,
                                      "num_attention_heads": 16, "num_key_value_heads": 8, "head_dim": 128,
                                      "vocab_size": 2048, "num_code_groups": 16, "rms_norm_eps": 1e-6,
                                      "rope_theta": 1000000},
        },
    }
    if variant == "base":
        cfg["speaker_encoder_config"] = {"enc_dim": h, "sample_rate": 24000}
    return json.dumps(cfg)


@pytest.mark.parametrize("size,variant,label,spec", [
    ("0b6", "base", "0.6B Base", S.SPEC_0_6B), ("1b7", "base", "1.7B Base", S.SPEC_1_7B),
    ("0b6", "custom_voice", "0.6B CustomVoice", S.SPEC_0_6B), ("1b7", "voice_design", "1.7B VoiceDesign", S.SPEC_1_7B)])
def test_parsed_model_config_variants(size, variant, label, spec):
    """config.rs tests :625-688, on configs with the published dimensions."""
    cfg = F.ParsedModelConfig.from_json(_hf_config(size, variant))
    assert cfg.model_type == variant and cfg.model_size == size and cfg.label() == label
    assert cfg.talker_hidden_size == spec.hidden and cfg.talker_intermediate_size == spec.inter
    assert cfg.cp_hidden_size == 1024 and cfg.mrope_section == (24, 20, 20)
    assert (cfg.speaker_enc_dim == spec.hidden) if variant == "base" else (cfg.speaker_enc_dim is None)
    got = cfg.to_spec(name=spec.name)
    assert got == spec                                            # the whole dimension table, field for field
    assert got.has_cp_proj == (size == "1b7")


def test_parsed_model_config_defaults_and_odd_values():
    """`unwrap_or` defaults (config.rs:244-306): absent, null or wrongly typed keys take the 0.6B Base values."""
    cfg = F.ParsedModelConfig.from_json("{}")
    assert cfg == F.ParsedModelConfig() and cfg.label() == "unknown Base" and cfg.mrope_section is None
    assert cfg.to_spec(name="0.6b") == S.SPEC_0_6B
    cfg = F.ParsedModelConfig.from_json(json.dumps({
        "tts_model_type": "something_else", "tts_model_size": 17,
        "talker_config": {"hidden_size": "2048", "num_hidden_layers": -3, "rms_norm_eps": 1e-5, "rope_theta": 10000,
                          "rope_scaling": {"mrope_section": [24, 20]}, "code_predictor_config": None}}))
    assert cfg.model_type == "base" and cfg.model_size == "unknown"
    assert cfg.talker_hidden_size == 1024 and cfg.talker_num_hidden_layers == 28
    assert cfg.talker_rms_norm_eps == 1e-5 and cfg.talker_rope_theta == 10000.0 and cfg.mrope_section is None
    with pytest.raises(ValueError):                               # CP constants differ from the talker's: refused, not ignored
        cfg.to_spec()
    with pytest.raises(ValueError):
        F.ParsedModelConfig.from_json(json.dumps({"talker_config": {"head_dim": 64}})).to_spec()


def test_config_file_errors_and_generated_config(tmp_path):
    with pytest.raises(OSError, match="Failed to read config"):   # config.rs:239-240
        F.ParsedModelConfig.from_file(str(tmp_path / "nope.json"))
    p = tmp_path / "config.json"
    p.write_text("{ not json")
    with pytest.raises(ValueError, match="Failed to parse config"):
        F.ParsedModelConfig.from_file(str(p))
    for spec, mt in ((S.SPEC_1_7B, "custom_voice"), (S.SPEC_0_6B, "base"), (S.SPEC_TINY_PROJ, "voice_design")):
        p.write_text(F.config_json_for_spec(spec, mt))
        cfg = F.ParsedModelConfig.from_file(str(p))
        assert cfg.model_type == mt and cfg.to_spec(name=spec.name, vocoder=spec.vocoder) == spec
        assert F.vocoder_spec_from_json(F.vocoder_config_json(spec.vocoder)) == spec.vocoder
    assert F.vocoder_spec_from_json("{}") == S.VocoderSpec()       # Decoder12HzConfig::default
    assert F.ParsedModelConfig.from_json(F.config_json_for_spec(S.SPEC_1_7B)).label() == "1.7B CustomVoice"


def test_detect_spec_from_weights():
    """lib.rs:370-381."""
    assert F.detect_spec_from_weights({"talker.model.norm.weight": [2048]}) is S.SPEC_1_7B
    assert F.detect_spec_from_weights({"talker.model.norm.weight": [1024]}) is S.SPEC_0_6B
    with pytest.raises(KeyError, match="Missing talker.model.norm.weight"):
        F.detect_spec_from_weights({})


# ---- checkpoint directory --------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def tiny_checkpoint(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("ckpt") / "tiny")
    spec = S.SPEC_TINY
    tw, vw = W.make_talker_weights(spec), W.make_vocoder_weights(spec.vocoder)
    extra = dict(tw)
    extra["speaker_encoder.blocks.0.conv.weight"] = torch.zeros(4, 4)     # front-end tensors share the file
    F.export_checkpoint(d, spec, extra, dict(vw, **{"encoder.downsample.conv.weight": torch.zeros(2, 2, 2)}))
    return d, spec, tw, vw


def test_load_checkpoint_reads_exactly_the_hot_path(tiny_checkpoint):
    d, spec, tw, vw = tiny_checkpoint
    assert sorted(os.listdir(d)) == ["config.json", "model.safetensors", "speech_tokenizer"]
    ck = F.load_checkpoint(d)
    assert ck.config.model_type == "custom_voice" and ck.spec == dataclasses.replace(spec, name=ck.spec.name)
    assert set(ck.talker_weights) == set(tw) and set(ck.vocoder_weights) == set(vw)
    for k in tw:
        assert ck.talker_weights[k].dtype == torch.bfloat16 and torch.equal(ck.talker_weights[k].view(torch.int16),
                                                                            tw[k].view(torch.int16)), k
    for k in vw:
        assert ck.vocoder_weights[k].dtype == torch.float32 and torch.equal(ck.vocoder_weights[k], vw[k]), k


def test_load_checkpoint_lookup_rules_and_errors(tiny_checkpoint, tmp_path):
    """lib.rs:200-254: error texts, speech tokenizer beside the model directory, config.json optional."""
    d, spec, tw, vw = tiny_checkpoint
    with pytest.raises(FileNotFoundError, match="Model weights not found at .*model.safetensors. Please download the model first."):
        F.load_checkpoint(str(tmp_path / "missing"))
    m = tmp_path / "root" / "model"
    m.mkdir(parents=True)
    F.save_safetensors(tw, str(m / "model.safetensors"))
    with pytest.raises(FileNotFoundError, match="Speech tokenizer weights not found"):
        F.load_checkpoint(str(m))
    (tmp_path / "root" / "speech_tokenizer").mkdir()
    F.save_safetensors(vw, str(tmp_path / "root" / "speech_tokenizer" / "model.safetensors"))
    (tmp_path / "root" / "speech_tokenizer" / "config.json").write_text(F.vocoder_config_json(spec.vocoder))
    # no config.json: weight inspection picks the 0.6B table (norm length != 2048), whose tensors this file lacks
    with pytest.raises(KeyError, match="Missing weight: talker.model.layers.3"):
        F.load_checkpoint(str(m))
    (m / "config.json").write_text(F.config_json_for_spec(spec))
    ck = F.load_checkpoint(str(m) + "/")                          # trailing slash: parent look-up still works
    assert set(ck.talker_weights) == set(tw) and ck.spec.hidden == spec.hidden
    (m / "config.json").write_text("{ broken")                    # unparsable config falls back to inspection
    with pytest.raises(KeyError, match="Missing weight"):
        F.load_checkpoint(str(m))
    short = {k: v for k, v in tw.items() if k != "talker.codec_head.weight"}
    F.save_safetensors(short, str(m / "model.safetensors"))
    (m / "config.json").write_text(F.config_json_for_spec(spec))
    with pytest.raises(KeyError, match="Missing weight: talker.codec_head.weight"):
        F.load_checkpoint(str(m))


def test_formats_match_the_c_oracle(tmp_path):
    """Product-side byte formats against oracle/c's plain-C restatement (codes dump, PCM16 rule) on seeded inputs."""
    import ctypes as C
    from oracle import build_ref
    lib = C.CDLL(build_ref.build_c())
    rng = np.random.default_rng(7)
    codes = rng.integers(0, 3072, size=(37, 16), dtype=np.uint32)
    want = np.zeros(37 * 16 * 8, dtype=np.uint8)
    lib.q3o_codes_dump(codes.ctypes.data_as(C.c_void_p), 37, want.ctypes.data_as(C.c_void_p))
    p = str(tmp_path / "c.bin")
    F.save_codes_binary(codes.tolist(), p)
    assert open(p, "rb").read() == want.tobytes()
    x = np.concatenate([rng.uniform(-1.5, 1.5, 100000), [1.0, -1.0, 0.0, -0.0, 1e-30]]).astype(np.float32)
    pcm = np.zeros(x.size, dtype=np.int16)
    lib.q3o_pcm16(x.ctypes.data_as(C.c_void_p), x.size, pcm.ctypes.data_as(C.c_void_p))
    assert np.array_equal(F.pcm_f32_to_i16(x), pcm)


def test_generate_audio_file_set(tmp_path):
    """generate_audio.rs:137-144 (frames from --duration) and :686-741 (file names, metadata fields)."""
    assert F.max_frames_from_args(2048, None) == 2048 and F.max_frames_from_args(2048, 10.0) == 125
    assert F.max_frames_from_args(7, 0.1) == 1                     # (0.1 * 12.5) as usize
    codes = [[q for q in range(16)] for _ in range(3)]
    audio = np.zeros(3 * 1920, np.float32)
    paths = F.write_generation_outputs(str(tmp_path / "o"), 42, codes, audio, "Hello", [9, 8], 0.7, 50, 0.9)
    assert sorted(os.listdir(tmp_path / "o")) == ["audio_seed42_frames3.bin", "audio_seed42_frames3.wav",
                                                  "codes_seed42_frames3.bin", "metadata_seed42_frames3.json"]
    meta = json.load(open(paths["metadata"]))
    assert list(meta) == ["text", "seed", "num_frames", "temperature", "top_k", "top_p", "input_ids", "codes_shape",
                          "audio_samples", "sample_rate"]
    assert meta["codes_shape"] == [1, 16, 3] and meta["audio_samples"] == 5760 and meta["sample_rate"] == 24000
    rep = F.compare_with_reference(str(tmp_path / "o"), 42, 3, codes, audio)   # a run compares clean against itself
    assert rep.codes_match and rep.max_diff == 0.0
    paths = F.write_generation_outputs(str(tmp_path / "o"), 1, codes, audio, "", [], 0.7, 50, 0.9,
                                       wav_path=str(tmp_path / "elsewhere" / "x.wav"))
    assert os.path.exists(tmp_path / "elsewhere" / "x.wav") and not os.path.exists(tmp_path / "o" / "audio_seed1_frames3.wav")


def test_generate_audio_tool_exports_a_loadable_checkpoint(tmp_path):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = str(tmp_path / "synth")
    subprocess.run([sys.executable, os.path.join(root, "tools", "generate_audio.py"), "--export-synthetic", "tiny_proj",
                    "--model-dir", d, "--model-type", "voice_design"], check=True, capture_output=True)
    ck = F.load_checkpoint(d)
    assert ck.config.model_type == "voice_design" and ck.spec.has_cp_proj and ck.spec.vocoder == S.TINY_VOCODER
    assert "talker.code_predictor.small_to_mtp_projection.bias" in ck.talker_weights
