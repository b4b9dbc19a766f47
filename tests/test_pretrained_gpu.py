"""GPU test of the checkpoint-directory entry point (`Qwen3TTS::from_pretrained`, src/lib.rs:183-262) and of the
file formats on the output side: a model loaded from an exported directory must behave bit for bit like the same
weights handed over in memory, and its codes / PCM must survive the reference's dump formats unchanged."""
import numpy as np
import pytest

from qwen3_tts_rs_b200 import api, formats as F, spec as S, weights as W
from conftest import talker_weights, vocoder_weights

pytestmark = pytest.mark.gpu


def test_from_pretrained_equals_from_weights(tmp_path):
    spec = S.SPEC_TINY
    tw, vw = talker_weights(spec), vocoder_weights(spec.vocoder, spec.name)
    d = str(tmp_path / "tiny-customvoice")
    F.export_checkpoint(d, spec, tw, vw, "custom_voice")
    a = api.Qwen3TTS.from_weights(spec, tw, vw)
    b = api.Qwen3TTS.from_pretrained(d)
    assert b.model_type == "custom_voice" and b.supports_preset_speakers() and not b.supports_voice_design()
    assert a.model_type is None and a.supports_preset_speakers()
    ids = [W.synthetic_prompt(i, spec) for i in range(2)]
    opts = api.SynthesisOptions(max_length=8)
    ca = a.generate_codes(ids, options=opts, seeds=[42, 43])
    cb = b.generate_codes(ids, options=opts, seeds=[42, 43])
    assert ca == cb and len(ca[0]) > 0                        # same weights, same kernels: identical, not merely close
    pa, pb = a.decode_codes(ca[0]), b.decode_codes(cb[0])
    assert np.array_equal(pa.samples, pb.samples)
    # output side: dumps and WAV
    F.save_codes_binary(cb[0], str(tmp_path / "codes_seed42_frames8.bin"))
    F.save_audio_binary(pb.samples, str(tmp_path / "audio_seed42_frames8.bin"))
    rep = F.compare_with_reference(str(tmp_path), 42, 8, ca[0], pa.samples)
    assert rep.codes_match and rep.audio_found and rep.max_diff == 0.0
    pb.save(str(tmp_path / "out.wav"))
    back = api.AudioBuffer.load(str(tmp_path / "out.wav"))
    # x -> trunc(32767 x) / 32768: off by at most (|x| + 1) / 32768 for |x| <= 1
    assert len(back) == len(pb) and np.abs(back.samples - pb.samples).max() <= 2.0 / 32768 + 1e-7
