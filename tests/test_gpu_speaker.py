"""GPU parity of the ECAPA-TDNN speaker encoder (voice-clone front end) against the oracle's F32
restatement of SpeakerEncoder::forward (speaker.rs:448-476), through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import speaker as OS
from qwen3_tts_rs_b200 import api, spec as S, weights as W

pytestmark = pytest.mark.gpu


def _encoder(cfg):
    w = W.make_speaker_weights(cfg)
    m = api.Model(S.SPEC_TINY)
    m.load(w).finalize()                       # a model may hold the speaker encoder alone
    return api.Qwen3TTS(m), OS.SpeakerEncoder(cfg, w)


@pytest.mark.parametrize("cfg,T", [(S.TINY_SPEAKER, 37), (S.SpeakerSpec(), 100), (S.SpeakerSpec(), 333)], ids=["tiny-37", "full-100", "full-333"])
def test_speaker_embedding_matches_the_oracle(cfg, T):
    """Tolerance: the 1x1 convs run on the tensor cores with the bf16x3 split (relative error ~2^-16 per product), everything
    else in F32 with a different summation order than torch: rms(cuda - oracle) <= 1e-4 rms(oracle), max <= 1e-3 rms."""
    tts, ref = _encoder(cfg)
    g = torch.Generator().manual_seed(T)
    mel = torch.randn(2, cfg.mel_dim, T, generator=g) * 2.0 - 3.0        # log-mel-like range
    got = tts.speaker_encode(mel.numpy())
    want = ref.forward(mel)
    assert got.shape == want.shape == (2, cfg.enc_dim)
    rms = float(want.pow(2).mean().sqrt())
    err = got - want
    assert float(err.pow(2).mean().sqrt()) <= 1e-4 * rms, (float(err.pow(2).mean().sqrt()), rms)
    assert float(err.abs().max()) <= 1e-3 * rms
    # rows are independent (one utterance at a time, as SpeakerEncoder::encode)
    assert torch.equal(tts.speaker_encode(mel[1:].numpy())[0], got[1])


def test_speaker_encoder_errors():
    tts, _ = _encoder(S.TINY_SPEAKER)
    with pytest.raises(api.L.Q3Error):           # shorter than the reflect padding of the k = 5 conv
        tts.speaker_encode(np.zeros((1, S.TINY_SPEAKER.mel_dim, 2), np.float32))
    plain = api.Model(S.SPEC_TINY)
    plain.load(W.make_talker_weights(S.SPEC_TINY)).finalize()
    with pytest.raises(api.L.Q3Error):
        api.Qwen3TTS(plain).speaker_encode(np.zeros((1, 32, 16), np.float32))


def test_speaker_embedding_feeds_the_voice_clone_prompt():
    """End of the x-vector path: mel -> ECAPA embedding -> voice-clone prefill -> codes; the embedding's width must be the
    talker's hidden size (lib.rs:930, talker.rs:538)."""
    spec = S.SPEC_TINY
    cfg = S.SpeakerSpec(mel_dim=32, enc_dim=spec.hidden, enc_channels=(64, 64, 64, 64, 192), enc_attention_channels=32,
                        enc_res2net_scale=4, enc_se_channels=32)
    w = dict(W.make_talker_weights(spec))
    w.update(W.make_speaker_weights(cfg))
    tts = api.Qwen3TTS.from_weights(spec, w)
    emb = tts.speaker_encode(np.random.default_rng(0).standard_normal((1, 32, 50)).astype(np.float32))[0]
    ids = W.synthetic_prompt(0, spec)
    codes = tts.generate_codes_voice_clone([ids], [api.VoiceClonePrompt(emb)], options=api.SynthesisOptions(max_length=4, eos_token_id=None), seeds=[3])
    assert len(codes[0]) == 4 and all(len(f) == 16 for f in codes[0])
