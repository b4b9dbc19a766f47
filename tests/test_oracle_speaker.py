"""The oracle's ECAPA-TDNN speaker encoder (oracle/speaker.py) against (a) every known answer of the reference's own unit
tests (src/models/speaker.rs:478-571) and (b) transformers' ECAPA_TimeDelayNet (the Qwen2.5-Omni token2wav speaker encoder:
same block names, so the synthetic weights load by name)."""
import pytest
import torch

from oracle import speaker as OS
from qwen3_tts_rs_b200 import spec as S, weights as W


def test_reference_unit_test_vectors():
    x = torch.arange(5.0).reshape(1, 1, 5)
    assert OS.reflect_pad_1d(x, 0, 0).shape == (1, 1, 5)                                          # test_reflect_pad_1d_no_pad
    assert OS.reflect_pad_1d(x, 2, 0).flatten().tolist() == [2.0, 1.0, 0.0, 1.0, 2.0, 3.0, 4.0]   # ..._left
    assert OS.reflect_pad_1d(x, 0, 2).flatten().tolist() == [0.0, 1.0, 2.0, 3.0, 4.0, 3.0, 2.0]   # ..._right
    assert OS.reflect_pad_1d(x, 2, 2).flatten().tolist() == [2.0, 1.0, 0.0, 1.0, 2.0, 3.0, 4.0, 3.0, 2.0]   # ..._both
    assert abs(float(OS.sigmoid(torch.zeros(1))) - 0.5) < 1e-5                                     # test_sigmoid
    assert torch.relu(torch.tensor([-1.0, 0.0, 1.0, 2.0])).tolist() == [0.0, 0.0, 1.0, 2.0]        # test_relu


def test_forward_shape_of_the_default_config():
    """test_speaker_encoder_forward_shape (speaker.rs:557-570): mel [1, 128, 100] -> [1, 1024]."""
    cfg = S.SpeakerSpec()
    enc = OS.SpeakerEncoder(cfg, W.make_speaker_weights(cfg))
    out = enc.forward(torch.randn(1, 128, 100))
    assert out.shape == (1, 1024) and torch.isfinite(out).all()
    assert len(list(W.speaker_tensor_specs(cfg))) == 76


def test_matches_transformers_ecapa():
    """transformers' ECAPA_TimeDelayNet with the same weights.  Its attentive pooling clamps the variances at 1e-12 where the
    reference adds 1e-5 (speaker.rs:303, 344) -- visible with the small synthetic activations -- so the oracle is run with
    that one rule switched to transformers' for the comparison: everything else must then agree to F32 rounding."""
    mod = pytest.importorskip("transformers.models.qwen2_5_omni.modeling_qwen2_5_omni")
    cfgm = pytest.importorskip("transformers.models.qwen2_5_omni.configuration_qwen2_5_omni")
    cfg = S.TINY_SPEAKER
    hf_cfg = cfgm.Qwen2_5OmniDiTConfig(mel_dim=cfg.mel_dim, enc_dim=cfg.enc_dim, enc_channels=list(cfg.enc_channels),
                                       enc_kernel_sizes=list(cfg.enc_kernel_sizes), enc_dilations=list(cfg.enc_dilations),
                                       enc_attention_channels=cfg.enc_attention_channels, enc_res2net_scale=cfg.enc_res2net_scale,
                                       enc_se_channels=cfg.enc_se_channels)
    hf = mod.ECAPA_TimeDelayNet(hf_cfg).eval()
    w = W.make_speaker_weights(cfg)
    missing, unexpected = hf.load_state_dict({k[len("speaker_encoder."):]: v for k, v in w.items()}, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    mel = torch.randn(2, cfg.mel_dim, 41)
    with torch.no_grad():
        want = hf(mel.transpose(1, 2))           # transformers takes [B, T, mel]
    got = OS.SpeakerEncoder(cfg, w, hf_variance_rule=True).forward(mel)
    rms = float(want.pow(2).mean().sqrt())
    assert float((got - want).abs().max()) <= 2e-5 * rms, (float((got - want).abs().max()), rms)
    ref = OS.SpeakerEncoder(cfg, w).forward(mel)             # the reference's rule: a small, explained difference
    assert 0 < float((ref - want).abs().max()) <= 1e-2 * rms
