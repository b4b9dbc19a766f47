"""TEST INFRASTRUCTURE ONLY (oracle): torch F32 restatement of the reference's ECAPA-TDNN speaker encoder,
`SpeakerEncoder::forward` (src/models/speaker.rs:448-476) -- mel spectrogram [B, mel_dim, T] -> embedding [B, enc_dim].
The mel front end (`MelSpectrogram::compute_for_speaker_encoder`, src/audio/mel.rs) is out of scope.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def reflect_pad_1d(x: torch.Tensor, pad_left: int, pad_right: int) -> torch.Tensor:
    """speaker.rs:26-53: mirror around the first / last sample (PyTorch padding_mode="reflect"), by index_select."""
    if pad_left == 0 and pad_right == 0:
        return x
    t = x.shape[-1]
    idx = list(range(pad_left, 0, -1)) + list(range(t)) + [t - 2 - i for i in range(pad_right)]
    return x[..., torch.tensor(idx, dtype=torch.long)]


def sigmoid(x: torch.Tensor) -> torch.Tensor:
    """speaker.rs:59-63: 1 / (exp(-x) + 1)."""
    return 1.0 / (torch.exp(-x) + 1.0)


class SpeakerEncoder:
    """speaker.rs:352-476.  Weight names: speaker_encoder.{blocks.0.conv, blocks.{1,2,3}.{tdnn1.conv, res2net_block.blocks.j.conv,
    tdnn2.conv, se_block.conv1, se_block.conv2}, mfa.conv, asp.tdnn.conv, asp.conv, fc}.{weight, bias}."""

    def __init__(self, cfg, w: Dict[str, torch.Tensor], hf_variance_rule: bool = False):
        # hf_variance_rule: sqrt(clamp(var, 1e-12)) as transformers' ECAPA pooling does, instead of the reference's
        # sqrt(var + 1e-5) -- only for the cross-check in tests/test_oracle_speaker.py
        self.hf_rule = hf_variance_rule
        self.cfg = cfg
        self.w = {k: v.to(torch.float32) for k, v in w.items() if k.startswith("speaker_encoder.")}

    def _conv(self, name: str, x: torch.Tensor, dilation: int = 1) -> torch.Tensor:
        """ReflectPadConv1d (speaker.rs:68-108): "same" length, pad_left = total / 2, pad_right = total - pad_left."""
        wt, b = self.w[f"speaker_encoder.{name}.weight"], self.w[f"speaker_encoder.{name}.bias"]
        total = dilation * (wt.shape[2] - 1)
        left = total // 2
        return F.conv1d(reflect_pad_1d(x, left, total - left), wt, b, dilation=dilation)

    def _tdnn(self, name: str, x: torch.Tensor, dilation: int = 1) -> torch.Tensor:
        """TimeDelayNetBlock (speaker.rs:112-139): conv + ReLU."""
        return torch.relu(self._conv(name + ".conv", x, dilation))

    def _res2net(self, name: str, x: torch.Tensor, dilation: int) -> torch.Tensor:
        """Res2NetBlock (speaker.rs:141-190): chunk 0 passes through; chunk i+1 (+ the previous output for i > 0) -> TDNN."""
        scale = self.cfg.enc_res2net_scale
        cs = x.shape[1] // scale
        outs = [x[:, :cs]]
        for i in range(scale - 1):
            chunk = x[:, (i + 1) * cs:(i + 2) * cs]
            inp = chunk if i == 0 else chunk + outs[-1]
            outs.append(self._tdnn(f"{name}.blocks.{i}", inp, dilation))
        return torch.cat(outs, 1)

    def _se(self, name: str, x: torch.Tensor) -> torch.Tensor:
        """SqueezeExcitationBlock (speaker.rs:192-222): mean over T -> conv1 + ReLU -> conv2 + sigmoid -> scale."""
        s = x.mean(-1, keepdim=True)
        s = torch.relu(self._conv(name + ".conv1", s))
        s = sigmoid(self._conv(name + ".conv2", s))
        return x * s

    def _se_res2net(self, i: int, x: torch.Tensor) -> torch.Tensor:
        """SqueezeExcitationRes2NetBlock (speaker.rs:224-268)."""
        d = self.cfg.enc_dilations[i]
        out = self._tdnn(f"blocks.{i}.tdnn1", x)
        out = self._res2net(f"blocks.{i}.res2net_block", out, d)
        out = self._tdnn(f"blocks.{i}.tdnn2", out)
        out = self._se(f"blocks.{i}.se_block", out)
        return out + x

    def _asp(self, x: torch.Tensor) -> torch.Tensor:
        """AttentiveStatisticsPooling (speaker.rs:270-347) -> [B, 2C, 1]."""
        sd = (lambda v: torch.sqrt(v.clamp(1e-12))) if self.hf_rule else (lambda v: torch.sqrt(v + 1e-5))
        mean = x.mean(-1, keepdim=True)
        std = sd(((x - mean) ** 2).mean(-1, keepdim=True))
        attn_in = torch.cat([x, mean.expand_as(x), std.expand_as(x)], 1)
        attn = torch.tanh(self._tdnn("asp.tdnn", attn_in))
        attn = torch.softmax(self._conv("asp.conv", attn), -1)
        w_mean = (x * attn).sum(-1, keepdim=True)
        w_std = sd((((x - w_mean) ** 2) * attn).sum(-1, keepdim=True))
        return torch.cat([w_mean, w_std], 1)

    def forward(self, mel: torch.Tensor) -> torch.Tensor:
        """speaker.rs:448-476: initial TDNN, 3 SE-Res2Net blocks (outputs concatenated: MFA), MFA TDNN, ASP, FC.  No L2
        normalisation (speaker.rs:473-474)."""
        c = self.cfg
        h = self._tdnn("blocks.0", mel.to(torch.float32), c.enc_dilations[0])
        outs = []
        for i in range(1, 4):
            h = self._se_res2net(i, h)
            outs.append(h)
        h = self._tdnn("mfa", torch.cat(outs, 1), c.enc_dilations[4])
        return self._conv("fc", self._asp(h))[:, :, 0]
