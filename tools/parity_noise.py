"""Noise calibration of the follow-mode parity bar (tests/test_gpu_parity.py): for one model size, how far is the CUDA path
from the oracle's bf16 mode, and how far is the oracle's bf16 mode from the oracle's f32 mode (the rounding noise floor of
the reference's own CUDA-path arithmetic)?  Run on a GPU box:  python tools/parity_noise.py 1.7b 8 3
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import generate as OG, model as OM   # noqa: E402
from qwen3_tts_rs_b200 import api, spec as S, weights as W   # noqa: E402
from conftest import talker_weights   # noqa: E402
from helpers import oracle_cfg   # noqa: E402
from test_gpu_parity import run_tapped   # noqa: E402


def stats(name, gpu, bf, f32):
    gpu, bf, f32 = [torch.as_tensor(np.asarray(x, dtype=np.float32)).flatten() for x in (gpu, bf, f32)]
    rms = float(f32.pow(2).mean().sqrt())
    r = lambda x: float(x.pow(2).mean().sqrt()) / rms
    m = lambda x: float(x.abs().max()) / rms
    print(f"{name:28s} rms {rms:8.4f} | gpu-bf16 rms {r(gpu - bf):.5f} max {m(gpu - bf):.4f} | bf16-f32 rms {r(bf - f32):.5f} max {m(bf - f32):.4f}"
          f" | gpu-f32 rms {r(gpu - f32):.5f} max {m(gpu - f32):.4f}")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "mid"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    F = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    rows = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
    spec = S.SPECS[name]
    w = talker_weights(spec)
    opts = api.SynthesisOptions(max_length=F)
    prompts = [W.synthetic_prompt(i, spec) for i in range(B)]
    seeds = [42 + i for i in range(B)]
    tts = api.Qwen3TTS.from_weights(spec, w)
    tapped, taps = run_tapped(tts, prompts, seeds, opts, F)
    cfg = oracle_cfg(opts)
    for b in rows:
        got = tapped[b]
        n = len(got)
        out = {}
        for mode, prec in (("bf16", OM.BF16P), ("f32", OM.F32P)):
            tk, cp = OM.Talker(spec, w, prec), OM.CodePredictor(spec, w, prec)
            emb = tk.custom_voice_embeds(prompts[b], S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"])
            out[mode] = OG.follow(tk, cp, emb, prompts[b], cfg, seeds[b], got, first_logits=taps["first_logits"][b],
                                  frame_logits=[taps["logits"][f, b] for f in range(n)], kv_max=F + 64)
        print(f"== {name} batch {B} row {b}: {n} frames")
        stats("prefill logits", taps["first_logits"][b], out["bf16"]["prefill_logits"], out["f32"]["prefill_logits"])
        for f in range(n):
            ob, of = out["bf16"]["frames"][f], out["f32"]["frames"][f]
            stats(f"frame {f} cp logits (all)", taps["cp_logits"][f, :, b], ob["cp_logits"].float().numpy(), of["cp_logits"].float().numpy())
            for g in (0, 7, 14):
                stats(f"frame {f} cp logits pass {g}", taps["cp_logits"][f, g, b], ob["cp_logits"][g].float().numpy(), of["cp_logits"][g].float().numpy())
            stats(f"frame {f} talker logits", taps["logits"][f, b], ob["logits"], of["logits"])
            same = sum(int(a == c) for a, c in zip(ob["own_codes"], got[f][1:]))
            print(f"   bf16-oracle arg-max == emitted code: {same}/15; f32-oracle: {sum(int(a == c) for a, c in zip(of['own_codes'], got[f][1:]))}/15")


if __name__ == "__main__":
    main()
