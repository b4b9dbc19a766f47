"""All five BASELINE.json configurations on one GPU (profiles/r1_baseline_configs.txt).  Wall-clock, host-buffer API
(what a caller of the session API sees): RTF = wall / audio seconds and frames/s as benches/e2e_bench.rs:334-344;
TTFA = time from the request (prompt ids on the host) to the first streamed chunk on the host, e2e_bench.rs:226-230."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qwen3_tts_rs_b200 import api, spec as S, weights as W

def load(spec):
    return api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), W.make_vocoder_weights(spec.vocoder))

def ids(seed, n, spec):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 151643, (n,), generator=g).tolist()

def run_batch(tts, prompts, frames, reps=3):
    opts = api.SynthesisOptions(max_length=frames)
    seeds = [42 + i for i in range(len(prompts))]
    tts.synthesize_with_voice(prompts, options=opts, seeds=seeds)            # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        audio = tts.synthesize_with_voice(prompts, options=opts, seeds=seeds)
    dt = (time.perf_counter() - t0) / reps
    nf = sum(len(a.samples) for a in audio) // 1920
    return nf, dt

def line(name, nf, dt, extra=""):
    print(f"{name:62s} frames {nf:6d}  wall {dt*1e3:9.1f} ms  frames/s {nf/dt:8.1f}  RTF {dt/(nf*0.08):.5f} {extra}", flush=True)

spec06, spec17 = S.SPECS["0.6b"], S.SPECS["1.7b"]
print("configs[0] (0.6B, CPU F32, seed 42) is the oracle-only case: see bench.py --impl reference / cpu_baseline", flush=True)
tts = load(spec06)
for nm, n in (("short (14 ids)", 14), ("long (140 ids)", 140)):
    nf, dt = run_batch(tts, [ids(42, n, spec06)], 128)
    line(f"configs[1] 0.6B Base-size, batch 1, {nm}, 128 frames", nf, dt)
del tts
tts = load(spec17)
prompts = [W.synthetic_prompt(i, spec17) for i in range(32)]
nf, dt = run_batch(tts, prompts[:8], 256)
line("configs[2] 1.7B CustomVoice(ryan), batch 8, 256 frames", nf, dt)
# configs[3]: VoiceDesign, batch 1, streaming
instr = ids(7, 24, spec17)
text = ids(43, 40, spec17)
opts = api.SynthesisOptions(max_length=120, seed=42)
for rep in range(3):
    t0 = time.perf_counter()
    st = tts.synthesize_voice_design_streaming(text, instr, options=opts)
    first = None; n = 0
    for chunk in st:
        if first is None: first = time.perf_counter() - t0
        n += len(chunk.samples)
    dt = time.perf_counter() - t0
line("configs[3] 1.7B VoiceDesign, batch 1, streaming (chunk 10 frames)", n // 1920, dt, f" TTFA {first*1e3:.1f} ms")
# the same stream, STATEFUL (carried vocoder state): 2-frame chunks, PCM identical to the non-streamed decode
opts2 = api.SynthesisOptions(max_length=120, seed=42, chunk_frames=2, stream_left_context=-1)
for rep in range(3):
    t0 = time.perf_counter()
    st = tts.synthesize_voice_design_streaming(text, instr, options=opts2)
    first = None; n = 0
    for chunk in st:
        if first is None: first = time.perf_counter() - t0
        n += len(chunk.samples)
    dt = time.perf_counter() - t0
line("configs[3] stateful streaming, chunk 2 frames (exact PCM)", n // 1920, dt, f" TTFA {first*1e3:.1f} ms")
for b in (4, 32):
    nf, dt = run_batch(tts, prompts[:b], 128)
    line(f"configs[4] 1.7B CustomVoice, {b} utterances on this GPU (8-GPU share = 4)", nf, dt)
