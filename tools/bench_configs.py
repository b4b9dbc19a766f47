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
# stateful, first chunk 2 frames then chunks of 10 (q3_session_set_first_chunk): the TTFA of the 2-frame stream at the throughput of 10-frame chunks
opts3 = api.SynthesisOptions(max_length=120, seed=42, chunk_frames=10, stream_left_context=-1, stream_first_chunk=2)
for rep in range(3):
    t0 = time.perf_counter()
    st = tts.synthesize_voice_design_streaming(text, instr, options=opts3)
    first = None; n = 0
    for chunk in st:
        if first is None: first = time.perf_counter() - t0
        n += len(chunk.samples)
    dt = time.perf_counter() - t0
line("configs[3] stateful streaming, first chunk 2 then 10 frames", n // 1920, dt, f" TTFA {first*1e3:.1f} ms")
for b in (4, 32):
    nf, dt = run_batch(tts, prompts[:b], 128)
    line(f"configs[4] 1.7B CustomVoice, {b} utterances on this GPU (8-GPU share = 4)", nf, dt)
# SURVEY 8(f) row 4 (not a BASELINE configuration): the voice-clone front end on the 0.6B Base dimensions (speaker embedding
# width = talker hidden = 1024): ECAPA speaker encoder on a 3 s mel spectrogram (24 kHz, hop 256 -> 282 frames), then an ICL
# voice-clone request (60 reference frames + transcript) through synthesize_voice_clone
del tts
w = dict(W.make_talker_weights(spec06)); vw = dict(W.make_vocoder_weights(spec06.vocoder)); vw.update(W.make_speaker_weights(S.SpeakerSpec()))
tts = api.Qwen3TTS.from_weights(spec06, w, vw)
mel = (np.random.default_rng(0).standard_normal((1, 128, 282)) * 2 - 3).astype(np.float32)
tts.speaker_encode(mel)
t0 = time.perf_counter()
for _ in range(5): emb = tts.speaker_encode(mel)
print(f"{'f4 speaker encoder (ECAPA-TDNN), 3 s of audio (282 mel frames)':62s} wall {(time.perf_counter() - t0) / 5 * 1e3:9.2f} ms per utterance", flush=True)
ref = np.random.default_rng(1).integers(0, 2048, size=(60, 16)).astype(np.uint32)
vc = api.VoiceClonePrompt(emb[0], ref, ids(9, 20, spec06))
o = api.SynthesisOptions(max_length=128, eos_token_id=None)
tts.synthesize_voice_clone([ids(44, 30, spec06)], [vc], options=o, seeds=[42])
t0 = time.perf_counter()
for _ in range(3): audio = tts.synthesize_voice_clone([ids(44, 30, spec06)], [vc], options=o, seeds=[42])
dt = (time.perf_counter() - t0) / 3
line("f4 0.6B voice clone (ICL: 60 reference frames), batch 1", len(audio[0].samples) // 1920, dt)
