import os, sys, time
sys.path.insert(0, "/root/repo")
from qwen3_tts_rs_b200 import api, spec as S, weights as W
spec = S.SPECS["1.7b"]
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), W.make_vocoder_weights(spec.vocoder))
for B in (1, 8):
    prompts = [W.synthetic_prompt(i, spec) for i in range(B)]
    pp = [tts.custom_voice_prompt(t, "ryan", "english") for t in prompts]
    for it in range(3):
        t0 = time.perf_counter()
        sess = api.Session(tts.model, B, api.SynthesisOptions(max_length=256), [42 + i for i in range(B)], max_seq=512)
        t1 = time.perf_counter()
        sess.prefill_ids([p[0] for p in pp], [p[1] for p in pp]); sess.set_trailing_ids([list(t[1:]) for t in prompts]); sess.synchronize()
        t2 = time.perf_counter()
        sess.generate(16); t3 = time.perf_counter()
        sess.vocode(16); t4 = time.perf_counter()
        sess.close(); t5 = time.perf_counter()
        print(f"B={B} it={it}: create {1e3*(t1-t0):.1f}  prefill+trailing {1e3*(t2-t1):.1f}  generate16 {1e3*(t3-t2):.1f}  vocode16 {1e3*(t4-t3):.1f}  close {1e3*(t5-t4):.1f} ms", flush=True)
