"""Where does the wall time of one bench step go?  Host-timed segments with a stream sync after each."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_tts_rs_b200 import api, lib as L, spec as S, weights as W
spec = S.SPECS["1.7b"]; B, F = 8, 256
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), W.make_vocoder_weights(spec.vocoder))
prompts = [W.synthetic_prompt(i, spec) for i in range(B)]
seeds = [42 + i for i in range(B)]
pp = [tts.custom_voice_prompt(t, "ryan", "english") for t in prompts]
lmax = max(len(p[0]) for p in pp)
sess = api.Session(tts.model, B, api.SynthesisOptions(max_length=F), seeds, max_seq=lmax + F + 8)
trailing = [list(t[1:]) for t in prompts]
def seg(f):
    t0 = time.perf_counter(); f(); sess.synchronize(); return (time.perf_counter() - t0) * 1e3
for it in range(4):
    r = seg(lambda: sess.reset(seeds))
    p = seg(lambda: sess.prefill_ids([q[0] for q in pp], [q[1] for q in pp]))
    t = seg(lambda: sess.set_trailing_ids(trailing))
    g = seg(lambda: sess.generate_async(F))
    v = seg(lambda: sess.vocode(F, to_host=False))
    tm = sess.timing()
    print(f"iter {it}: reset {r:.2f}  prefill_ids {p:.2f} (device {tm.prefill_ms:.2f})  trailing {t:.2f}  generate {g:.2f}  vocode {v:.2f} (device {tm.decode_ms:.2f})  sum {r+p+t+g+v:.2f}")
t0 = time.perf_counter()
for it in range(3):
    sess.reset(seeds); sess.prefill_ids([q[0] for q in pp], [q[1] for q in pp]); sess.set_trailing_ids(trailing)
    sess.generate_async(F); sess.vocode(F, to_host=False)
sess.synchronize()
print("back to back per step", (time.perf_counter() - t0) * 1e3 / 3)
