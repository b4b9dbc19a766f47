"""ms per decode frame of the persistent kernel (CUDA events on the session stream), for quick A/B runs:
   Q3_MEGA=2 python tools/time_frames.py --model 1.7b --batch 1,8,16 --frames 128"""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qwen3_tts_rs_b200 import api, lib as L, spec as S, weights as W
ap = argparse.ArgumentParser(); ap.add_argument("--model", default="1.7b"); ap.add_argument("--batch", default="8")
ap.add_argument("--frames", type=int, default=128); ap.add_argument("--reps", type=int, default=2); a = ap.parse_args()
spec = S.SPECS[a.model]
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), None)
lib = L.load()
for B in [int(x) for x in a.batch.split(",")]:
    prompts = [W.synthetic_prompt(i, spec) for i in range(B)]
    pp = [tts.custom_voice_prompt(t, "ryan", "english") for t in prompts]
    seeds = [42 + i for i in range(B)]
    sess = api.Session(tts.model, B, api.SynthesisOptions(max_length=a.frames), seeds, max_seq=a.frames + 40)
    stream = torch.cuda.ExternalStream(lib.q3_session_stream(sess.handle))
    best = None
    for r in range(a.reps + 1):
        sess.reset(seeds)
        sess.prefill_ids([p[0] for p in pp], [p[1] for p in pp]); sess.set_trailing_ids([list(t[1:]) for t in prompts])
        sess.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); sess.generate_async(a.frames); e1.record(stream); sess.synchronize()
        ms = e0.elapsed_time(e1) / a.frames
        if r > 0: best = ms if best is None else min(best, ms)
    codes, n = sess.get_codes(a.frames)
    print(f"Q3_MEGA={os.environ.get('Q3_MEGA','default')} lib={'dev' if L.IS_DEV else 'product'} model={a.model} batch={B}: {best:.3f} ms/frame "
          f"({B / best * 1e3:.0f} frames/s), frames {int(n.max())}, checksum {int(codes.astype('int64').sum())}", flush=True)
    sess.close()
