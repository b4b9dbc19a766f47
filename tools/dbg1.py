import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, torch
from qwen3_tts_rs_b200 import api, spec as S, weights as W
from helpers import gpu_tts, oracle_run, bf16_ulp_diff
for spec in (S.SPEC_TINY, S.SPEC_MID):
    opts = api.SynthesisOptions(max_length=6, seed=42)
    text_ids = W.synthetic_prompt(3, spec)
    frames, tr, emb = oracle_run(spec, text_ids, 42, opts, trace=True)
    tts = gpu_tts(spec)
    sess = api.Session(tts.model, 1, opts, [42], max_seq=64)
    sess.prefill_embeds([emb[0]])
    for fr in tr.frames[:4]:
        codes, lg = sess.code_predictor_frame(fr["cp_in_hidden"][0, 0], [fr["tok"]], want_logits=True)
        ol = fr["cp_logits"].float().numpy()
        d = np.abs(lg[0] - ol)
        print(spec.name, 'frame', fr['frame'], 'cp logits maxdiff per group', np.round(d.max(axis=1), 4), 'scale', np.abs(ol).max())
        print(' codes gpu', codes[0].tolist()); print(' codes ora', fr['codes'])
        hid, logits = sess.talker_step(fr["step_input"][0, 0])
        dh = (hid[0].float() - fr["hidden"][0, 0]).abs()
        print(' hidden maxdiff', float(dh.max()), 'rms', float(fr["hidden"].pow(2).mean().sqrt()), 'ulpdiff max', float(bf16_ulp_diff(hid[0], fr["hidden"][0,0]).max()))
        dl = np.abs(logits[0] - fr["logits"][0]); print(' logits maxdiff', dl.max(), 'scale', np.abs(fr["logits"][0]).max())
    sess.close()
