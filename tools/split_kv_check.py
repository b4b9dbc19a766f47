"""Split-KV attention checks (m2_attn_units): (1) with the unit-mapped path on but no row long enough to split
(Q3_SPLIT_KV=100000) the codes equal the unsplit kernel's (Q3_SPLIT_KV=0) -- run as two processes and compare checksums;
(2) long-context timing.   python tools/split_kv_check.py <batch> <frames> <instruct_len>"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qwen3_tts_rs_b200 import api, spec as S, weights as W
B = int(sys.argv[1]); F = int(sys.argv[2]); NI = int(sys.argv[3])
spec = S.SPECS[sys.argv[4] if len(sys.argv) > 4 else "1.7b"]
tts = api.Qwen3TTS.from_weights(spec, W.make_talker_weights(spec), None)
g = torch.Generator().manual_seed(5)
prompts = [W.synthetic_prompt(i, spec) for i in range(B)]
instr = [torch.randint(0, 150000, (NI if i % 2 == 0 else 7,), generator=g).tolist() for i in range(B)]
pp = [tts.voice_design_prompt(t, ins, "english") for t, ins in zip(prompts, instr)]
opts = api.SynthesisOptions(max_length=F, eos_token_id=None)
def run(rows):
    sess = api.Session(tts.model, len(rows), opts, [42 + r for r in rows], max_seq=NI + F + 64)
    sess.prefill_ids([pp[r][0] for r in rows], [pp[r][1] for r in rows]); sess.set_trailing_ids([list(prompts[r][1:]) for r in rows])
    sess.synchronize(); t0 = time.perf_counter()
    codes, n = sess.generate(F); dt = time.perf_counter() - t0
    sess.close()
    return codes, dt
codes, dt = run(list(range(B)))
codes2, dt = run(list(range(B)))
print(f"Q3_SPLIT_KV={os.environ.get('Q3_SPLIT_KV', 'default')} B={B} F={F} instruct {NI}: {dt / F * 1e3:.3f} ms/frame, checksum {int(codes.astype('int64').sum())}, repeat identical {np.array_equal(codes, codes2)}")
# row independence: row 0 (long) and row 1 (short) alone
for r in range(min(B, 2)):
    c1, _ = run([r])
    print(f"   row {r} alone == row {r} of the batch: {np.array_equal(c1[0], codes[r])}")
