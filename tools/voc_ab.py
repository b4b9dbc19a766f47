"""A/B of the vocoder's convolution kernels (Q3_VOC_UMMA bit mask, vocoder.cu umma_mask): decodes the same codes with the
full-size Decoder12Hz, saves the PCM, compares it with a saved run and times the decode.
   Q3_VOC_UMMA=0 python tools/voc_ab.py save /tmp/base.npy ;  Q3_VOC_UMMA=15 python tools/voc_ab.py cmp /tmp/base.npy"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qwen3_tts_rs_b200 import api, spec as S, weights as W
mode, path = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
T = int(sys.argv[4]) if len(sys.argv) > 4 else 64
vs = S.VocoderSpec()
vw = W.make_vocoder_weights(vs)
m = api.Model(S.SPEC_TINY.__class__(**{**S.SPEC_TINY.to_dict(), "name": "tiny_fullvoc", "vocoder": vs}))
m.load(vw).finalize()
tts = api.Qwen3TTS(m)
g = torch.Generator().manual_seed(11)
codes = torch.randint(0, vs.codebook_size, (B, 16, T), generator=g, dtype=torch.int64).numpy()
pcm = tts.decode_tensor(codes)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): pcm = tts.decode_tensor(codes)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 3 * 1e3
pcm = np.asarray(pcm)
tag = f"Q3_VOC_UMMA={os.environ.get('Q3_VOC_UMMA', 'default')} B={B} T={T}: {ms:.2f} ms/decode"
if mode == "save":
    np.save(path, pcm); print(tag, "saved", pcm.shape, "rms", float(np.sqrt((pcm ** 2).mean())))
else:
    ref = np.load(path)
    d = pcm - ref
    print(tag, f"rms(ref) {np.sqrt((ref**2).mean()):.4f}  rms(diff) {np.sqrt((d**2).mean()):.3e}  max|diff| {np.abs(d).max():.3e}  nan {int(np.isnan(pcm).sum())}")
