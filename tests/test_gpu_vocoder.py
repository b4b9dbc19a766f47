"""GPU parity tests for the vocoder and the streaming session, through the C ABI.
Tolerance (north star): PCM within 1e-3 RMS of the oracle; asserted here as rms(err) <= 1e-3 and
max|err| <= 2e-2 on clamped [-1,1] PCM."""
import numpy as np
import pytest
import torch

from oracle import generate as OG
from oracle import vocoder as OV
from qwen3_tts_rs_b200 import api, spec as S, weights as W
from conftest import talker_weights, vocoder_weights
from helpers import gpu_tts, oracle_cfg, oracle_models

pytestmark = pytest.mark.gpu


def _rand_codes(B, T, seed):
    g = np.random.default_rng(seed)
    codes = g.integers(0, 2048, size=(B, 16, T), dtype=np.int64)
    codes[:, 0] = g.integers(0, 3072, size=(B, T))       # semantic ids may exceed 2048 (mod rule, decoder_12hz.rs:423)
    return codes


def _check(pcm, ref, what):
    err = pcm - ref
    rms = float(np.sqrt(np.mean(err ** 2)))
    print(what, "rms err", rms, "max err", float(np.abs(err).max()), "ref rms", float(np.sqrt(np.mean(ref ** 2))))
    assert rms <= 1e-3, (what, rms)
    assert float(np.abs(err).max()) <= 2e-2, what


@pytest.mark.parametrize("B,T", [(1, 1), (1, 2), (2, 7), (3, 33)])
def test_vocoder_tiny_vs_oracle(B, T):
    spec = S.SPEC_TINY
    vw = vocoder_weights(spec.vocoder, "tiny")
    tts = api.Qwen3TTS.from_weights(spec, {}, vw) if False else gpu_tts(spec, with_vocoder=True, vkey="tiny")
    codes = _rand_codes(B, T, 5 + T)
    pcm = tts.decode_tensor(codes)
    assert pcm.shape == (B, T * 1920)                     # reference_validation.rs:2266-2268
    ref = OV.Vocoder(spec.vocoder, vw).decode(codes)[:, 0].numpy()
    _check(pcm, ref, f"tiny B={B} T={T}")


@pytest.mark.parametrize("T", [2, 12])
def test_vocoder_full_size_vs_oracle(T):
    """Full Decoder12Hz dimensions (114 M parameters), short T so the CPU oracle finishes in seconds."""
    vs = S.VocoderSpec()
    vw = vocoder_weights(vs, "full")
    m = api.Model(S.SPEC_TINY.__class__(**{**S.SPEC_TINY.to_dict(), "name": "tiny_fullvoc", "vocoder": vs}))
    m.load(vw).finalize()
    tts = api.Qwen3TTS(m)
    codes = _rand_codes(2, T, 11)
    pcm = tts.decode_tensor(codes)
    assert pcm.shape == (2, T * 1920)
    assert np.abs(pcm).max() <= 1.0
    ref = OV.Vocoder(vs, vw).decode(codes)[:, 0].numpy()
    _check(pcm, ref, f"full T={T}")


def test_vocoder_is_causal_and_batch_independent():
    """Property at any size: decoding the first T1 frames alone equals the first T1*1920 samples of a longer
    decode (every op is causal), and rows do not influence each other."""
    spec = S.SPEC_TINY
    tts = gpu_tts(spec, with_vocoder=True, vkey="tiny")
    codes = _rand_codes(2, 40, 3)
    full = tts.decode_tensor(codes)
    part = tts.decode_tensor(codes[:, :, :13].copy())
    assert np.array_equal(part, full[:, : 13 * 1920])
    solo = tts.decode_tensor(codes[1:2].copy())
    assert np.array_equal(solo[0], full[1])


def test_decode_codes_empty_and_layout():
    spec = S.SPEC_TINY
    tts = gpu_tts(spec, with_vocoder=True, vkey="tiny")
    assert api.codes_to_tensor([]).shape == (1, 16, 0)            # lib.rs:2016-2021
    t = api.codes_to_tensor([list(range(16)), list(range(100, 116))])
    assert t.shape == (1, 16, 2) and t[0, 0].tolist() == [0, 100] and t[0, 1].tolist() == [1, 101]   # lib.rs:2031-2050
    audio = tts.decode_codes([[5] * 16])
    assert len(audio) == 1920 and audio.sample_rate == 24000


def test_out_of_range_codes_are_an_error():
    """decoder_12hz.rs:429, 443: index_select fails on an acoustic code >= codebook_size or a negative code; semantic codes
    are reduced modulo the codebook size (decoder_12hz.rs:423-427), so 3071 is fine there."""
    spec = S.SPEC_TINY
    tts = gpu_tts(spec, with_vocoder=True, vkey="tiny")
    codes = _rand_codes(1, 4, 1)
    codes[0, 0, 2] = 3071
    assert tts.decode_tensor(codes).shape == (1, 4 * 1920)
    for q, f, v in ((3, 1, 2048), (15, 0, -1), (0, 3, -5)):
        bad = codes.copy()
        bad[0, q, f] = v
        with pytest.raises(api.L.Q3Error) as e:
            tts.decode_tensor(bad)
        assert e.value.status == "Q3_ERR_INVALID" and "out of range" in str(e.value)


def test_streaming_session_matches_oracle_chunks():
    """StreamingSession (lib.rs:1650-1759): chunk_frames = 4, 10 frames -> chunks of 4,4,2 frames, each vocoded
    independently.  Held unconditionally: the streamed codes equal the non-streamed run's codes value for value; every
    chunk's PCM equals (rms <= 1e-3) the oracle vocoder's decode of THAT chunk's codes alone (no state crosses chunks,
    lib.rs:1755-1758); chunk sizes equal the oracle session's; total samples == frames * 1920 (streaming_e2e.rs:150-157)."""
    spec = S.SPEC_TINY
    vw = vocoder_weights(spec.vocoder, "tiny")
    tts = gpu_tts(spec, with_vocoder=True, vkey="tiny")
    ids = W.synthetic_prompt(2, spec)
    opts = api.SynthesisOptions(max_length=10, chunk_frames=4, seed=99)
    tk, cp = oracle_models(spec)
    voc = OV.Vocoder(spec.vocoder, vw)
    emb = tk.custom_voice_embeds(ids, S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"])
    osess = OG.StreamingSession(tk, cp, lambda c: voc.decode(c)[0, 0].numpy(), emb, ids, oracle_cfg(opts), 99, chunk_frames=4)
    ochunks = list(osess)
    # the API object (chunk sizes, done flag, frame count) ...
    sess = tts.synthesize_streaming(ids, options=opts)
    chunks = list(sess)
    assert sess.is_done()
    assert [len(c) for c in chunks] == [len(c) for c in ochunks] == [4 * 1920, 4 * 1920, 2 * 1920]
    assert sum(len(c) for c in chunks) == sess.frames_generated() * 1920
    # ... and the same stream through the session call that also returns each chunk's codes
    prompts = [tts.custom_voice_prompt(ids, "ryan", "english")]
    low = tts._new_session([ids], prompts, opts, [99])
    streamed, pcm_chunks = [], []
    while True:
        codes, pcm, n, done = low.stream_next()
        if n[0]:
            streamed.append(codes[0, : n[0]].tolist())
            pcm_chunks.append(pcm[0, : n[0] * 1920].copy())
        if done:
            break
    low.close()
    nonstream = tts.generate_codes([ids], options=opts, seeds=[99])[0]
    assert [f for ch in streamed for f in ch] == nonstream                 # streamed codes == non-streamed codes
    assert len(pcm_chunks) == len(chunks)
    for i, (ch, p, c) in enumerate(zip(streamed, pcm_chunks, chunks)):
        assert np.array_equal(p, c.samples)                                # both stream calls give the same samples
        _check(p, voc.decode(OG.codes_to_tensor(ch))[0, 0].numpy(), f"chunk {i}")


def test_synthesize_with_voice_end_to_end():
    """synthesize_with_timing shape contract: audio length == frames * 1920, timing fields populated."""
    spec = S.SPEC_TINY
    tts = gpu_tts(spec, with_vocoder=True, vkey="tiny")
    prompts = [W.synthetic_prompt(i, spec) for i in range(2)]
    audio, timing = tts.synthesize_with_voice(prompts, options=api.SynthesisOptions(max_length=8), seeds=[1, 2], with_timing=True)
    assert all(len(a) == 8 * 1920 for a in audio)
    assert timing.generation_frames == 8 and timing.generation_ms > 0 and timing.decode_ms > 0


def test_synthesize_voice_clone_budget_and_cut_rule():
    """synthesize_voice_clone (lib.rs:895-1060) through the mirror: ICL rows get the reference's frame budget
    min(max_length, max(75, 6 x text tokens)) and repetition penalty >= 1.5 (lib.rs:913-927); the reference frames are decoded
    in front of the generated ones and ref_len / total_len of the waveform is cut from its start (lib.rs:1021-1040); an
    x-vector-only batch decodes the generated frames alone; a mixed batch is refused."""
    spec = S.SPEC_TINY
    tts = gpu_tts(spec, with_vocoder=True, vkey="tiny")
    g = torch.Generator().manual_seed(8)
    spk = lambda: torch.randn(spec.hidden, generator=g) * 0.05
    ref = lambda t: torch.randint(0, 2048, (t, 16), generator=g).numpy().astype(np.uint32)
    texts = [W.synthetic_prompt(0, spec)[:4], W.synthetic_prompt(1, spec)[:15]]
    icl = [api.VoiceClonePrompt(spk(), ref(5), [3, 4, 5]), api.VoiceClonePrompt(spk(), ref(9), [6, 7])]
    opts = api.SynthesisOptions(max_length=80, eos_token_id=None)
    codes = tts.generate_codes_voice_clone(texts, icl, options=opts, seeds=[1, 2])
    assert [len(c) for c in codes] == [75, 80]           # max(75, 6*4) = 75 ; min(80, max(75, 6*15)) = 80
    audio = tts.synthesize_voice_clone(texts, icl, options=opts, seeds=[1, 2])
    for a, c, p in zip(audio, codes, icl):
        total = (len(p.ref_codes) + len(c)) * 1920
        assert len(a) == total - len(p.ref_codes) * total // (len(p.ref_codes) + len(c))
    # row 0 alone (batch 1, its own 75-frame budget) gives the same codes: rows are independent
    assert tts.generate_codes_voice_clone(texts[:1], icl[:1], options=opts, seeds=[1])[0] == codes[0]
    xv = [api.VoiceClonePrompt(spk()), api.VoiceClonePrompt(spk())]
    audio = tts.synthesize_voice_clone(texts, xv, options=api.SynthesisOptions(max_length=6, eos_token_id=None), seeds=[1, 2])
    assert all(len(a) == 6 * 1920 for a in audio)
    with pytest.raises(ValueError):
        tts.generate_codes_voice_clone(texts, [icl[0], xv[0]], options=opts)


def test_concurrent_sessions_on_one_model_from_threads():
    """a finalized model is read-only and shareable across threads, one session per request.  Three threads
    synthesize different utterances at the same time on one model (ctypes drops the GIL inside the C ABI, so the persistent
    cooperative kernels, the vocoder kernels and the session pools of the three requests really interleave); every result must
    equal the same request run alone, bit for bit."""
    import threading
    spec = S.SPEC_TINY
    tts = gpu_tts(spec, with_vocoder=True, vkey="tiny")
    reqs = [([W.synthetic_prompt(3 * i + j, spec) for j in range(1 + i)], [100 + 3 * i + j for j in range(1 + i)]) for i in range(3)]
    opts = api.SynthesisOptions(max_length=24)
    alone = [tts.synthesize_with_voice(p, options=opts, seeds=s) for p, s in reqs]
    for rounds in range(3):
        out, errs = [None] * 3, []
        def work(i):
            try:
                out[i] = tts.synthesize_with_voice(reqs[i][0], options=opts, seeds=reqs[i][1])
            except Exception as e:      # noqa: BLE001 -- reported below
                errs.append((i, repr(e)))
        th = [threading.Thread(target=work, args=(i,)) for i in range(3)]
        for t in th: t.start()
        for t in th: t.join()
        assert not errs, errs
        for i in range(3):
            assert len(out[i]) == len(alone[i])
            for a, b in zip(out[i], alone[i]):
                assert np.array_equal(a.samples, b.samples), ("request differs when run concurrently", i, rounds)
