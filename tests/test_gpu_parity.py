"""GPU parity of the FREE-RUNNING production loop (q3_generate) at every frame, with no token exemptions.

A free-running CUDA run and a free-running oracle run part ways at the first near-tie (bf16 logit noise), after which
nothing can be compared directly.  These tests close that hole from both sides ("parity holes"):

  * the CUDA run is TAPPED (q3_debug_generate_tapped: same kernels and phase program, frames run one per launch so the
    taps can be read back; the codes are asserted bit-identical to the untapped q3_generate run): per frame the raw f32
    talker logits the sampler drew from, the code-predictor logits of all 15 passes, the PCG state and the talker input;
  * the oracle FOLLOWS the emitted codes (oracle/generate.py follow()): every decision input is the CUDA path's, every
    tensor is the oracle's own.

Held at every frame of every followed row:
  1. RNG stream: the PCG state before every draw equals the oracle's (sampling.rs:84-94) -- bit-exact.
  2. sampled token == the reference sampler (penalties, suppression, min_new_tokens, top-k, top-p, multinomial;
     sampling.rs:140-319, lib.rs:1271-1322) run by the oracle on the CUDA path's own logits with the oracle's RNG and
     penalty mask.  Only a draw within 2e-6 of a CDF boundary is exempt (CUDA expf vs libm; same bar as the q3_sample
     tests); exemptions are counted and must be zero in practice.
  3. acoustic codes == arg-max (lowest index among ties) of the CUDA path's own code-predictor logits -- bit-exact.
  4. talker input == bf16(bf16(sem + sum_i E_i[c_i]) + trailing_text_row_or_tts_pad) computed by the oracle from the
     emitted codes (lib.rs:612-622, code_predictor.rs:497-519) -- bit-exact (the trailing-text rule: row
     frame_idx while frame_idx < trailing length, tts_pad after).  The text rows are the ones the device projected
     (q3_debug_get_trailing), themselves held to the oracle's projection by the bar of item 5, so that a one-ulp
     difference of the projection GEMM does not mask -- or fake -- an error of the add.
  5. code-predictor logits, talker logits and prefill logits against the oracle with a NOISE-CALIBRATED bar: the oracle
     follows the same codes twice, in bf16 mode (the reference's CUDA-path arithmetic) and in f32 mode, and for every
     tensor   rms(cuda - f32) <= 1.5 rms(bf16 - f32) + 1e-3 rms   and   |cuda - f32| <= 8 rms(bf16 - f32) + 1e-2 rms + 2^-6 |f32|,
     i.e. the CUDA path must be as close to exact arithmetic as the reference's own bf16 rounding is (measured on B200,
     tools/parity_noise.py: 1.02-1.08 at 1.7B where a fixed 2^-4 rms element bar fails on 0.1 % of the logits after 28
     layers).  A wrong operand, position or rounding point shows up at the scale of rms itself, ~50x this bar.
  6. where the oracle's own arg-max differs from the emitted code, the oracle's top-2 margin is below 2^-5 |top1|.
"""
import numpy as np
import pytest
import torch

from oracle import generate as OG
from qwen3_tts_rs_b200 import api, spec as S, weights as W
from helpers import gpu_tts, oracle_cfg, oracle_models

pytestmark = pytest.mark.gpu


def close_to_f32(cuda, bf, f32, what, stats):
    """Item 5 of the module docstring for one tensor; `stats` collects the worst ratios for the report."""
    cuda, bf, f32 = [torch.as_tensor(np.asarray(x, dtype=np.float32)).flatten() for x in (cuda, bf, f32)]
    rms = float(f32.pow(2).mean().sqrt())
    sigma = float((bf - f32).pow(2).mean().sqrt())
    e_rms = float((cuda - f32).pow(2).mean().sqrt())
    e_max = float((cuda - f32).abs().max())
    stats["rms_ratio"] = max(stats.get("rms_ratio", 0.0), e_rms / max(sigma, 1e-12))
    stats["max_over_sigma"] = max(stats.get("max_over_sigma", 0.0), e_max / max(sigma, 1e-12))
    assert e_rms <= 1.5 * sigma + 1e-3 * rms, f"{what}: rms(cuda - f32) {e_rms:.4g} vs rms(bf16 - f32) {sigma:.4g} (tensor rms {rms:.4g})"
    # element-wise: 8 sigma + 1 % of rms + 4 bf16 ulp of the element itself (a logit of magnitude 60 has a bf16 ulp of 0.25)
    over = (cuda - f32).abs() - (8.0 * sigma + 1e-2 * rms + 2.0 ** -6 * f32.abs())
    assert float(over.max()) <= 0.0, (f"{what}: {int((over > 0).sum())}/{over.numel()} elements outside the bar, max|cuda - f32| {e_max:.4g} vs "
                                      f"rms(bf16 - f32) {sigma:.4g} (tensor rms {rms:.4g})")


def run_tapped(tts, prompts, seeds, opts, frames, instruct=None, clone=None):
    if clone is not None:     # voice clone (talker.rs:511-564, 646-705): speaker embedding, optional ICL reference
        pp = [tts.voice_clone_prompt(t, c, "english") for t, c in zip(prompts, clone)]
        sess = api.Session(tts.model, len(prompts), opts, seeds, max_seq=max(len(p[0]) for p in pp) + frames + 40)
        sess.prefill_voice_clone([p[0] for p in pp], [p[1] for p in pp], [c.speaker_embedding for c in clone],
                                 [c.ref_codes if c.is_icl else None for c in clone])
        sess.set_trailing_ids([p[2] for p in pp])
    else:
        if instruct is None:
            pp = [tts.custom_voice_prompt(t, "ryan", "english") for t in prompts]
        else:       # VoiceDesign prefill (talker.rs:585-627): the instruct ids sit in front of the role prefix
            pp = [tts.voice_design_prompt(t, ins, "english") for t, ins in zip(prompts, instruct)]
        sess = tts._new_session(prompts, pp, opts, seeds, max_seq=max(len(p[0]) for p in pp) + frames + 40)
    try:
        text = sess.trailing_rows(cap=max([len(t) for t in prompts] + [len(p[2] or []) for p in pp if len(p) > 2]) + 1)
        codes, n, taps = sess.generate_tapped(frames)
        taps["text"] = text
    finally:
        sess.close()
    return [codes[b, : n[b]].tolist() for b in range(len(prompts))], taps


def check_follow(spec, tts, prompts, seeds, opts, frames, rows, tapped=None, taps=None, models=None, models32=None, instruct=None,
                 clone=None):
    """Runs the tapped CUDA loop (unless given) and holds rows `rows` to the oracle as the module docstring says.
    Returns a report dict (counts only; every violation asserts)."""
    if tapped is None:
        tapped, taps = run_tapped(tts, prompts, seeds, opts, frames, instruct, clone)
    tk, cp = models if models is not None else oracle_models(spec)
    tk32, cp32 = models32 if models32 is not None else oracle_models(spec, bf16=False)
    cfg = oracle_cfg(opts)
    rep = dict(rows=len(rows), frames=0, sampled=0, sample_exempt=0, codes=0, oracle_argmax_differs=0)
    for b in rows:
        got = tapped[b]
        n = len(got)
        def prefill_embeds(t, c):
            """-> (prefill embeddings, trailing-text override or None) of oracle talker t / code predictor c"""
            if clone is not None:
                from oracle import model as OM
                pr = clone[b]
                emb_, tr_ = OM.voice_clone_prompt(t, c, prompts[b], pr.speaker_embedding, S.LANGUAGE_IDS["english"],
                                                  pr.ref_codes if pr.is_icl else None, pr.ref_text_ids if pr.is_icl else None)
                return emb_, (tr_ if pr.is_icl else None)
            if instruct is None:
                return t.custom_voice_embeds(prompts[b], S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"]), None
            return t.voice_design_embeds(prompts[b], instruct[b], S.LANGUAGE_IDS["english"]), None
        emb, tr_over = prefill_embeds(tk, cp)
        kv_max = int(emb.shape[1]) + frames + 64
        tr_rows, tr_lt, tr_pad = taps["text"]
        gpu_text = (tr_rows[b, : int(tr_lt[b])][None], tr_pad[None, None])
        if int(tr_lt[b]) == 0:      # no trailing rows on the device (ICL prompt consumed the text) == the reference's [tts_pad]
            gpu_text = (tr_pad[None, None], tr_pad[None, None])
        fo = OG.follow(tk, cp, emb, prompts[b], cfg, seeds[b], got, first_logits=taps["first_logits"][b],
                       frame_logits=[taps["logits"][f, b] for f in range(n)], kv_max=kv_max, text_rows=gpu_text,
                       trailing_override=tr_over)
        # the device's text projection (trailing rows + tts_pad) vs the oracle's: the same noise-calibrated bar as the logits
        o_tr = torch.cat([fo["text"]["trailing"][0], fo["text"]["pad"][0]], 0)
        g_tr = torch.cat([gpu_text[0][0].float(), gpu_text[1][0].float()], 0)
        emb32, tr_over32 = prefill_embeds(tk32, cp32)
        f_tr32, f_len, f_pad32 = tk32.build_trailing_text(prompts[b])
        if tr_over32 is not None:
            f_tr32 = tr_over32
        close_to_f32(g_tr, o_tr, torch.cat([f_tr32[0], f_pad32[0]], 0), f"trailing text rows row {b}", rep)
        f32 = OG.follow(tk32, cp32, emb32, prompts[b], cfg, seeds[b], got, kv_max=kv_max, trailing_override=tr_over32)      # noise calibration (item 5)
        # 1. RNG stream
        assert [int(x) for x in taps["rng"][: n + 1, b]] == fo["rng_states"], ("rng stream", b)
        # 2. sampler replay on the CUDA path's own logits
        want = [fr[0] for fr in got]
        for f in range(n + 1):
            if f < n:
                expect = want[f]
            elif n < frames and opts.eos_token_id is not None:
                expect = opts.eos_token_id          # the row stopped early: the token after its last frame is EOS
            else:
                break                               # the token drawn after the last requested frame is not emitted
            rep["sampled"] += 1
            if fo["replayed"][f] != expect:
                assert fo["margins"][f] is not None and fo["margins"][f] <= 2e-6, ("sampled token", b, f, fo["replayed"][f], expect, fo["margins"][f])
                rep["sample_exempt"] += 1
        close_to_f32(taps["first_logits"][b], fo["prefill_logits"], f32["prefill_logits"], f"prefill logits row {b}", rep)
        for f in range(n):
            o = fo["frames"][f]
            # 3. greedy codes are the arg-max of the path's own logits
            for g in range(15):
                rep["codes"] += 1
                assert int(np.argmax(taps["cp_logits"][f, g, b])) == got[f][1 + g], ("code != argmax of own logits", b, f, g)
            # 4. talker input, bit-exact
            assert torch.equal(taps["step_input"][f, b].view(torch.int16),
                               o["step_input"][0, 0].to(torch.bfloat16).view(torch.int16)), ("step_input", b, f)
            # 5. tensors vs the oracle
            o32 = f32["frames"][f]
            close_to_f32(taps["cp_logits"][f, :, b], o["cp_logits"].float().numpy(), o32["cp_logits"].float().numpy(),
                         f"cp logits row {b} frame {f}", rep)
            close_to_f32(taps["logits"][f, b], o["logits"], o32["logits"], f"talker logits row {b} frame {f}", rep)
            # 6. the oracle's own arg-max
            for g in range(15):
                if o["own_codes"][g] != got[f][1 + g]:
                    top2 = torch.topk(o["cp_logits"][g].float(), 2).values
                    margin = float(top2[0] - top2[1])
                    assert margin <= 2.0 ** -5 * abs(float(top2[0])) + 1e-6, ("oracle arg-max differs away from a tie", b, f, g, margin)
                    rep["oracle_argmax_differs"] += 1
        rep["frames"] += n
    return rep


@pytest.mark.parametrize("spec,mega", [(S.SPEC_TINY, None), (S.SPEC_TINY_PROJ, None), (S.SPEC_MID, None), (S.SPEC_RING, None),
                                       (S.SPEC_RING, "4"), (S.SPEC_RING, "5"), (S.SPEC_MID, "5")],
                         ids=["tiny", "tiny_proj", "mid", "ring", "ring-mega4", "ring-mega5", "mid-mega5"])
def test_free_running_generate_follows_the_oracle_every_frame(spec, mega, monkeypatch):
    """64 frames, batch 8 (4 launches of 16 frames in the production path): tapped run == untapped run bit for bit, and
    three rows are held to the oracle at every frame (module docstring, items 1-6).  "ring" has the 1.7B's matrix shapes
    with 2 + 2 layers; "ring-mega4" runs it on the TMA-ring generation of the persistent kernel (Q3_MEGA=4)."""
    if mega is not None:
        monkeypatch.setenv("Q3_MEGA", mega)
    B, F = 8, 64
    opts = api.SynthesisOptions(max_length=F)
    prompts = [W.synthetic_prompt(200 + i, spec) for i in range(B)]
    seeds = [4242 + i for i in range(B)]
    tts = gpu_tts(spec)
    if mega is not None:
        probe = api.Session(tts.model, B, opts, seeds)
        assert probe.decode_generation() == int(mega), "the requested generation was replaced by a fallback"
        probe.close()
    plain = tts.generate_codes(prompts, options=opts, seeds=seeds)
    tapped, taps = run_tapped(tts, prompts, seeds, opts, F)
    assert tapped == plain                       # one frame per launch == 16 frames per launch, bit-identical
    rep = check_follow(spec, tts, prompts, seeds, opts, F, rows=(0, 3, 7), tapped=tapped, taps=taps)
    print("follow report:", rep)
    assert rep["frames"] == 3 * F and rep["sample_exempt"] == 0
    assert rep["oracle_argmax_differs"] <= rep["codes"] // 20       # near-ties (each one verified above) stay a small minority


@pytest.mark.parametrize("spec,batch,mega", [(S.SPEC_1_7B, 8, None), (S.SPEC_1_7B, 1, None), (S.SPEC_0_6B, 8, None), (S.SPEC_1_7B, 8, "4"), (S.SPEC_1_7B, 8, "5"), (S.SPEC_0_6B, 8, "5")],
                         ids=["1.7b-b8", "1.7b-b1", "0.6b-b8", "1.7b-b8-mega4", "1.7b-b8-mega5", "0.6b-b8-mega5"])
def test_baseline_dimensions_follow_the_oracle(spec, batch, mega, monkeypatch):
    """BASELINE.json's model dimensions (1.7B: hidden 2048, 28 layers, 16/8 heads, inter 6144, small_to_mtp projection;
    0.6B: hidden 1024, inter 3072, no projection), batch 8 and batch 1, 3 frames: the same six checks, i.e. the K = 6144
    down-projection, the 16/8-head GQA mapping, 28-layer error growth and the register-resident vs streaming kernel
    variants at their real sizes are compared with the oracle, through the production loop."""
    if mega is not None:
        monkeypatch.setenv("Q3_MEGA", mega)
    F = 3
    opts = api.SynthesisOptions(max_length=F)
    prompts = [W.synthetic_prompt(i, spec) for i in range(batch)]
    seeds = [42 + i for i in range(batch)]
    tts = gpu_tts(spec)
    plain = tts.generate_codes(prompts, options=opts, seeds=seeds)
    tapped, taps = run_tapped(tts, prompts, seeds, opts, F)
    assert tapped == plain
    rows = (0, 5) if batch > 1 else (0,)
    rep = check_follow(spec, tts, prompts, seeds, opts, F, rows=rows, tapped=tapped, taps=taps)
    print("follow report:", rep)
    assert rep["frames"] == len(rows) * F and rep["sample_exempt"] == 0


@pytest.mark.parametrize("mega", ["2", "4", "5"])
def test_code_predictor_frame_repeats_bit_for_bit_at_1p7b(mega, monkeypatch):
    """Race detector at BASELINE dimensions: the code-predictor frame (405 dependent phases of the persistent kernel, one
    launch) run 150 times on identical inputs must give bit-identical logits every time, on both generations.  (This is the
    test that caught the early slot release of the TMA-ring kernel: mbarrier.arrive scheduled ahead of the MMAs that consume
    the slot's fragments, 73 of 300 repetitions differed at batch 8; tools/cp_repeat.py, tools/cp_bisect.py.)"""
    monkeypatch.setenv("Q3_MEGA", mega)
    spec, B = S.SPEC_1_7B, 8
    tts = gpu_tts(spec)
    sess = api.Session(tts.model, B, api.SynthesisOptions(max_length=64), list(range(B)), max_seq=128)
    assert sess.decode_generation() == int(mega)
    g = torch.Generator().manual_seed(5)
    hid = (torch.randn(B, spec.hidden, generator=g) * 0.7).to(torch.bfloat16)
    toks = [100 + 37 * i for i in range(B)]
    ref_codes, ref = sess.code_predictor_frame(hid, toks, want_logits=True)
    for it in range(150):
        codes, lg = sess.code_predictor_frame(hid, toks, want_logits=True)
        assert np.array_equal(lg, ref) and np.array_equal(codes, ref_codes), ("repetition differs", it)
    sess.close()


def test_eos_row_follows_the_oracle_to_its_last_frame():
    """A row that stops early (EOS forced by a dominant codec_head row, as in test_gpu_model's EOS test): the replayed
    sampler must draw EOS right after the row's last emitted frame, from the CUDA path's own logits."""
    spec = S.SPEC_TINY
    from conftest import talker_weights
    from oracle import model as OM
    w = dict(talker_weights(spec))
    tk = OM.Talker(spec, w, OM.BF16P)
    ids = W.synthetic_prompt(0, spec)
    emb = tk.custom_voice_embeds(ids, S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"])
    hidden, _ = tk.run_prefill_layers(emb, tk.new_kv_caches(64))
    head = w["talker.codec_head.weight"].clone().float()
    head[2150] = 8.0 * hidden[0, -1] / hidden[0, -1].norm()
    w["talker.codec_head.weight"] = head.to(torch.bfloat16)
    tts = api.Qwen3TTS.from_weights(spec, w)
    opts = api.SynthesisOptions(max_length=12)
    prompts, seeds = [ids, ids], [42, 43]
    tapped, taps = run_tapped(tts, prompts, seeds, opts, 12)
    assert [len(r) for r in tapped] == [2, 2]
    # follow with the modified weights
    models = (OM.Talker(spec, w, OM.BF16P), OM.CodePredictor(spec, w, OM.BF16P))
    models32 = (OM.Talker(spec, w, OM.F32P), OM.CodePredictor(spec, w, OM.F32P))
    rep = check_follow(spec, tts, prompts, seeds, opts, 12, rows=(0, 1), tapped=tapped, taps=taps, models=models, models32=models32)
    assert rep["sampled"] == 2 * 3 and rep["sample_exempt"] == 0      # first token, token after frame 0, EOS after frame 1


def test_long_context_voice_design_prefill_follows_the_oracle():
    """Long context at BASELINE dimensions (0.6B: 28 layers, 16 query / 8 kv heads): a VoiceDesign prompt whose 700 instruct
    ids sit in the prefill (talker.rs:585-627: context = 700 + 9 positions before the first frame), batch 2 with one short
    row, then 3 decode frames -- the prefill attention, the decode attention over ~710 cached positions of 8 kv heads (above the
    split-KV threshold of 512: the long row's attention runs cut over 4 CTAs, m2_attn_units, the short row's unsplit) and the
    RoPE table far from position 0 are held to the oracle by the same six checks."""
    spec = S.SPEC_0_6B
    F = 3
    opts = api.SynthesisOptions(max_length=F)
    prompts = [W.synthetic_prompt(i, spec) for i in range(2)]
    g = torch.Generator().manual_seed(99)
    instruct = [torch.randint(0, 150000, (700,), generator=g).tolist(), torch.randint(0, 150000, (5,), generator=g).tolist()]
    seeds = [7, 8]
    tts = gpu_tts(spec)
    tapped, taps = run_tapped(tts, prompts, seeds, opts, F, instruct)
    # one frame per launch (tapped) == 16 frames per launch (production) at a split-KV context as well
    pp = [tts.voice_design_prompt(t, ins, "english") for t, ins in zip(prompts, instruct)]
    sess = tts._new_session(prompts, pp, opts, seeds, max_seq=max(len(p[0]) for p in pp) + F + 40)
    codes, n = sess.generate(F)
    sess.close()
    assert [codes[b, : n[b]].tolist() for b in range(2)] == tapped
    rep = check_follow(spec, tts, prompts, seeds, opts, F, rows=(0, 1), tapped=tapped, taps=taps, instruct=instruct)
    print("follow report:", rep)
    assert rep["frames"] == 2 * F and rep["sample_exempt"] == 0


@pytest.mark.parametrize("spec", [S.SPEC_TINY_PROJ, S.SPEC_0_6B], ids=["tiny_proj", "0.6b"])
def test_voice_clone_prompts_follow_the_oracle(spec):
    """SURVEY 8(f) row 4, talker side: voice-clone prefill with a continuous speaker embedding (talker.rs:511-564) and the ICL
    block (sum_ref_codec_embeddings lib.rs:1239-1257 + build_icl_prompt talker.rs:646-705, run here as part of one causal
    prefill).  Batch 3: x-vector only; ICL whose text outlasts the reference codes (trailing = the text remainder); ICL whose
    reference codes outlast the text (text padded with tts_pad, no trailing rows).  Same six follow-mode checks."""
    F = 3
    opts = api.SynthesisOptions(max_length=F, repetition_penalty=1.5)
    g = torch.Generator().manual_seed(2024)
    H = spec.hidden
    hi = min(151643, spec.text_vocab - 300) if spec.text_vocab > 4096 else spec.text_vocab
    rnd_ids = lambda n: torch.randint(0, hi, (n,), generator=g).tolist()
    def ref_codes(t):
        c = torch.randint(0, 2048, (t, 16), generator=g).numpy().astype(np.uint32)
        return c
    spk = lambda: (torch.randn(H, generator=g) * 0.05)
    prompts = [rnd_ids(9), rnd_ids(14), rnd_ids(4)]
    clone = [api.VoiceClonePrompt(spk()),
             api.VoiceClonePrompt(spk(), ref_codes(6), rnd_ids(5)),          # n_text = 5 + 14 + 1 = 20 > n_codec = 7
             api.VoiceClonePrompt(spk(), ref_codes(12), rnd_ids(3))]         # n_text = 3 + 4 + 1 = 8 < n_codec = 13
    seeds = [5, 6, 7]
    tts = gpu_tts(spec)
    # the batch API refuses to mix ICL and x-vector rows (different repetition penalties); sessions themselves do not care
    tapped, taps = run_tapped(tts, prompts, seeds, opts, F, clone=clone)
    rep = check_follow(spec, tts, prompts, seeds, opts, F, rows=(0, 1, 2), tapped=tapped, taps=taps, clone=clone)
    print("follow report:", rep)
    assert rep["frames"] == 3 * F and rep["sample_exempt"] == 0


def test_split_kv_attention_is_row_independent_and_the_unsplit_form_is_untouched(monkeypatch):
    """Split-KV decode attention (m2_attn_units): rows whose context has reached the threshold are cut over 4 CTAs that
    exchange the softmax maximum / normaliser and their partial P*V sums.  (a) With the unit-mapped path active but no row
    long enough to be split, the codes are those of the unsplit kernel, bit for bit; (b) with splitting on, a batch that
    mixes long and short rows is deterministic and every row equals its batch-1 run (whether a row is split depends on its
    own context only).  The comparison with the oracle at a split context is
    test_long_context_voice_design_prefill_follows_the_oracle (710 positions >= the default threshold of 512)."""
    spec = S.SPEC_RING
    tts = gpu_tts(spec)
    F = 20
    g = torch.Generator().manual_seed(17)
    prompts = [W.synthetic_prompt(i, spec) for i in range(4)]
    instr = [torch.randint(0, 1500, (n,), generator=g).tolist() for n in (640, 9, 580, 30)]
    pp = [tts.voice_design_prompt(t, ins, "english") for t, ins in zip(prompts, instr)]
    opts = api.SynthesisOptions(max_length=F, eos_token_id=None)

    def run(rows):
        sess = api.Session(tts.model, len(rows), opts, [42 + r for r in rows], max_seq=640 + F + 64)
        sess.prefill_ids([pp[r][0] for r in rows], [pp[r][1] for r in rows])
        sess.set_trailing_ids([list(prompts[r][1:]) for r in rows])
        codes, n = sess.generate(F)
        sess.close()
        return codes

    monkeypatch.setenv("Q3_SPLIT_KV", "0")
    unsplit = run([0, 1, 2, 3])
    monkeypatch.setenv("Q3_SPLIT_KV", "100000")          # unit-mapped path, nothing long enough to split
    assert np.array_equal(run([0, 1, 2, 3]), unsplit)
    monkeypatch.setenv("Q3_SPLIT_KV", "512")
    split = run([0, 1, 2, 3])
    assert np.array_equal(run([0, 1, 2, 3]), split)       # deterministic
    assert np.array_equal(split[1], unsplit[1]) and np.array_equal(split[3], unsplit[3])     # short rows: untouched
    for r in range(4):
        assert np.array_equal(run([r])[0], split[r]), ("row differs from its batch-1 run", r)
    # the threshold is crossed inside the run: 505 + 9 prefill positions, 20 frames
    instr2 = torch.randint(0, 1500, (495,), generator=g).tolist()
    pp[0] = tts.voice_design_prompt(prompts[0], instr2, "english")
    a, b2 = run([0, 1]), run([0])
    assert np.array_equal(a[0], b2[0])
