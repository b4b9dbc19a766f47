"""CPU tests: the C-ABI library loads and exports every symbol include/q3tts.h declares (no compute calls),
host-side prompt assembly mirrors the oracle, and the product package never imports the oracle."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from qwen3_tts_rs_b200 import api, lib as L, spec as S, weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "q3tts.h")).read()
    declared = set(re.findall(r"\b(q3_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"q3_status"}
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.q3_abi_version() == 1
    assert lib.q3_kernel_launch_count() >= 0


def test_struct_layouts_match_the_header():
    """ctypes mirrors of q3_model_desc / q3_gen_config must match the C layout (checked by compiling a probe)."""
    src = '#include <stdio.h>\n#include "q3tts.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(q3_model_desc), sizeof(q3_gen_config), sizeof(q3_timing));return 0;}'
    exe = os.path.join("/tmp", "q3_layout_probe")
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src, text=True, check=True)
    a, b, c = map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    assert (a, b, c) == (C.sizeof(L.ModelDesc), C.sizeof(L.GenConfig), C.sizeof(L.Timing))


def test_no_gpu_means_loud_failure(lib):
    """There is no CPU fallback: without a usable sm_100 device model creation fails with Q3_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(L.Q3Error) as e:
        api.Model(S.SPEC_TINY)
    assert e.value.status == "Q3_ERR_CUDA"


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "qwen3_tts_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
    r = subprocess.run([sys.executable, "-c",
                        "import sys; import qwen3_tts_rs_b200.api, qwen3_tts_rs_b200.lib; "
                        "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"],
                       cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_prompt_id_lists_mirror_the_oracle_embeddings():
    """custom_voice_prompt / voice_design_prompt produce, position by position, the (text id, codec id) pairs whose
    embeddings the oracle's prefill builders add up."""
    import torch
    from oracle import model as OM
    from conftest import talker_weights
    spec = S.SPEC_TINY
    tk = OM.Talker(spec, talker_weights(spec), OM.BF16P)
    tts = api.Qwen3TTS.__new__(api.Qwen3TTS)
    tts.spec = spec
    ids = W.synthetic_prompt(4, spec)

    def embed(text, codec):
        rows = []
        for t, c in zip(text, codec):
            e = None
            if t >= 0:
                e = tk.projected_text([t])[0, 0]
            if c >= 0:
                ce = tk.codec_embedding[c]
                e = ce if e is None else OM.BF16P.r(e + ce)
            rows.append(e)
        return torch.stack(rows)[None]
    text, codec = tts.custom_voice_prompt(ids, "ryan", "english")
    assert len(text) == len(codec) == 10
    assert torch.equal(embed(text, codec), tk.custom_voice_embeds(ids, S.SPEAKER_IDS["ryan"], S.LANGUAGE_IDS["english"]))
    ins = list(range(50, 74))
    text, codec = tts.voice_design_prompt(ids, ins, "german")
    assert len(text) == 24 + 9
    assert torch.equal(embed(text, codec), tk.voice_design_embeds(ids, ins, S.LANGUAGE_IDS["german"]))
    text, codec = tts.custom_voice_prompt([], "ryan", "english")
    assert len(text) == 9                      # no text token -> no 10th position (talker.rs:484-488)


def test_voice_clone_prompt_lists_mirror_the_oracle_embeddings():
    """voice_clone_prompt (x-vector only, ICL with a text remainder, ICL with padded text): the (text id, codec part) pairs,
    with Q3_POS_SPEAKER and Q3_POS_REF_FRAME(t) parts resolved as the CUDA kernel resolves them, add up to the oracle's
    prefill_voice_clone ++ build_icl_prompt embeddings bit for bit, and the trailing ids to its trailing text."""
    import numpy as np
    import torch
    from oracle import model as OM
    from conftest import talker_weights
    spec = S.SPEC_TINY_PROJ
    w = talker_weights(spec)
    tk, cp = OM.Talker(spec, w, OM.BF16P), OM.CodePredictor(spec, w, OM.BF16P)
    tts = api.Qwen3TTS.__new__(api.Qwen3TTS)
    tts.spec = spec
    g = torch.Generator().manual_seed(3)
    spk = torch.randn(spec.hidden, generator=g) * 0.05
    r = OM.BF16P.r

    def embed(text, codec, prompt):
        rows = []
        for t, c in zip(text, codec):
            ce = None
            if c >= 0:
                ce = tk.codec_embedding[c]
            elif c == api.POS_SPEAKER:
                ce = r(spk.float())
            elif c <= -16:
                fr = [int(x) for x in prompt.ref_codes[-16 - c]]
                ce = tk.codec_embedding[fr[0]]
                for gi in range(1, 16):
                    ce = r(ce + cp.codec_embeddings[gi - 1][fr[gi]])
            e = tk.projected_text([t])[0, 0] if t >= 0 else None
            rows.append(ce if e is None else (e if ce is None else r(e + ce)))
        return torch.stack(rows)[None]

    ids = W.synthetic_prompt(4, spec)
    lang = S.LANGUAGE_IDS["english"]
    for ref_t, ref_text in ((None, None), (5, [11, 12, 13]), (40, [11, 12])):
        rc = None if ref_t is None else torch.randint(0, 2048, (ref_t, 16), generator=g).numpy().astype(np.uint32)
        prompt = api.VoiceClonePrompt(spk, rc, ref_text)
        text, codec, trailing = tts.voice_clone_prompt(ids, prompt, "english")
        emb, tr = OM.voice_clone_prompt(tk, cp, ids, spk, lang, rc, ref_text)
        assert len(text) == len(codec) == emb.shape[1]
        assert torch.equal(embed(text, codec, prompt), emb)
        if trailing is None:                       # no rows: every frame adds tts_pad, as the reference's [tts_pad] trailing does
            assert torch.equal(tr, tk.tts_pad_embed())
        else:
            assert torch.equal(torch.cat([tk.projected_text(trailing), tk.tts_eos_embed()], 1), tr)
    assert len(tts.voice_clone_prompt(ids, api.VoiceClonePrompt(spk), "english")[0]) == 10
    assert len(tts.voice_clone_prompt(ids, api.VoiceClonePrompt(spk, np.zeros((5, 16), np.uint32), [1]), "english")[0]) == 9 + 6


def test_synthesis_options_defaults_and_gen_config():
    o = api.SynthesisOptions()                 # lib.rs:1822-1836
    assert (o.max_length, o.temperature, o.top_k, o.top_p, o.repetition_penalty, o.eos_token_id, o.chunk_frames,
            o.min_new_tokens, o.seed) == (2048, 0.9, 50, 0.9, 1.05, 2150, 10, 2, None)
    g = o.to_gen_config()
    assert g.max_new_tokens == 2048 and g.eos_token_id == 2150 and abs(g.temperature - 0.9) < 1e-12
    assert api.SynthesisOptions(eos_token_id=None).to_gen_config().eos_token_id == -1
    assert api.CODEC_EOS_TOKEN_ID == 2150 and api.SAMPLES_PER_FRAME == 1920


def test_roofline_byte_counts_match_the_survey():
    """weight bytes per step."""
    assert S.talker_weight_bytes(S.SPEC_1_7B) == 2_831_403_008
    assert S.cp_weight_bytes_per_frame(S.SPEC_1_7B) == 2_359_672_320 + 62_945_280 + 62_914_560
    assert S.talker_weight_bytes(S.SPEC_0_6B) == 887_226_368
    assert S.cp_weight_bytes_per_frame(S.SPEC_0_6B) == 2_359_672_320 + 62_914_560
    assert S.kv_bytes_per_position(S.SPEC_1_7B) == 114_688


def test_weights_are_order_independent_and_named_like_the_checkpoint():
    a = W.make_tensor("talker.model.norm.weight", (16,), "norm")
    b = W.make_tensor("talker.model.norm.weight", (16,), "norm")
    assert (a == b).all()
    names = {n for n, _, _ in W.talker_tensor_specs(S.SPEC_1_7B)}
    assert "talker.code_predictor.small_to_mtp_projection.bias" in names
    assert "talker.code_predictor.small_to_mtp_projection.weight" not in {n for n, _, _ in W.talker_tensor_specs(S.SPEC_0_6B)}
    assert sum(int(np.prod(s)) for n, s, _ in W.vocoder_tensor_specs(S.VocoderSpec()) if "cluster_usage" not in n) > 100e6


def test_rust_ffi_declares_every_symbol_of_the_header():
    """rust/src/b200/ffi.rs is shipped as source (no Rust toolchain here): at least keep it in step with the header --
    every q3_* function and both config structs' fields, in order."""
    hdr = open(os.path.join(ROOT, "include", "q3tts.h")).read()
    rs = open(os.path.join(ROOT, "rust", "src", "b200", "ffi.rs")).read()
    declared = set(re.findall(r"\b(q3_[a-z0-9_]+)\s*\(", hdr)) - {"q3_status"}
    in_rust = set(re.findall(r"pub fn (q3_[a-z0-9_]+)\s*\(", rs))
    assert declared == in_rust, declared ^ in_rust
    for struct in ("q3_model_desc", "q3_gen_config", "q3_timing"):
        body = re.search(r"typedef struct " + struct + r" \{(.*?)\} " + struct, hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        c_fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                c_fields += [re.sub(r"\[.*", "", f.strip()) for f in decl.split(None, 1)[1].split(",")]
        rbody = re.search(r"pub struct " + struct + r" \{(.*?)\n\}", rs, re.S).group(1)
        r_fields = re.findall(r"pub ([a-z0-9_]+):", rbody)
        assert c_fields == r_fields, (struct, c_fields, r_fields)
