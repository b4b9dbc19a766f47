"""Counterpart of the reference's `generate_audio` binary (src/bin/generate_audio.rs) for the decode path: load a
checkpoint directory, generate codes with a seed, decode to audio, and leave the same file set in --output-dir
(codes_seed{S}_frames{N}.bin, audio_..wav, audio_..bin, metadata_..json); --compare reads another build's dumps from
--reference-dir.  Tokenisation is outside the hot path, so the text is given as token ids (--input-ids "1,2,3", or
--synthetic-prompt I for utterance I of the synthetic prompt set).

  # write the synthetic 1.7B CustomVoice checkpoint where a reference build can load it too (CPU only)
  python tools/generate_audio.py --export-synthetic 1.7b --model-dir /tmp/synth-1.7b
  # generate on the GPU from it and compare with dumps made elsewhere
  python tools/generate_audio.py --model-dir /tmp/synth-1.7b --synthetic-prompt 0 --seed 42 --frames 64 \
      --output-dir out --compare --reference-dir ref_dumps
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from qwen3_tts_rs_b200 import formats as F, spec as S, weights as W  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--input-ids", default=None, help="comma-separated text token ids")
    ap.add_argument("--synthetic-prompt", type=int, default=None)
    ap.add_argument("-s", "--seed", type=int, default=42)
    ap.add_argument("-f", "--frames", type=int, default=2048)
    ap.add_argument("-d", "--duration", type=float, default=None)
    ap.add_argument("--temperature", type=float, default=0.7)          # the binary's default, not SynthesisOptions'
    ap.add_argument("--top-k", type=int, default=50)
    ap.add_argument("--top-p", type=float, default=0.9)
    ap.add_argument("--repetition-penalty", type=float, default=1.05)
    ap.add_argument("-m", "--model-dir", default="test_data/model")
    ap.add_argument("-o", "--output-dir", default="test_data/rust_audio")
    ap.add_argument("--output", default=None)
    ap.add_argument("-c", "--compare", action="store_true")
    ap.add_argument("--reference-dir", default="test_data/reference_audio")
    ap.add_argument("--speaker", default="ryan")
    ap.add_argument("--language", default="english")
    ap.add_argument("--instruct-ids", default=None, help="comma-separated token ids of the voice description (VoiceDesign)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--export-synthetic", default=None, choices=sorted(S.SPECS),
                    help="write the synthetic checkpoint of this dimension table into --model-dir and exit")
    ap.add_argument("--model-type", default="custom_voice", choices=F.MODEL_TYPES)
    a = ap.parse_args(argv)

    if a.export_synthetic:
        spec = S.SPECS[a.export_synthetic]
        F.export_checkpoint(a.model_dir, spec, W.make_talker_weights(spec), W.make_vocoder_weights(spec.vocoder), a.model_type)
        print(f"wrote {a.model_dir}: config.json, model.safetensors, speech_tokenizer/model.safetensors ({spec.name})")
        return 0

    from qwen3_tts_rs_b200 import api
    tts = api.Qwen3TTS.from_pretrained(a.model_dir, a.device)
    if a.input_ids is not None:
        ids = [int(t) for t in a.input_ids.split(",") if t.strip()]
    else:
        ids = W.synthetic_prompt(a.synthetic_prompt or 0, tts.spec)
    mf = F.max_frames_from_args(a.frames, a.duration)
    opts = api.SynthesisOptions(max_length=mf, temperature=a.temperature, top_k=a.top_k, top_p=a.top_p,
                                repetition_penalty=a.repetition_penalty, seed=a.seed)
    print(f"model: {tts.model_type or 'variant unknown'} {tts.spec.name}; {len(ids)} text tokens; seed {a.seed}; max {mf} frames")
    if a.instruct_ids is not None:
        inst = [int(t) for t in a.instruct_ids.split(",") if t.strip()]
        prompts = [tts.voice_design_prompt(ids, inst, a.language)]
    else:
        prompts = [tts.custom_voice_prompt(ids, a.speaker, a.language)]
    # one session: prefill -> generate_codes -> vocoder, the body of synthesize_with_timing (lib.rs:425-501)
    sess = tts._new_session([ids], prompts, opts, [a.seed], max_seq=max(mf + 256, len(prompts[0][0]) + mf))
    try:
        codes_arr, n = sess.generate(mf)
        pcm = sess.vocode(mf)
        timing = sess.timing()
        codes = codes_arr[0, : n[0]].tolist()
        samples = pcm[0, : n[0] * api.SAMPLES_PER_FRAME].copy()
    finally:
        sess.close()
    print(f"frames: {len(codes)}; audio samples: {samples.size} ({samples.size / 24000.0:.3f}s at 24kHz); "
          f"prefill {timing.prefill_ms:.1f} ms, generation {timing.generation_ms:.1f} ms, decode {timing.decode_ms:.1f} ms")
    paths = F.write_generation_outputs(a.output_dir, a.seed, codes, samples, "", ids, a.temperature, a.top_k, a.top_p, a.output)
    for k, p in paths.items():
        print(f"saved {k}: {p}")
    if a.compare:
        rep = F.compare_with_reference(a.reference_dir, a.seed, len(codes), codes, samples)
        print("=== comparing with reference dumps ===")
        if not rep.codes_found:
            print("Codes: reference not found")
        elif rep.codes_match:
            print(f"Codes: MATCH (all {rep.n_ref_codes} values identical)")
        else:
            print(f"Codes: MISMATCH  reference {rep.n_ref_codes} values, ours {rep.n_our_codes}, {rep.n_code_diffs} differences; first: {rep.first_code_diffs}")
        if rep.audio_found:
            print(f"Audio: {rep.n_audio_compared} samples compared, max diff {rep.max_diff:.6g}, mean diff {rep.mean_diff:.6g}, rmse {rep.rmse:.6g}")
        else:
            print("Audio: reference not found")
    return 0


if __name__ == "__main__":
    sys.exit(main())
