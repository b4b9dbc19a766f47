/*
 * q3tts.h -- C ABI of libq3tts_b200.so, the B200 (sm_100a) implementation of the Qwen3-TTS
 * autoregressive decode hot path of TrevorS/qwen3-tts-rs.
 *
 * This header is the drop-in boundary.  Every entry point names the reference
 * interface it replaces as `ref: file:line` (paths relative to the reference repository).  The
 * reference-side binding a maintainer would add (Rust `extern "C"` block + safe wrappers behind
 * src/generation and src/models) is in INTEGRATION.md.
 *
 * Conventions
 *   - every call returns a q3_status; 0 = ok.  q3_last_error() returns a thread-local message.
 *     No C++ exception or abort crosses the boundary (ref: anyhow::Result at the API,
 *     candle_core::Result inside ops; bail! on overflow/shape errors).
 *   - all handles are opaque; the caller owns every output buffer it passes in.
 *   - a model is immutable after q3_model_finalize and may be shared by sessions/threads
 *     (ref: `&self` methods, src/lib.rs:530-541); a session owns all mutable state (KV caches,
 *     penalty masks, RNG states, offsets) and one CUDA stream, and is not thread-safe
 *     (ref: StreamingSession fields, src/lib.rs:1484-1508).
 *   - "host" pointers are ordinary (ideally pinned) host memory; "dev" pointers are CUDA device
 *     memory on the model's device.  bf16 values are passed as uint16_t bit patterns.
 *   - there is NO CPU fallback: every compute entry point fails with Q3_ERR_CUDA when no
 *     sm_100 device is available.
 */
#ifndef Q3TTS_H
#define Q3TTS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Q3_ABI_VERSION 1
#define Q3_CODEC_VOCAB 3072         /* ref: src/models/talker.rs:54  CODEC_VOCAB_SIZE */
#define Q3_CODEC_EOS 2150           /* ref: src/lib.rs:1466          CODEC_EOS_TOKEN_ID */
#define Q3_SAMPLES_PER_FRAME 1920   /* ref: src/lib.rs:1469          SAMPLES_PER_FRAME */
#define Q3_CODES_PER_FRAME 16       /* ref: src/lib.rs:119-121       [semantic, a0..a14] */

typedef enum q3_status {
  Q3_OK = 0,
  Q3_ERR_INVALID = 1,        /* bad argument / shape (ref: bail! on non-contiguous or wrong dims) */
  Q3_ERR_CUDA = 2,           /* CUDA runtime error or no usable device */
  Q3_ERR_KV_OVERFLOW = 3,    /* ref: src/models/kv_cache.rs:293-300 "KV cache overflow" */
  Q3_ERR_MISSING_WEIGHT = 4, /* ref: src/models/codec/decoder_12hz.rs:176-181 "Missing weight" */
  Q3_ERR_STATE = 5,          /* call made in the wrong session/model state */
  Q3_ERR_UNSUPPORTED = 6
} q3_status;

typedef enum q3_dtype { Q3_BF16 = 0, Q3_F32 = 1 } q3_dtype;

typedef struct q3_model q3_model;
typedef struct q3_session q3_session;

/* Dimension table.  ref: TalkerConfig (src/models/talker.rs:208-274), CodePredictorConfig
 * (src/models/code_predictor.rs:48-113), Decoder12HzConfig (src/models/codec/decoder_12hz.rs:47-67).
 * Parsing config.json stays with the caller. */
typedef struct q3_model_desc {
  int32_t hidden, inter, layers, heads, kv_heads, head_dim;
  int32_t codec_vocab, text_vocab, text_embed_dim;
  float rope_theta, rms_eps;
  int32_t cp_hidden, cp_inter, cp_layers, cp_heads, cp_kv_heads, cp_vocab, groups;
  int32_t cp_rope_positions, cp_max_seq;
  /* vocoder */
  int32_t v_codebook_dim, v_vq_dim, v_latent_dim, v_hidden, v_layers, v_heads, v_head_dim, v_inter;
  int32_t v_quantizers, v_codebook_size, v_decoder_dim;
  int32_t v_n_upsampling, v_upsampling[4];
  int32_t v_n_rates, v_rates[8];
  float v_rms_eps, v_rope_theta;
  int32_t device;            /* CUDA device ordinal */
} q3_model_desc;

/* ref: GenerationConfig (src/generation/sampling.rs:100-115) + SynthesisOptions.chunk_frames
 * (src/lib.rs:1786-1805).  eos_token_id < 0 means None. */
typedef struct q3_gen_config {
  int32_t max_new_tokens;
  double temperature;
  int32_t top_k;
  double top_p;
  double repetition_penalty;
  int32_t eos_token_id;
  int32_t min_new_tokens;
  int32_t chunk_frames;
} q3_gen_config;

const char* q3_last_error(void);
int q3_abi_version(void);
/* number of CUDA kernels this library has launched in the calling process (monotone counter) */
uint64_t q3_kernel_launch_count(void);

/* ---- model ------------------------------------------------------------------------------
 * ref: Qwen3TTS::from_weights (src/lib.rs:267-274) with the HuggingFace tensor names the
 * reference loads (docs/QWEN3_TTS_ARCHITECTURE.md:431-459; decoder_12hz.rs:191-381). */
q3_status q3_model_create(const q3_model_desc* desc, q3_model** out);
/* The library copies (and re-packs) the tensor; the caller's buffer may be freed afterwards.
 * Talker/code-predictor tensors are stored bf16 (an F32 source is rounded), vocoder tensors F32. */
q3_status q3_model_set_tensor(q3_model* m, const char* hf_name, const void* data, q3_dtype dtype,
                              const int64_t* shape, int32_t ndim, int32_t on_device);
q3_status q3_model_finalize(q3_model* m);
void q3_model_destroy(q3_model* m);

/* ---- session ----------------------------------------------------------------------------
 * ref: TalkerModel::new_kv_caches (src/models/talker.rs:891-913) + SamplingContext::new
 * (src/generation/sampling.rs:32-51) + mask setup (src/lib.rs:543-571).  One session holds
 * `batch` independent utterances (row i == an independent batch-1 reference run with seed i). */
q3_status q3_session_create(const q3_model* m, int32_t batch, int32_t max_seq, const q3_gen_config* cfg,
                            const uint64_t* seeds /*[batch]*/, q3_session** out);
q3_status q3_session_reset(q3_session* s, const uint64_t* seeds /*[batch]*/);
void q3_session_destroy(q3_session* s);
/* CUDA stream (cudaStream_t) the session launches on, for callers that time with events. */
void* q3_session_stream(q3_session* s);
q3_status q3_session_synchronize(q3_session* s);

/* Prefill from ready-made input embeddings.  embeds: bf16 [batch][l_max][hidden] (host), lens[batch].
 * ref: TalkerModel::run_prefill_layers (src/models/talker.rs:823-841). */
q3_status q3_prefill_embeds(q3_session* s, const uint16_t* embeds, const int32_t* lens, int32_t l_max);
/* Prefill with on-device prompt assembly: position p of row b is
 *   text_proj(text_embedding[text_ids[b][p]])  (if text_ids >= 0)  (+)  codec_embedding[codec_ids[b][p]]  (if >= 0)
 * which covers prefill_custom_voice / prefill_voice_design (src/models/talker.rs:451-491, 585-627). */
q3_status q3_prefill_ids(q3_session* s, const int32_t* text_ids, const int32_t* codec_ids,
                         const int32_t* lens, int32_t l_max);
/* ECAPA-TDNN speaker encoder on a mel spectrogram (voice-clone front end).
 * ref: SpeakerEncoder::forward (src/models/speaker.rs:448-476): initial TDNN, three squeeze-excitation Res2Net blocks, multi-layer
 * feature aggregation, attentive statistics pooling, FC; raw (un-normalised) embedding.  Weights `speaker_encoder.*` (F32) through
 * q3_model_set_tensor; dilations are the reference's config defaults (src/models/config.rs:144-146).  The mel front end
 * (MelSpectrogram::compute_for_speaker_encoder, src/audio/mel.rs) stays on the host side of the boundary.
 * mel: f32 [batch][mel_dim][t] (host); embed_out: f32 [batch][q3_speaker_embed_dim(m)] (host). */
int32_t q3_speaker_embed_dim(const q3_model* m);
q3_status q3_speaker_encode(const q3_model* m, const float* mel, int32_t batch, int32_t t, float* embed_out);
/* Voice-clone prompt, x-vector only or with an in-context (ICL) reference (talker side: the speaker and
 * speech encoders that produce `speaker_embeds` and `ref_codes` are out of scope).  As q3_prefill_ids, with two more kinds of
 * codec part per position:
 *   codec_ids[b][p] == Q3_POS_SPEAKER (-2)      : speaker_embeds[b] (bf16 [batch][hidden]), the continuous speaker embedding of
 *                                                 prefill_voice_clone (src/models/talker.rs:511-564, position 7)
 *   codec_ids[b][p] == Q3_POS_REF_FRAME(t)      : the 16-way embedding sum of reference frame t of row b,
 *                                                 codec_embedding[c0] + code_predictor.codec_embedding[g-1][c_g], g = 1..15 in
 *                                                 order (sum_ref_codec_embeddings, src/lib.rs:1239-1257)
 * ref_codes: u32 [batch][t_ref_max][16], t_ref[batch] frames valid (NULL / 0 for x-vector only).  The reference runs the ICL
 * block as a second causal chunk behind the 9-position prefill (src/lib.rs:953-987); one causal prefill over the
 * concatenation fills the same KV cache and ends in the same last hidden state and logits, so the host mirror passes
 * [prefill positions ++ build_icl_prompt positions (streaming overlay, src/models/talker.rs:684-704)] in one call. */
#define Q3_POS_SPEAKER (-2)
#define Q3_POS_REF_FRAME(t) (-16 - (t))
q3_status q3_prefill_voice_clone(q3_session* s, const int32_t* text_ids, const int32_t* codec_ids, const int32_t* lens,
                                 int32_t l_max, const uint16_t* speaker_embeds, const uint32_t* ref_codes,
                                 const int32_t* t_ref, int32_t t_ref_max);
/* ref: Qwen3TTS::build_trailing_text (src/lib.rs:508-519).  trailing: bf16 [batch][lt_max][hidden]. */
q3_status q3_set_trailing_text(q3_session* s, const uint16_t* trailing, const int32_t* lt, int32_t lt_max,
                               const uint16_t* tts_pad /*[hidden]*/);
/* Same from token ids: rows are text_proj(ids[b][0..n-1]) ++ text_proj(tts_eos); pad = text_proj(tts_pad).
 * n[b] == -1: no trailing rows at all -- every frame adds tts_pad (an ICL prompt that consumed the whole text,
 * src/models/talker.rs:691-703). */
q3_status q3_set_trailing_ids(q3_session* s, const int32_t* ids, const int32_t* n, int32_t n_max,
                              int32_t tts_eos_id, int32_t tts_pad_id);

/* The production path.  ref: Qwen3TTS::generate_codes (src/lib.rs:530-656).  Runs up to max_frames
 * frames for every row with no per-frame host sync; codes: u32 [batch][max_frames][16] (host),
 * n_frames[batch].  Rows stop at their own EOS; the EOS token's frame is not emitted. */
q3_status q3_generate(q3_session* s, int32_t max_frames, uint32_t* codes, int32_t* n_frames);
/* Same loop, results left on the device (inputs/outputs resident in HBM): enqueue only. */
q3_status q3_generate_async(q3_session* s, int32_t max_frames);
q3_status q3_get_codes(q3_session* s, int32_t max_frames, uint32_t* codes, int32_t* n_frames);

/* ref: StreamingSession::next_chunk (src/lib.rs:1650-1759).  Generates up to chunk_frames frames
 * per row and vocodes each row's chunk independently (no vocoder state crosses chunks).
 * codes: u32 [batch][chunk_frames][16]; pcm: f32 [batch][chunk_frames*1920]; n_frames[batch];
 * *done != 0 when every row has finished and nothing is buffered. */
q3_status q3_stream_next(q3_session* s, uint32_t* codes, float* pcm, int32_t* n_frames, int32_t* done);
/* Opt-in extension (no reference counterpart -- the reference decodes every chunk without left
 * context, src/lib.rs:1755-1758, so streamed PCM differs from non-streamed PCM at chunk starts).
 *   left_context_frames  = 0 (default): the reference's stateless chunks.
 *   left_context_frames  < 0: STATEFUL streaming.  The session carries the vocoder's cross-chunk state (keys / values of the
 *     pre-transformer for every frame so far, and the last 10 frames of the conv stack's input, whose look-back is 9.4
 *     frames), so every q3_stream_next costs O(chunk + 10 frames) of vocoder work however long the utterance is, and the
 *     streamed PCM is identical to q3_vocode_session's (every vocoder op is causal).  Any chunk size >= 1 frame.
 *   left_context_frames  = c > 0: re-decode form -- the previous min(c, frames so far) frames are decoded again in front of
 *     the chunk and their samples dropped; approximate for finite c (the pre-transformer sees only c frames of history). */
q3_status q3_session_set_stream_context(q3_session* s, int32_t left_context_frames);
/* Extension (no reference counterpart; StreamingSession::next_chunk always emits chunk_frames, src/lib.rs:1650-1759): the FIRST
 * q3_stream_next of the session generates only min(first_chunk_frames, chunk_frames) frames, later ones chunk_frames -- with the
 * stateful form above the waveform does not depend on where the chunks are cut, so a 2-frame first chunk gives the time to first
 * audio of 2-frame chunks at the throughput of 10-frame chunks.  0 (default) = every chunk has chunk_frames. */
q3_status q3_session_set_first_chunk(q3_session* s, int32_t first_chunk_frames);

/* ref: Decoder12Hz::decode (src/models/codec/decoder_12hz.rs:411-505).  codes: i64 [B][16][T]
 * (host, the codes_to_tensor layout of src/lib.rs:1417-1431); pcm: f32 [B][T*1920] (host). */
q3_status q3_vocoder_decode(const q3_model* m, const int64_t* codes, int32_t batch, int32_t t, float* pcm);
/* Vocode the frames a session generated (device-resident codes, per-row lengths), pcm to host:
 * pcm f32 [batch][max_frames*1920]; rows are zero-filled past their own length.  Pass pcm == NULL
 * to leave the result on the device (bench `value` leg). */
q3_status q3_vocode_session(q3_session* s, int32_t max_frames, float* pcm);

/* ---- fine-grained entry points for per-op parity tests -------------------------------------- */
/* ref: TalkerModel::generate_step_with_embed (src/models/talker.rs:716-736).  step_input: bf16
 * [batch][hidden] (host).  Appends to the session KV cache at each row's offset.
 * hidden_out: bf16 [batch][hidden] (post-norm), logits_out: f32 [batch][codec_vocab]. */
q3_status q3_talker_step(q3_session* s, const uint16_t* step_input, uint16_t* hidden_out, float* logits_out);
/* ref: CodePredictor::generate_acoustic_codes (src/models/code_predictor.rs:320-416).
 * last_hidden: bf16 [batch][hidden]; sem_tokens[batch]; codes_out: u32 [batch][15];
 * logits_out (optional): f32 [batch][15][cp_vocab]. */
q3_status q3_code_predictor_frame(q3_session* s, const uint16_t* last_hidden, const uint32_t* sem_tokens,
                                  uint32_t* codes_out, float* logits_out);
/* ref: apply_generation_penalties_gpu + generation::sample + update_penalty_mask
 * (src/lib.rs:1271-1322, src/generation/sampling.rs:140-319, src/lib.rs:662-673).
 * logits: f32 [batch][vocab] (host); rng_states[batch] in/out (PCG state, sampling.rs:84-94);
 * seen_mask: u8 [batch][vocab] in/out; token_count: tokens sampled so far for these rows. */
q3_status q3_sample(const q3_model* m, const float* logits, int32_t batch, int32_t vocab, const q3_gen_config* cfg,
                    uint64_t* rng_states, uint8_t* seen_mask, int32_t token_count, uint32_t* tokens_out);
/* ref: FusedRmsNorm::forward_residual + kernels/fused_residual_rmsnorm.cu (src/models/fused_ops.rs:49-96):
 * sum = x + r (stored rounded), normed = rms_norm(sum) * w.  All pointers are DEVICE pointers of
 * `dtype`; out_normed/out_sum may be the two halves of one [2*rows, cols] buffer as the reference's
 * CustomOp2 returns it.  stream: cudaStream_t or NULL. */
q3_status q3_fused_residual_rmsnorm(const void* x, const void* r, const void* w, void* out_normed, void* out_sum,
                                    int32_t rows, int32_t cols, float eps, q3_dtype dtype, void* stream);
/* Host-buffer convenience wrapper of the same op (copies in/out), for tests without a CUDA allocator. */
q3_status q3_fused_residual_rmsnorm_host(const void* x, const void* r, const void* w, void* out_normed, void* out_sum,
                                         int32_t rows, int32_t cols, float eps, q3_dtype dtype, int32_t device);

/* Per-stage device timing of the last q3_generate / q3_vocode_session call, in milliseconds
 * (ref: SynthesisTiming, src/lib.rs:136-147). */
typedef struct q3_timing { float prefill_ms, generation_ms, decode_ms; int32_t generation_frames; } q3_timing;
q3_status q3_session_timing(q3_session* s, q3_timing* out);

#ifdef __cplusplus
}
#endif
#endif /* Q3TTS_H */
