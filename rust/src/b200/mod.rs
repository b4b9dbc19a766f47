//! B200 backend for the decode hot path: safe wrappers over libq3tts_b200.so that slot in behind the crate's existing
//! API (`Qwen3TTS::synthesize_with_voice`, `synthesize_voice_design`, `generate_codes`, `decode_codes`,
//! `synthesize_streaming`).  Enabled by the `b200` cargo feature; nothing here uses candle for compute -- candle tensors
//! are only the source of the weight bytes.
//!
//! Call order per request (identical to include/q3tts.hpp and qwen3_tts_rs_b200/api.py, which are compiled / tested):
//!   B200Session::new -> prefill_ids -> set_trailing_ids -> generate -> vocode        (non-streaming)
//!   B200Session::new -> prefill_ids -> set_trailing_ids -> stream_next ...           (streaming)
pub mod ffi;

use std::collections::HashMap;
use std::ffi::CString;
use std::ptr;

use anyhow::{anyhow, bail, Result};
use candle_core::{DType, Tensor};

use crate::audio::AudioBuffer;
use crate::models::talker::{Language, Speaker};
use crate::{FrameCodes, SynthesisOptions, SynthesisTiming};

pub const SAMPLES_PER_FRAME: usize = 1920; // src/lib.rs:1469

// special token ids (src/models/talker.rs:31-54)
const IM_START: i32 = 151644;
const ASSISTANT: i32 = 77091;
const NEWLINE: i32 = 198;
const TTS_PAD: i32 = 151671;
const TTS_BOS: i32 = 151672;
const TTS_EOS: i32 = 151673;
const CODEC_PAD: i32 = 2148;
const CODEC_BOS: i32 = 2149;
const CODEC_THINK: i32 = 2154;
const CODEC_THINK_BOS: i32 = 2156;
const CODEC_THINK_EOS: i32 = 2157;

/// Dimension table for the published checkpoints (`large` = the 1.7B talker, hidden 2048).
pub fn default_desc(large: bool, device: i32) -> ffi::q3_model_desc {
    ffi::q3_model_desc {
        hidden: if large { 2048 } else { 1024 },
        inter: if large { 6144 } else { 3072 },
        layers: 28,
        heads: 16,
        kv_heads: 8,
        head_dim: 128,
        codec_vocab: 3072,
        text_vocab: 151936,
        text_embed_dim: 2048,
        rope_theta: 1_000_000.0,
        rms_eps: 1e-6,
        cp_hidden: 1024,
        cp_inter: 3072,
        cp_layers: 5,
        cp_heads: 16,
        cp_kv_heads: 8,
        cp_vocab: 2048,
        groups: 16,
        cp_rope_positions: 1024,
        cp_max_seq: 17,
        v_codebook_dim: 512,
        v_vq_dim: 256,
        v_latent_dim: 1024,
        v_hidden: 512,
        v_layers: 8,
        v_heads: 16,
        v_head_dim: 64,
        v_inter: 1024,
        v_quantizers: 16,
        v_codebook_size: 2048,
        v_decoder_dim: 1536,
        v_n_upsampling: 2,
        v_upsampling: [2, 2, 0, 0],
        v_n_rates: 4,
        v_rates: [8, 5, 4, 3, 0, 0, 0, 0],
        v_rms_eps: 1e-5,
        v_rope_theta: 10000.0,
        device,
    }
}

/// Dimension table from a parsed config.json (`TalkerConfig::from_parsed` + `CodePredictorConfig::from_parsed`).
pub fn desc_from_parsed(c: &crate::models::config::ParsedModelConfig, device: i32) -> Result<ffi::q3_model_desc> {
    if c.talker_head_dim != 128 || c.cp_head_dim != 128 {
        bail!("the B200 decode kernels are built for head_dim 128");
    }
    let mut d = default_desc(false, device);
    d.hidden = c.talker_hidden_size as i32;
    d.inter = c.talker_intermediate_size as i32;
    d.layers = c.talker_num_hidden_layers as i32;
    d.heads = c.talker_num_attention_heads as i32;
    d.kv_heads = c.talker_num_key_value_heads as i32;
    d.codec_vocab = c.talker_vocab_size as i32;
    d.text_vocab = c.talker_text_vocab_size as i32;
    d.text_embed_dim = c.talker_text_hidden_size as i32;
    d.rope_theta = c.talker_rope_theta as f32;
    d.rms_eps = c.talker_rms_norm_eps as f32;
    d.cp_hidden = c.cp_hidden_size as i32;
    d.cp_inter = c.cp_intermediate_size as i32;
    d.cp_layers = c.cp_num_hidden_layers as i32;
    d.cp_heads = c.cp_num_attention_heads as i32;
    d.cp_kv_heads = c.cp_num_key_value_heads as i32;
    d.cp_vocab = c.cp_vocab_size as i32;
    d.groups = c.cp_num_code_groups as i32;
    Ok(d)
}

/// Weights resident on one B200.  Immutable after `from_weights`, shareable across threads (`&self` everywhere),
/// like the reference's `Qwen3TTS` model fields.
pub struct B200Model {
    raw: *mut ffi::q3_model,
    desc: ffi::q3_model_desc,
}
unsafe impl Send for B200Model {}
unsafe impl Sync for B200Model {}

impl Drop for B200Model {
    fn drop(&mut self) {
        unsafe { ffi::q3_model_destroy(self.raw) }
    }
}

impl B200Model {
    /// Counterpart of `Qwen3TTS::from_weights` (src/lib.rs:267-274): the two safetensors maps as candle loaded them,
    /// under their HF names.  `talker.*` goes over as bf16, `decoder.*` as f32; everything else (speaker encoder, the
    /// speech tokenizer's encoder) is not part of the decode path and is skipped.
    pub fn from_weights(
        model_weights: &HashMap<String, Tensor>,
        decoder_weights: &HashMap<String, Tensor>,
        desc: ffi::q3_model_desc,
    ) -> Result<Self> {
        let mut raw = ptr::null_mut();
        ffi::check(unsafe { ffi::q3_model_create(&desc, &mut raw) })?;
        let model = Self { raw, desc };
        for (name, t) in model_weights.iter().filter(|(k, _)| k.starts_with("talker.")) {
            model.set_tensor(name, t, DType::BF16)?;
        }
        for (name, t) in decoder_weights.iter().filter(|(k, _)| k.starts_with("decoder.")) {
            model.set_tensor(name, t, DType::F32)?;
        }
        ffi::check(unsafe { ffi::q3_model_finalize(model.raw) })?;
        Ok(model)
    }

    fn set_tensor(&self, name: &str, t: &Tensor, dtype: DType) -> Result<()> {
        let cname = CString::new(name)?;
        let shape: Vec<i64> = if t.rank() == 0 { vec![1] } else { t.dims().iter().map(|&d| d as i64).collect() };
        let flat = t.to_device(&candle_core::Device::Cpu)?.to_dtype(dtype)?.flatten_all()?;
        let code = match dtype {
            DType::BF16 => {
                let v: Vec<half::bf16> = flat.to_vec1()?;
                unsafe { ffi::q3_model_set_tensor(self.raw, cname.as_ptr(), v.as_ptr() as *const _, ffi::Q3_BF16, shape.as_ptr(), shape.len() as i32, 0) }
            }
            _ => {
                let v: Vec<f32> = flat.to_vec1()?;
                unsafe { ffi::q3_model_set_tensor(self.raw, cname.as_ptr(), v.as_ptr() as *const _, ffi::Q3_F32, shape.as_ptr(), shape.len() as i32, 0) }
            }
        };
        ffi::check(code)
    }

    pub fn desc(&self) -> &ffi::q3_model_desc {
        &self.desc
    }

    /// `Decoder12Hz::decode` for one utterance (src/lib.rs:881-890): `[n_frames][16]` -> 24 kHz audio.
    pub fn decode_codes(&self, codes: &FrameCodes) -> Result<AudioBuffer> {
        let t = codes.len();
        let mut tensor = vec![0i64; 16 * t]; // codes_to_tensor layout, src/lib.rs:1417-1431
        for (f, frame) in codes.iter().enumerate() {
            for (q, &c) in frame.iter().enumerate() {
                tensor[q * t + f] = c as i64;
            }
        }
        let mut pcm = vec![0f32; t * SAMPLES_PER_FRAME];
        if t > 0 {
            ffi::check(unsafe { ffi::q3_vocoder_decode(self.raw, tensor.as_ptr(), 1, t as i32, pcm.as_mut_ptr()) })?;
        }
        Ok(AudioBuffer::new(pcm, 24000))
    }
}

/// Position-wise (text id, codec id) pairs of a prompt; -1 = absent.
pub struct Prompt {
    pub text: Vec<i32>,
    pub codec: Vec<i32>,
}

/// `prefill_custom_voice` (src/models/talker.rs:451-488) as id lists; the embedding math runs on the device.
pub fn custom_voice_prompt(input_ids: &[u32], speaker: Speaker, language: Language) -> Prompt {
    let mut text = vec![IM_START, ASSISTANT, NEWLINE, TTS_PAD, TTS_PAD, TTS_PAD, TTS_PAD, TTS_PAD, TTS_BOS];
    let mut codec = vec![
        -1, -1, -1, CODEC_THINK, CODEC_THINK_BOS, language.token_id() as i32, CODEC_THINK_EOS, speaker.token_id() as i32, CODEC_PAD,
    ];
    if let Some(&first) = input_ids.first() {
        text.push(first as i32);
        codec.push(CODEC_BOS);
    }
    Prompt { text, codec }
}

/// `prefill_voice_design` (src/models/talker.rs:585-624).
pub fn voice_design_prompt(input_ids: &[u32], instruct_ids: &[u32], language: Language) -> Prompt {
    let mut text: Vec<i32> = instruct_ids.iter().map(|&t| t as i32).collect();
    let mut codec = vec![-1; instruct_ids.len() + 3];
    text.extend_from_slice(&[IM_START, ASSISTANT, NEWLINE, TTS_PAD, TTS_PAD, TTS_PAD, TTS_PAD, TTS_BOS]);
    codec.extend_from_slice(&[CODEC_THINK, CODEC_THINK_BOS, language.token_id() as i32, CODEC_THINK_EOS, CODEC_PAD]);
    if let Some(&first) = input_ids.first() {
        text.push(first as i32);
        codec.push(CODEC_BOS);
    }
    Prompt { text, codec }
}

/// All mutable per-request state (KV caches, RNG, penalty mask, offsets) and one CUDA stream.  `Send`, not `Sync`,
/// like `StreamingSession<'a>` (src/lib.rs:1484-1485); borrows the model.
pub struct B200Session<'a> {
    raw: *mut ffi::q3_session,
    model: &'a B200Model,
    chunk_frames: usize,
    frames_generated: usize,
    done: bool,
}
unsafe impl Send for B200Session<'_> {}

impl Drop for B200Session<'_> {
    fn drop(&mut self) {
        unsafe { ffi::q3_session_destroy(self.raw) }
    }
}

impl<'a> B200Session<'a> {
    /// One utterance: session + prefill + trailing text (src/lib.rs:743-760; `max_seq = max_length + 256`, :756).
    pub fn new(model: &'a B200Model, prompt: &Prompt, input_ids: &[u32], options: &SynthesisOptions) -> Result<Self> {
        let seed = options.seed.ok_or_else(|| anyhow!("the B200 backend needs SynthesisOptions.seed (reproducible runs only)"))?;
        let cfg = ffi::q3_gen_config {
            max_new_tokens: options.max_length as i32,
            temperature: options.temperature,
            top_k: options.top_k as i32,
            top_p: options.top_p,
            repetition_penalty: options.repetition_penalty,
            eos_token_id: options.eos_token_id.map(|t| t as i32).unwrap_or(-1),
            min_new_tokens: options.min_new_tokens as i32,
            chunk_frames: options.chunk_frames as i32,
        };
        let max_seq = (options.max_length + 256).max(prompt.text.len() + options.max_length) as i32;
        let mut raw = ptr::null_mut();
        ffi::check(unsafe { ffi::q3_session_create(model.raw, 1, max_seq, &cfg, &seed, &mut raw) })?;
        let s = Self { raw, model, chunk_frames: options.chunk_frames.max(1), frames_generated: 0, done: false };
        let len = prompt.text.len() as i32;
        ffi::check(unsafe { ffi::q3_prefill_ids(s.raw, prompt.text.as_ptr(), prompt.codec.as_ptr(), &len, len) })?;
        // build_trailing_text (src/lib.rs:508-519): remaining text tokens, then tts_eos; tts_pad afterwards
        let trailing: Vec<i32> = input_ids.iter().skip(1).map(|&t| t as i32).collect();
        let n = trailing.len() as i32;
        let padded = if trailing.is_empty() { vec![0i32] } else { trailing };
        ffi::check(unsafe { ffi::q3_set_trailing_ids(s.raw, padded.as_ptr(), &n, padded.len() as i32, TTS_EOS, TTS_PAD) })?;
        Ok(s)
    }

    /// Voice-clone session (src/lib.rs:895-1003): `prefill_voice_clone` (src/models/talker.rs:511-564) followed, in ICL
    /// mode, by the streaming overlay of `build_icl_prompt` (talker.rs:646-705) -- passed as ONE causal prefill, which fills the
    /// same KV cache and ends in the same last hidden state / logits as the reference's two chunks.  `speaker_bf16`: the
    /// speaker embedding cast to bf16 bits (lib.rs:930); `ref_codes`: `[T_ref][16]` (ICL) or empty.  The caller applies the
    /// ICL adjustments of the generation config (lib.rs:913-927) to `options` first.
    pub fn new_voice_clone(model: &'a B200Model, input_ids: &[u32], speaker_bf16: &[u16], ref_codes: &[[u32; 16]],
                           ref_text_ids: Option<&[u32]>, language: i32, options: &SynthesisOptions) -> Result<Self> {
        const POS_SPEAKER: i32 = -2; // Q3_POS_SPEAKER
        let pos_ref = |t: usize| -16 - t as i32; // Q3_POS_REF_FRAME(t)
        let mut text = vec![IM_START, ASSISTANT, NEWLINE, TTS_PAD, TTS_PAD, TTS_PAD, TTS_PAD, TTS_PAD, TTS_BOS];
        let mut codec = vec![-1, -1, -1, CODEC_THINK, CODEC_THINK_BOS, language, CODEC_THINK_EOS, POS_SPEAKER, CODEC_PAD];
        let mut trailing: Option<Vec<i32>> = Some(input_ids.iter().skip(1).map(|&t| t as i32).collect());
        if let Some(ref_text) = ref_text_ids {
            let mut all: Vec<i32> = ref_text.iter().chain(input_ids.iter()).map(|&t| t as i32).collect();
            all.push(TTS_EOS);
            let n_codec = ref_codes.len() + 1;
            for i in 0..n_codec {
                text.push(if i < all.len() { all[i] } else { TTS_PAD });
                codec.push(if i == 0 { CODEC_BOS } else { pos_ref(i - 1) });
            }
            // the text that did not fit is the trailing text (it ends in tts_eos, which q3_set_trailing_ids appends);
            // otherwise there are no trailing rows and every frame adds tts_pad
            trailing = if all.len() > n_codec { Some(all[n_codec..all.len() - 1].to_vec()) } else { None };
        } else if let Some(&first) = input_ids.first() {
            text.push(first as i32);
            codec.push(CODEC_BOS);
        }
        let seed = options.seed.ok_or_else(|| anyhow!("the B200 backend needs SynthesisOptions.seed (reproducible runs only)"))?;
        let cfg = ffi::q3_gen_config {
            max_new_tokens: options.max_length as i32,
            temperature: options.temperature,
            top_k: options.top_k as i32,
            top_p: options.top_p,
            repetition_penalty: options.repetition_penalty,
            eos_token_id: options.eos_token_id.map(|t| t as i32).unwrap_or(-1),
            min_new_tokens: options.min_new_tokens as i32,
            chunk_frames: options.chunk_frames as i32,
        };
        let max_seq = (options.max_length + 256).max(text.len() + options.max_length) as i32;
        let mut raw = ptr::null_mut();
        ffi::check(unsafe { ffi::q3_session_create(model.raw, 1, max_seq, &cfg, &seed, &mut raw) })?;
        let s = Self { raw, model, chunk_frames: options.chunk_frames.max(1), frames_generated: 0, done: false };
        let (len, t_ref) = (text.len() as i32, ref_codes.len() as i32);
        let ref_ptr = if ref_codes.is_empty() { ptr::null() } else { ref_codes.as_ptr() as *const u32 };
        ffi::check(unsafe {
            ffi::q3_prefill_voice_clone(s.raw, text.as_ptr(), codec.as_ptr(), &len, len, speaker_bf16.as_ptr(), ref_ptr, &t_ref, t_ref)
        })?;
        let (n, padded) = match trailing {
            Some(t) if !t.is_empty() => (t.len() as i32, t),
            Some(_) => (0, vec![0i32]),
            None => (-1, vec![0i32]),
        };
        ffi::check(unsafe { ffi::q3_set_trailing_ids(s.raw, padded.as_ptr(), &n, padded.len() as i32, TTS_EOS, TTS_PAD) })?;
        Ok(s)
    }

    /// `generate_codes` (src/lib.rs:530-656): the whole loop runs on the device; one read-back at the end.
    pub fn generate(&mut self, max_frames: usize) -> Result<FrameCodes> {
        let mut codes = vec![0u32; max_frames * 16];
        let mut n = 0i32;
        ffi::check(unsafe { ffi::q3_generate(self.raw, max_frames as i32, codes.as_mut_ptr(), &mut n) })?;
        self.frames_generated = n as usize;
        Ok(codes.chunks(16).take(n as usize).map(|c| c.to_vec()).collect())
    }

    /// Vocoder over the frames this session generated.
    pub fn vocode(&mut self, max_frames: usize) -> Result<AudioBuffer> {
        let mut pcm = vec![0f32; max_frames * SAMPLES_PER_FRAME];
        ffi::check(unsafe { ffi::q3_vocode_session(self.raw, max_frames as i32, pcm.as_mut_ptr()) })?;
        pcm.truncate(self.frames_generated * SAMPLES_PER_FRAME);
        Ok(AudioBuffer::new(pcm, 24000))
    }

    /// `StreamingSession::next_chunk` (src/lib.rs:1650-1759).
    pub fn next_chunk(&mut self) -> Result<Option<AudioBuffer>> {
        if self.done {
            return Ok(None);
        }
        let mut codes = vec![0u32; self.chunk_frames * 16];
        let mut pcm = vec![0f32; self.chunk_frames * SAMPLES_PER_FRAME];
        let (mut n, mut done) = (0i32, 0i32);
        ffi::check(unsafe { ffi::q3_stream_next(self.raw, codes.as_mut_ptr(), pcm.as_mut_ptr(), &mut n, &mut done) })?;
        self.frames_generated += n as usize;
        self.done = done != 0;
        if n == 0 {
            return Ok(None);
        }
        pcm.truncate(n as usize * SAMPLES_PER_FRAME);
        Ok(Some(AudioBuffer::new(pcm, 24000)))
    }

    /// Opt-in: `frames` of left context per streamed chunk, -1 = whole history (streamed == non-streamed PCM).
    pub fn set_stream_context(&mut self, frames: i32) -> Result<()> {
        ffi::check(unsafe { ffi::q3_session_set_stream_context(self.raw, frames) })
    }

    /// Opt-in: the first streamed chunk has only `frames` frames (low time to first audio), later ones `chunk_frames`.
    pub fn set_first_chunk(&mut self, frames: i32) -> Result<()> {
        ffi::check(unsafe { ffi::q3_session_set_first_chunk(self.raw, frames) })
    }

    pub fn frames_generated(&self) -> usize {
        self.frames_generated
    }
    pub fn is_done(&self) -> bool {
        self.done
    }
    pub fn model(&self) -> &B200Model {
        self.model
    }

    pub fn timing(&mut self) -> Result<SynthesisTiming> {
        let mut t = ffi::q3_timing::default();
        ffi::check(unsafe { ffi::q3_session_timing(self.raw, &mut t) })?;
        Ok(SynthesisTiming {
            prefill_ms: t.prefill_ms as f64,
            generation_ms: t.generation_ms as f64,
            generation_frames: t.generation_frames as usize,
            decode_ms: t.decode_ms as f64,
        })
    }
}

/// Body of `Qwen3TTS::synthesize_with_timing` (src/lib.rs:425-501) on the B200 backend, after tokenisation:
/// `let input_ids = self.text_tokenizer.encode(text)?;` stays where it is, then
/// `b200::synthesize_with_voice(&self.b200, &input_ids, speaker, language, &options)`.
pub fn synthesize_with_voice(
    model: &B200Model,
    input_ids: &[u32],
    speaker: Speaker,
    language: Language,
    options: &SynthesisOptions,
) -> Result<(AudioBuffer, FrameCodes, SynthesisTiming)> {
    let prompt = custom_voice_prompt(input_ids, speaker, language);
    let mut s = B200Session::new(model, &prompt, input_ids, options)?;
    let codes = s.generate(options.max_length)?;
    let audio = s.vocode(options.max_length)?;
    let timing = s.timing()?;
    Ok((audio, codes, timing))
}

/// `synthesize_voice_design` (src/lib.rs:802-870) after tokenising the text and the voice description.
pub fn synthesize_voice_design(
    model: &B200Model,
    input_ids: &[u32],
    instruct_ids: &[u32],
    language: Language,
    options: &SynthesisOptions,
) -> Result<(AudioBuffer, FrameCodes, SynthesisTiming)> {
    let prompt = voice_design_prompt(input_ids, instruct_ids, language);
    let mut s = B200Session::new(model, &prompt, input_ids, options)?;
    let codes = s.generate(options.max_length)?;
    let audio = s.vocode(options.max_length)?;
    let timing = s.timing()?;
    Ok((audio, codes, timing))
}

/// `synthesize_streaming` (src/lib.rs:1070-1093): returns the session; call `next_chunk` until `None`.
pub fn synthesize_streaming<'a>(
    model: &'a B200Model,
    input_ids: &[u32],
    speaker: Speaker,
    language: Language,
    options: &SynthesisOptions,
) -> Result<B200Session<'a>> {
    B200Session::new(model, &custom_voice_prompt(input_ids, speaker, language), input_ids, options)
}
