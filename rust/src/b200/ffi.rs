//! `extern "C"` declarations of libq3tts_b200.so, one to one with include/q3tts.h (ABI version 1).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct q3_model {
    _p: [u8; 0],
}
#[repr(C)]
pub struct q3_session {
    _p: [u8; 0],
}

pub const Q3_OK: c_int = 0;
pub const Q3_ERR_INVALID: c_int = 1;
pub const Q3_ERR_CUDA: c_int = 2;
pub const Q3_ERR_KV_OVERFLOW: c_int = 3;
pub const Q3_ERR_MISSING_WEIGHT: c_int = 4;
pub const Q3_ERR_STATE: c_int = 5;
pub const Q3_ERR_UNSUPPORTED: c_int = 6;
pub const Q3_BF16: c_int = 0;
pub const Q3_F32: c_int = 1;

/// TalkerConfig (src/models/talker.rs:208-274), CodePredictorConfig (src/models/code_predictor.rs:48-113),
/// Decoder12HzConfig (src/models/codec/decoder_12hz.rs:47-67).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct q3_model_desc {
    pub hidden: i32,
    pub inter: i32,
    pub layers: i32,
    pub heads: i32,
    pub kv_heads: i32,
    pub head_dim: i32,
    pub codec_vocab: i32,
    pub text_vocab: i32,
    pub text_embed_dim: i32,
    pub rope_theta: f32,
    pub rms_eps: f32,
    pub cp_hidden: i32,
    pub cp_inter: i32,
    pub cp_layers: i32,
    pub cp_heads: i32,
    pub cp_kv_heads: i32,
    pub cp_vocab: i32,
    pub groups: i32,
    pub cp_rope_positions: i32,
    pub cp_max_seq: i32,
    pub v_codebook_dim: i32,
    pub v_vq_dim: i32,
    pub v_latent_dim: i32,
    pub v_hidden: i32,
    pub v_layers: i32,
    pub v_heads: i32,
    pub v_head_dim: i32,
    pub v_inter: i32,
    pub v_quantizers: i32,
    pub v_codebook_size: i32,
    pub v_decoder_dim: i32,
    pub v_n_upsampling: i32,
    pub v_upsampling: [i32; 4],
    pub v_n_rates: i32,
    pub v_rates: [i32; 8],
    pub v_rms_eps: f32,
    pub v_rope_theta: f32,
    pub device: i32,
}

/// GenerationConfig (src/generation/sampling.rs:100-115) + SynthesisOptions.chunk_frames (src/lib.rs:1786-1805).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct q3_gen_config {
    pub max_new_tokens: i32,
    pub temperature: f64,
    pub top_k: i32,
    pub top_p: f64,
    pub repetition_penalty: f64,
    pub eos_token_id: i32, // -1 = None
    pub min_new_tokens: i32,
    pub chunk_frames: i32,
}

/// SynthesisTiming (src/lib.rs:136-147).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct q3_timing {
    pub prefill_ms: f32,
    pub generation_ms: f32,
    pub decode_ms: f32,
    pub generation_frames: i32,
}

#[link(name = "q3tts_b200")]
extern "C" {
    pub fn q3_last_error() -> *const c_char;
    pub fn q3_abi_version() -> c_int;
    pub fn q3_kernel_launch_count() -> u64;

    pub fn q3_model_create(desc: *const q3_model_desc, out: *mut *mut q3_model) -> c_int;
    pub fn q3_model_set_tensor(
        m: *mut q3_model,
        hf_name: *const c_char,
        data: *const c_void,
        dtype: c_int,
        shape: *const i64,
        ndim: i32,
        on_device: i32,
    ) -> c_int;
    pub fn q3_model_finalize(m: *mut q3_model) -> c_int;
    pub fn q3_model_destroy(m: *mut q3_model);

    pub fn q3_session_create(
        m: *const q3_model,
        batch: i32,
        max_seq: i32,
        cfg: *const q3_gen_config,
        seeds: *const u64,
        out: *mut *mut q3_session,
    ) -> c_int;
    pub fn q3_session_reset(s: *mut q3_session, seeds: *const u64) -> c_int;
    pub fn q3_session_destroy(s: *mut q3_session);
    pub fn q3_session_stream(s: *mut q3_session) -> *mut c_void;
    pub fn q3_session_synchronize(s: *mut q3_session) -> c_int;
    pub fn q3_session_set_stream_context(s: *mut q3_session, left_context_frames: i32) -> c_int;
    pub fn q3_session_set_first_chunk(s: *mut q3_session, first_chunk_frames: i32) -> c_int;
    pub fn q3_session_timing(s: *mut q3_session, out: *mut q3_timing) -> c_int;

    pub fn q3_prefill_embeds(s: *mut q3_session, embeds: *const u16, lens: *const i32, l_max: i32) -> c_int;
    pub fn q3_prefill_ids(s: *mut q3_session, text_ids: *const i32, codec_ids: *const i32, lens: *const i32, l_max: i32) -> c_int;
    pub fn q3_speaker_embed_dim(m: *const q3_model) -> i32;
    pub fn q3_speaker_encode(m: *const q3_model, mel: *const f32, batch: i32, t: i32, embed_out: *mut f32) -> c_int;
    pub fn q3_prefill_voice_clone(s: *mut q3_session, text_ids: *const i32, codec_ids: *const i32, lens: *const i32, l_max: i32,
                                  speaker_embeds: *const u16, ref_codes: *const u32, t_ref: *const i32, t_ref_max: i32) -> c_int;
    pub fn q3_set_trailing_text(s: *mut q3_session, trailing: *const u16, lt: *const i32, lt_max: i32, tts_pad: *const u16) -> c_int;
    pub fn q3_set_trailing_ids(s: *mut q3_session, ids: *const i32, n: *const i32, n_max: i32, tts_eos_id: i32, tts_pad_id: i32) -> c_int;

    pub fn q3_generate(s: *mut q3_session, max_frames: i32, codes: *mut u32, n_frames: *mut i32) -> c_int;
    pub fn q3_generate_async(s: *mut q3_session, max_frames: i32) -> c_int;
    pub fn q3_get_codes(s: *mut q3_session, max_frames: i32, codes: *mut u32, n_frames: *mut i32) -> c_int;
    pub fn q3_stream_next(s: *mut q3_session, codes: *mut u32, pcm: *mut f32, n_frames: *mut i32, done: *mut i32) -> c_int;

    pub fn q3_vocoder_decode(m: *const q3_model, codes: *const i64, batch: i32, t: i32, pcm: *mut f32) -> c_int;
    pub fn q3_vocode_session(s: *mut q3_session, max_frames: i32, pcm: *mut f32) -> c_int;

    pub fn q3_talker_step(s: *mut q3_session, step_input: *const u16, hidden_out: *mut u16, logits_out: *mut f32) -> c_int;
    pub fn q3_code_predictor_frame(
        s: *mut q3_session,
        last_hidden: *const u16,
        sem_tokens: *const u32,
        codes_out: *mut u32,
        logits_out: *mut f32,
    ) -> c_int;
    pub fn q3_sample(
        m: *const q3_model,
        logits: *const f32,
        batch: i32,
        vocab: i32,
        cfg: *const q3_gen_config,
        rng_states: *mut u64,
        seen_mask: *mut u8,
        token_count: i32,
        tokens_out: *mut u32,
    ) -> c_int;
    pub fn q3_fused_residual_rmsnorm(
        x: *const c_void,
        r: *const c_void,
        w: *const c_void,
        out_normed: *mut c_void,
        out_sum: *mut c_void,
        rows: i32,
        cols: i32,
        eps: f32,
        dtype: c_int,
        stream: *mut c_void,
    ) -> c_int;
    pub fn q3_fused_residual_rmsnorm_host(
        x: *const c_void,
        r: *const c_void,
        w: *const c_void,
        out_normed: *mut c_void,
        out_sum: *mut c_void,
        rows: i32,
        cols: i32,
        eps: f32,
        dtype: c_int,
        device: i32,
    ) -> c_int;
}

/// Status code -> `anyhow::Result`, carrying the library's thread-local message (same convention as the crate).
pub fn check(code: c_int) -> anyhow::Result<()> {
    if code == Q3_OK {
        return Ok(());
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(q3_last_error()) }.to_string_lossy().into_owned();
    anyhow::bail!("q3tts_b200 error {code}: {msg}")
}
