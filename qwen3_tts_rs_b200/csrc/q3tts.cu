// libq3tts_b200: model / session management, the decode loop and the C ABI (include/q3tts.h).
// ref: src/lib.rs (generate_codes 530-656, StreamingSession 1484-1782), src/models/talker.rs,
// src/models/code_predictor.rs, src/generation/sampling.rs.
#include <algorithm>
#include <cmath>
#include <cstring>

#include <cstdlib>

#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges show up under nsys / ncu --nvtx, no-ops otherwise

#include "decode_kernels.cuh"
#include "gemm_tc.cuh"
#include "gemv.cuh"
#include "mega.cuh"
#include "mega2.cuh"
#include "mega3.cuh"
#include "mega4.cuh"
#include "mega5.cuh"
#include "model.h"

std::atomic<uint64_t> g_q3_launches{0};

// Stage ranges named after the reference's tracing spans (src/lib.rs:433-485 "synthesize" / "prefill" / "decode",
// 578 "generate_frames"): the per-frame spans of the reference (code_predictor / talker_step / sampling, lib.rs:593-637)
// live INSIDE one persistent kernel launch here and are timed by the kernel's own stamps (tools/profile_mega2.py).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
static thread_local std::string g_last_error;

// =================================================================================================
// session object
struct Scratch {
  DBuf x, qkv, q, attn, o, normed, h1, act, tmp_e, tmp_p, zeros;
  int tcap = 0;
};

static bool use_tc_gemm() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("Q3_TC");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

struct q3_session {
  const q3_model* m = nullptr;
  int B = 0, max_seq = 0, frames_cap = 0;
  q3_gen_config cfg{};
  cudaStream_t st = nullptr;
  // caches
  DBuf tk_k, tk_v, cp_k, cp_v, cos_tab, sin_tab;
  // per-row state (device)
  DBuf cur_tok, done, n_frames, token_count, offset, frame_idx, rng, seen, last_hidden, trailing, lt, tts_pad, codes,
      amax, frame_codes, logits, step_input, cp_x0, cp_xe, cp_logits, lens_dev;
  int* host_flags = nullptr;       // mapped pinned
  int* host_flags_dev = nullptr;
  int lt_max = 1;
  Scratch sc;
  FrameState fs{};
  // host mirrors
  std::vector<int> prefill_len;
  int frames_run = 0;              // loop iterations executed since prefill
  bool prefilled = false, first_sampled = false;
  // persistent frame kernel (batch <= 8)
  bool use_mega = false;
  int mega_ver = 2;                // 1: fence-based grid barriers (mega.cuh), 2: tagged dataflow phases (mega2.cuh)
  DBuf tr_ids, tr_proj;            // staging of q3_set_trailing_ids
  int stream_first = 0;            // q3_session_set_first_chunk: frames of the first streamed chunk (0 = chunk_frames)
  int split_thr = 0, split_ns = 4; // split-KV attention: context length from which a row is cut over split_ns CTAs (0 = never)
  DBuf pf_tid, pf_cid;             // staging of q3_prefill_ids
  DBuf pf_spk, pf_ref;             // q3_prefill_voice_clone: speaker embeddings, reference codes
  DBuf m2_x, m2_qkv, m2_attn, m2_h1, m2_act, m2_prog, m2_prog_tmp, m2_tag, m2_xchg;
  std::vector<DBuf> m2_progs;      // cached full-frame program of every row group
  std::vector<int> m2_group_nph;
  int m2_n_ph = 0;                 // phases of the cached full-frame program (0: not built)
  size_t m2_smem = 0, m3_smem = 0, m4_smem = 0, m5_smem = 0;
  int m4_slots = 0, m4_red2 = 0, m5_slots = 0;
  DBuf prof;                       // optional timestamp buffer (q3_debug_profile)
  float* tap_cp_logits = nullptr;  // q3_debug_generate_tapped: device [15][B][cp_vocab] written by every frame's CP heads
  size_t mega_smem = 0;
  int mega_grid = 0;
  DBuf bar, ss;
  // frame graph
  cudaGraphExec_t graph_exec = nullptr;
  bool graph_ok = false;
  uint64_t graph_launches = 0;     // kernels inside one replay of the frame graph
  // streaming
  std::vector<int> stream_emitted;
  int stream_left_ctx = 0;         // frames of left context re-decoded per chunk (0 = the reference's stateless chunks)
  std::unique_ptr<VocoderScratch> voc;   // borrowed from the model's pool at the first vocode, returned on destruction
  std::unique_ptr<VocoderStreamState> vstream;   // carried vocoder state of a stateful streamed decode (stream context -1)
  // timing
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_poll[2] = {nullptr, nullptr};
  q3_timing timing{};
  ~q3_session() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (auto& e : ev_poll) if (e) cudaEventDestroy(e);
    if (host_flags) cudaFreeHost(host_flags);
    if (st) cudaStreamDestroy(st);
    if (m && voc) m->release_scratch(std::move(voc));
  }
};

// =================================================================================================
// helpers
static const RawTensor& need_bf16(const q3_model* m, const std::string& name) {
  auto it = m->t.find(name);
  if (it == m->t.end()) throw Q3Error(Q3_ERR_MISSING_WEIGHT, "Missing weight: " + name);
  if (it->second.dtype != Q3_BF16) throw Q3Error(Q3_ERR_INVALID, "talker weight must be stored bf16: " + name);
  return it->second;
}
static const bf16* needb(const q3_model* m, const std::string& name) { return need_bf16(m, name).buf.as<bf16>(); }

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = f2bf(in[i]);
}
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = bf2f(in[i]);
}
__global__ void add_int_kernel(int* p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] += v;
}

// The frame graph bakes in buffer addresses and FrameState scalars; drop it when any of them changes.
static void invalidate_graph(q3_session* s) {
  if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
  s->graph_exec = nullptr;
  s->graph_ok = false;
}

static void ensure_scratch(q3_session* s, int T) {
  Scratch& sc = s->sc;
  if (T <= sc.tcap) return;
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  invalidate_graph(s);
  const q3_model_desc& d = s->m->d;
  const int H = std::max(d.hidden, d.cp_hidden), I = std::max(d.inter, d.cp_inter);
  const int nh = std::max(d.heads + 2 * d.kv_heads, d.cp_heads + 2 * d.cp_kv_heads);
  const int qd = std::max(d.heads, d.cp_heads) * 128;
  const int E = std::max(d.text_embed_dim, H);
  sc.x.alloc((size_t)T * H * 2);
  sc.qkv.alloc((size_t)T * nh * 128 * 2);
  sc.q.alloc((size_t)T * qd * 2);
  sc.attn.alloc((size_t)T * qd * 2);
  sc.o.alloc((size_t)T * H * 2);
  sc.normed.alloc((size_t)T * H * 2);
  sc.h1.alloc((size_t)T * H * 2);
  sc.act.alloc((size_t)T * I * 2);
  sc.tmp_e.alloc((size_t)T * E * 2);
  sc.tmp_p.alloc((size_t)T * E * 2);
  sc.zeros.alloc((size_t)T * H * 2);
  sc.zeros.zero(s->st);
  sc.tcap = T;
}

// One decoder stack (talker or code predictor) over T = B*S tokens, in place on x.
// ref: DecoderLayer::forward (transformer.rs:442-467) x layers.
static void layers_forward(q3_session* s, const std::vector<LayerW>& L, const StackDims& dm, bf16* x, int T, int S,
                           const int* pos_base, int pos_add, bf16* kc, bf16* vc, int cache_seq, const bf16* cos_tab,
                           const bf16* sin_tab) {
  const q3_model* m = s->m;
  Scratch& sc = s->sc;
  const int nh = dm.heads + 2 * dm.kv_heads;
  const float eps = m->d.rms_eps;
  const size_t layer_stride = (size_t)s->B * dm.kv_heads * cache_seq * 128;
  // multi-position passes (prompt prefill, S > 1) run their projections on the tcgen05 + TMA GEMM -- for every batch
  // size, so that a row's result never depends on how many other rows share the launch
  const bool tc = use_tc_gemm() && S > 1 && gemm_tc_supported(nh * 128, dm.H, T, dm.H, EPI_STORE) &&
                  gemm_tc_supported(dm.H, dm.heads * 128, T, dm.heads * 128, EPI_STORE) &&
                  gemm_tc_supported(dm.I, dm.H, T, dm.H, EPI_SWIGLU) && gemm_tc_supported(dm.H, dm.I, T, dm.I, EPI_RESIDUAL);
  for (int l = 0; l < dm.layers; ++l) {
    const LayerW& w = L[l];
    if (tc) {
      // input RMSNorm (rms_norm(x) == fused kernel with a zero residual), then [q;k;v] = xn W^T
      fused_residual_rmsnorm_launch<bf16>(x, sc.zeros.as<bf16>(), w.in_ln, sc.normed.as<bf16>(), sc.h1.as<bf16>(), T, dm.H, eps, s->st);
      GemmTcArgs ga{};
      ga.N = nh * 128; ga.K = dm.H; ga.T = T; ga.epi = EPI_STORE; ga.Y = sc.qkv.as<bf16>(); ga.ldy = nh * 128;
      gemm_tc_launch(w.wqkv, nullptr, sc.normed.as<bf16>(), dm.H, ga, s->st);
    } else {
      GemvArgs g{};
      g.W = w.wqkv; g.X = x; g.ldx = dm.H; g.norm_w = w.in_ln; g.eps = eps; g.N = nh * 128; g.K = dm.H; g.T = T;
      g.pro = PRO_RMSNORM; g.epi = EPI_STORE; g.Y = sc.qkv.as<bf16>(); g.ldy = nh * 128;
      gemv_launch(g, m->num_sms, s->st);
    }

    RopeArgs r{};
    r.qkv = sc.qkv.as<bf16>(); r.q_out = sc.q.as<bf16>();
    r.k_cache = kc + l * layer_stride; r.v_cache = vc + l * layer_stride;
    r.q_norm_w = w.q_norm; r.k_norm_w = w.k_norm; r.cos_tab = cos_tab; r.sin_tab = sin_tab;
    r.pos_base = pos_base; r.pos_add = pos_add; r.S = S; r.T = T; r.heads = dm.heads; r.kv_heads = dm.kv_heads;
    r.max_seq = cache_seq; r.eps = eps;
    qk_norm_rope_append_kernel<<<dim3(ceil_div(nh, 4), T), 128, 0, s->st>>>(r);
    Q3_COUNT_LAUNCH();
    Q3_LAUNCH_CHECK();

    AttnArgs a{};
    a.q = sc.q.as<bf16>(); a.k_cache = r.k_cache; a.v_cache = r.v_cache; a.out = sc.attn.as<bf16>();
    a.pos_base = pos_base; a.pos_add = pos_add; a.S = S; a.T = T; a.heads = dm.heads; a.kv_heads = dm.kv_heads;
    a.max_seq = cache_seq;
    attn_decode_kernel<<<dim3(dm.kv_heads, T), 256, attn_smem_bytes(cache_seq), s->st>>>(a);
    Q3_COUNT_LAUNCH();
    Q3_LAUNCH_CHECK();

    if (tc) {
      GemmTcArgs ga{};
      ga.N = dm.H; ga.K = dm.heads * 128; ga.T = T; ga.epi = EPI_STORE; ga.Y = sc.o.as<bf16>(); ga.ldy = dm.H;
      gemm_tc_launch(w.wo, nullptr, sc.attn.as<bf16>(), dm.heads * 128, ga, s->st);
    } else {
      GemvArgs o{};
      o.W = w.wo; o.X = sc.attn.as<bf16>(); o.ldx = dm.heads * 128; o.N = dm.H; o.K = dm.heads * 128; o.T = T;
      o.pro = PRO_NONE; o.epi = EPI_STORE; o.Y = sc.o.as<bf16>(); o.ldy = dm.H;
      gemv_launch(o, m->num_sms, s->st);
    }

    fused_residual_rmsnorm_launch<bf16>(sc.o.as<bf16>(), x, w.post_ln, sc.normed.as<bf16>(), sc.h1.as<bf16>(), T, dm.H,
                                        eps, s->st);

    if (tc) {
      GemmTcArgs gu{};
      gu.N = dm.I; gu.K = dm.H; gu.T = T; gu.epi = EPI_SWIGLU; gu.Y = sc.act.as<bf16>(); gu.ldy = dm.I;
      gemm_tc_launch(w.gate, w.up, sc.normed.as<bf16>(), dm.H, gu, s->st);
      GemmTcArgs dn{};
      dn.N = dm.H; dn.K = dm.I; dn.T = T; dn.epi = EPI_RESIDUAL; dn.R = sc.h1.as<bf16>(); dn.ldr = dm.H; dn.Y = x; dn.ldy = dm.H;
      gemm_tc_launch(w.down, nullptr, sc.act.as<bf16>(), dm.I, dn, s->st);
    } else {
      GemvArgs gu{};
      gu.W = w.gate; gu.W2 = w.up; gu.X = sc.normed.as<bf16>(); gu.ldx = dm.H; gu.N = dm.I; gu.K = dm.H; gu.T = T;
      gu.pro = PRO_NONE; gu.epi = EPI_SWIGLU; gu.Y = sc.act.as<bf16>(); gu.ldy = dm.I;
      gemv_launch(gu, m->num_sms, s->st);

      GemvArgs dn{};
      dn.W = w.down; dn.X = sc.act.as<bf16>(); dn.ldx = dm.I; dn.N = dm.H; dn.K = dm.I; dn.T = T;
      dn.pro = PRO_NONE; dn.epi = EPI_RESIDUAL; dn.R = sc.h1.as<bf16>(); dn.ldr = dm.H; dn.Y = x; dn.ldy = dm.H;
      gemv_launch(dn, m->num_sms, s->st);
    }
  }
}

// text_proj(text_embedding[ids]) for n ids (device int array) -> out [n][hidden]   (talker.rs:294-321)
static void text_project(q3_session* s, const int* ids_dev, int n, bf16* out) {
  const q3_model* m = s->m;
  const q3_model_desc& d = m->d;
  ensure_scratch(s, n);
  gather_rows_kernel<<<n, 128, 0, s->st>>>(m->text_emb, ids_dev, d.text_embed_dim, d.text_vocab, s->sc.tmp_e.as<bf16>());
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
  const int E = d.text_embed_dim;
  if (use_tc_gemm() && gemm_tc_supported(E, E, n, E, EPI_BIAS_SILU) && gemm_tc_supported(d.hidden, E, n, E, EPI_BIAS)) {
    GemmTcArgs a{};
    a.N = E; a.K = E; a.T = n; a.epi = EPI_BIAS_SILU; a.bias = m->fc1_b; a.Y = s->sc.tmp_p.as<bf16>(); a.ldy = E;
    gemm_tc_launch(m->fc1_w, nullptr, s->sc.tmp_e.as<bf16>(), E, a, s->st);
    GemmTcArgs b{};
    b.N = d.hidden; b.K = E; b.T = n; b.epi = EPI_BIAS; b.bias = m->fc2_b; b.Y = out; b.ldy = d.hidden;
    gemm_tc_launch(m->fc2_w, nullptr, s->sc.tmp_p.as<bf16>(), E, b, s->st);
    return;
  }
  GemvArgs a{};
  a.W = m->fc1_w; a.bias = m->fc1_b; a.X = s->sc.tmp_e.as<bf16>(); a.ldx = E; a.N = E;
  a.K = E; a.T = n; a.pro = PRO_NONE; a.epi = EPI_BIAS_SILU; a.Y = s->sc.tmp_p.as<bf16>(); a.ldy = E;
  gemv_launch(a, m->num_sms, s->st);
  GemvArgs b{};
  b.W = m->fc2_w; b.bias = m->fc2_b; b.X = s->sc.tmp_p.as<bf16>(); b.ldx = E; b.N = d.hidden;
  b.K = E; b.T = n; b.pro = PRO_NONE; b.epi = EPI_BIAS; b.Y = out; b.ldy = d.hidden;
  gemv_launch(b, m->num_sms, s->st);
}

static SampleArgs make_sample_args(const q3_gen_config& c, int V, int B) {
  SampleArgs a{};
  a.V = V; a.B = B;
  a.use_temp = (c.temperature != 1.0 && c.temperature > 0.0) ? 1 : 0;      // sampling.rs:148
  a.inv_temp = (float)(1.0 / c.temperature);
  a.greedy = c.temperature < 0.01 ? 1 : 0;                                 // sampling.rs:155
  a.top_k = c.top_k;
  a.use_top_p = (c.top_p < 1.0 && c.top_p > 0.0) ? 1 : 0;                  // sampling.rs:167
  a.top_p = (float)c.top_p;
  a.use_pen = (c.repetition_penalty != 1.0 && std::fabs(c.repetition_penalty - 1.0) >= 1e-9) ? 1 : 0;
  a.pen = (float)c.repetition_penalty;
  a.inv_pen = 1.0f / (float)c.repetition_penalty;                          // sampling.rs:389-390 (f32 divide)
  a.eos = c.eos_token_id;
  a.min_new_tokens = c.min_new_tokens;
  return a;
}

static void launch_sampler(q3_session* s, int advance) {
  SampleArgs a = make_sample_args(s->cfg, s->m->d.codec_vocab, s->B);
  a.logits = s->logits.as<float>();
  a.seen = s->fs.seen; a.rng = s->fs.rng; a.tok_out = s->fs.cur_tok; a.token_count = s->fs.token_count;
  a.done = s->fs.done; a.offset = s->fs.offset; a.frame_idx = s->fs.frame_idx; a.host_flags = nullptr;
  a.advance = advance;
  sample_kernel<<<s->B, 1024, 0, s->st>>>(a);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
}

// talker final norm + codec head; writes last_hidden (post-norm) and f32 logits (talker.rs:730-733)
static void talker_head(q3_session* s, const bf16* x, int ldx) {
  const q3_model* m = s->m;
  GemvArgs h{};
  h.W = m->codec_head; h.X = x; h.ldx = ldx; h.norm_w = m->t_norm; h.eps = m->d.rms_eps; h.xn_out = s->fs.last_hidden;
  h.N = m->d.codec_vocab; h.K = m->d.hidden; h.T = s->B; h.pro = PRO_RMSNORM; h.epi = EPI_LOGITS; h.Yf = s->logits.as<float>();
  gemv_launch(h, m->num_sms, s->st);
}

// The code-predictor part of one frame (code_predictor.rs:320-416): 15 dependent passes.
static void cp_frame(q3_session* s, float* logits_out /* [15][B][cp_vocab] or null */) {
  const q3_model* m = s->m;
  const q3_model_desc& d = m->d;
  const int B = s->B, n_ac = d.groups - 1, H = d.hidden, C = d.cp_hidden;
  ensure_scratch(s, 2 * B);
  cp_begin_kernel<<<B, 256, 0, s->st>>>(s->fs, m->codec_emb, s->cp_x0.as<bf16>(), H, B, n_ac);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
  bf16* x = s->sc.x.as<bf16>();
  auto project = [&](const bf16* src, int T) {
    if (m->cp_proj_w) {
      GemvArgs p{};
      p.W = m->cp_proj_w; p.bias = m->cp_proj_b; p.X = src; p.ldx = H; p.N = C; p.K = H; p.T = T;
      p.pro = PRO_NONE; p.epi = EPI_BIAS; p.Y = x; p.ldy = C;
      gemv_launch(p, m->num_sms, s->st);
    } else {
      Q3_CHECK_CUDA(cudaMemcpyAsync(x, src, (size_t)T * C * 2, cudaMemcpyDeviceToDevice, s->st));
    }
  };
  auto head = [&](int g, const bf16* xin, int ldx) {
    GemvArgs h{};
    h.W = m->cp_head[g]; h.X = xin; h.ldx = ldx; h.norm_w = m->cp_norm; h.eps = d.rms_eps; h.N = d.cp_vocab; h.K = C;
    h.T = B; h.pro = PRO_RMSNORM; h.epi = EPI_LOGITS; h.amax = s->fs.amax + (size_t)g * B;
    h.Yf = logits_out ? logits_out + (size_t)g * B * d.cp_vocab : nullptr;
    gemv_launch(h, m->num_sms, s->st);
  };
  // pass 0: two positions per row [talker_hidden, semantic_embed], causal, offset 0
  project(s->cp_x0.as<bf16>(), 2 * B);
  layers_forward(s, m->cl, m->cdims(), x, 2 * B, 2, nullptr, 0, s->cp_k.as<bf16>(), s->cp_v.as<bf16>(), d.cp_max_seq,
                 m->cp_cos, m->cp_sin);
  head(0, x + C, 2 * C);                      // logits from position 1 of every row
  for (int g = 1; g < n_ac; ++g) {
    cp_embed_kernel<<<B, 256, 0, s->st>>>(s->fs, m->cp_emb[g - 1], s->cp_xe.as<bf16>(), H, B, g);
    Q3_COUNT_LAUNCH();
    Q3_LAUNCH_CHECK();
    project(s->cp_xe.as<bf16>(), B);
    layers_forward(s, m->cl, m->cdims(), x, B, 1, nullptr, g + 1, s->cp_k.as<bf16>(), s->cp_v.as<bf16>(), d.cp_max_seq,
                   m->cp_cos, m->cp_sin);
    head(g, x, C);
  }
}

// everything of one loop iteration of generate_codes (lib.rs:580-652)
static void frame_body(q3_session* s) {
  const q3_model* m = s->m;
  const q3_model_desc& d = m->d;
  cp_frame(s, s->tap_cp_logits);
  EmbTable tab{};
  for (int i = 0; i < d.groups - 1; ++i) tab.e[i] = m->cp_emb[i];
  frame_finish_kernel<<<s->B, 256, 0, s->st>>>(s->fs, tab, m->codec_emb, s->step_input.as<bf16>(), d.hidden, s->B,
                                               d.groups - 1);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
  bf16* x = s->sc.x.as<bf16>();
  Q3_CHECK_CUDA(cudaMemcpyAsync(x, s->step_input.p, (size_t)s->B * d.hidden * 2, cudaMemcpyDeviceToDevice, s->st));
  layers_forward(s, m->tl, m->tdims(), x, s->B, 1, s->fs.offset, 0, s->tk_k.as<bf16>(), s->tk_v.as<bf16>(), s->max_seq,
                 s->cos_tab.as<bf16>(), s->sin_tab.as<bf16>());
  talker_head(s, x, d.hidden);
  launch_sampler(s, 1);
}

// ---- persistent frame kernel ----------------------------------------------------------------------------
static MegaArgs mega_args(q3_session* s) {
  const q3_model* m = s->m;
  const q3_model_desc& d = m->d;
  MegaArgs a{};
  a.tk = {m->tl_dev, d.layers, d.hidden, d.inter, d.heads, d.kv_heads};
  a.cp = {m->cl_dev, d.cp_layers, d.cp_hidden, d.cp_inter, d.cp_heads, d.cp_kv_heads};
  a.codec_emb = m->codec_emb; a.t_norm = m->t_norm; a.codec_head = m->codec_head;
  a.cp_proj_w = m->cp_proj_w; a.cp_proj_b = m->cp_proj_b; a.cp_norm = m->cp_norm;
  for (int i = 0; i < 15; ++i) { a.cp_emb[i] = m->cp_emb[i]; a.cp_head[i] = m->cp_head[i]; }
  a.cp_cos = m->cp_cos; a.cp_sin = m->cp_sin; a.t_cos = s->cos_tab.as<bf16>(); a.t_sin = s->sin_tab.as<bf16>();
  a.H = d.hidden; a.C = d.cp_hidden; a.V = d.codec_vocab; a.cpV = d.cp_vocab; a.n_ac = d.groups - 1; a.B = s->B;
  a.eps = d.rms_eps;
  a.fs = s->fs;
  a.tk_k = s->tk_k.as<bf16>(); a.tk_v = s->tk_v.as<bf16>(); a.cp_k = s->cp_k.as<bf16>(); a.cp_v = s->cp_v.as<bf16>();
  a.max_seq = s->max_seq; a.cp_max_seq = d.cp_max_seq;
  a.x = s->sc.x.as<bf16>(); a.qkv = s->sc.qkv.as<bf16>(); a.attn = s->sc.attn.as<bf16>(); a.o = s->sc.o.as<bf16>();
  a.h1 = s->sc.h1.as<bf16>(); a.act = s->sc.act.as<bf16>(); a.step_input = s->step_input.as<bf16>();
  a.logits = s->logits.as<float>();
  a.bar = s->bar.as<unsigned>();
  a.ssA = s->ss.as<float>();
  a.ssB = s->ss.as<float>() + (size_t)s->mega_grid * MEGA_TMAX;
  SampleArgs sa = make_sample_args(s->cfg, d.codec_vocab, s->B);
  sa.logits = s->logits.as<float>();
  sa.seen = s->fs.seen; sa.rng = s->fs.rng; sa.tok_out = s->fs.cur_tok; sa.token_count = s->fs.token_count;
  sa.done = s->fs.done; sa.offset = s->fs.offset; sa.frame_idx = s->fs.frame_idx; sa.host_flags = nullptr; sa.advance = 1;
  a.smp = sa;
  const char* e1 = std::getenv("Q3_BAR_MODE");
  const char* e2 = std::getenv("Q3_PREFETCH");
  a.bar_mode = e1 ? std::atoi(e1) : 0;
  a.prefetch_mode = e2 ? std::atoi(e2) : 3;
  const char* e3 = std::getenv("Q3_SMALL");
  a.small_mode = e3 ? std::atoi(e3) : 1;
  return a;
}

static void mega_launch(q3_session* s, MegaArgs& a) {
  const unsigned magic = (unsigned)(((1ull << 24) + (unsigned)s->mega_grid - 1) / (unsigned)s->mega_grid);
  Q3_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_grid_magic, &magic, 4, 0, cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaMemsetAsync(s->bar.p, 0, 4, s->st));
  if (s->prof.p) { a.prof = s->prof.as<unsigned long long>(); a.prof_cap = (int)(s->prof.bytes / 8); }
  void* params[] = {(void*)&a};
  Q3_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)decode_frames_mega_kernel, dim3(s->mega_grid), dim3(MEGA_THREADS), params,
                                            s->mega_smem, s->st));
  Q3_COUNT_LAUNCH();
}


// ---- persistent frame kernel, dataflow generation (mega2.cuh) ---------------------------------------------
// The phase program of one frame for the given mode; every pointer is resolved here, once.
// Rows [r0, r0 + Bg) of the session form one launch group (batches above 16 run as groups of 16, back to back).
static std::vector<M2Phase> m2_build_program(q3_session* s, bool do_cp, bool do_finish, bool do_talker, bool do_sample,
                                             const bf16* ext_in, float* cp_logits, int r0 = 0, int Bg = -1) {
  const q3_model* m = s->m;
  const q3_model_desc& d = m->d;
  const int Btot = s->B, G = s->mega_grid, H = d.hidden, C = d.cp_hidden;
  const int B = Bg < 0 ? Btot : Bg;                               // rows of this group
  unsigned long long* amax_g = s->fs.amax + (size_t)(r0 / MEGA_TMAX) * 15 * MEGA_TMAX;   // the group's arg-max keys [15][B]
  std::vector<M2Phase> pr;
  u64* xT = s->m2_x.as<u64>();
  u64* qkvT = s->m2_qkv.as<u64>();
  u64* attnT = s->m2_attn.as<u64>();
  u64* h1T = s->m2_h1.as<u64>();
  u64* actT = s->m2_act.as<u64>();
  const char* e_samew = std::getenv("Q3_DEBUG_SAMEW");
  const bool samew = e_samew && e_samew[0] == '1';
  auto blank = [](int kind) {
    M2Phase p;
    memset(&p, 0, sizeof(p));
    p.kind = kind; p.xf = XF_NONE; p.rf = XF_NONE; p.yf = XF_NONE;
    return p;
  };
  auto pick_small = [&](M2Phase& p) {
    const int units = p.N / 8, per_cta = (units + G - 1) / G, tiles = (per_cta + 1) / 2;
    const bool dual = p.flags & PF_DUAL, norm = p.flags & PF_NORM;
    p.small = 0;
    if (dual) {
      if (norm && p.xf == XF_F32T && p.K == 1024 && tiles <= 2) p.small = 0x21;
    } else if (norm) {
      if (p.xf == XF_BF16T && p.K == 1024 && tiles <= 2) p.small = 0x21;
      else if (p.xf == XF_BF16T && p.K == 2048 && tiles <= 2 && p.T <= 8) p.small = 0x22;
    } else if (p.xf == XF_GATHER) {
      if (p.K == 2048 && tiles <= 1) p.small = 0x12;
    } else if (p.xf == XF_BF16T) {
      if (p.K == 2048 && tiles <= 1) p.small = 0x12;
      else if (p.K == 3072 && tiles <= 1) p.small = 0x13;
    }
    const char* e = std::getenv("Q3_SMALL");
    if (e && e[0] == '0') p.small = 0;
  };
  auto layers = [&](const std::vector<LayerW>& L, const StackDims& dm, int T, int S, bool cp, int pos_add, bf16* kc, bf16* vc,
                    int cache_seq, int row0) {
    const int nh = dm.heads + 2 * dm.kv_heads;
    const size_t layer_stride = (size_t)Btot * dm.kv_heads * cache_seq * 128;
    const size_t row_off = (size_t)(r0 + row0) * dm.kv_heads * cache_seq * 128;   // first batch row of this group in the caches
    for (int l = 0; l < dm.layers; ++l) {
      const LayerW& w = samew ? L[0] : L[l];    // Q3_DEBUG_SAMEW=1: timing experiment with L2-resident weights (wrong results)
      M2Phase q = blank(M2_GEMV);      // rms_norm(x) -> [q;k;v]
      q.W = w.wqkv; q.N = nh * 128; q.K = dm.H; q.T = T; q.flags = PF_NORM; q.aux = w.in_ln; q.X = xT; q.ldx = dm.H;
      q.xf = XF_BF16T; q.epi = EPI_STORE; q.Y = qkvT; q.ldy = nh * 128; q.yf = XF_BF16T;
      pick_small(q); pr.push_back(q);
      M2Phase at = blank(M2_ATTN);     // QK-norm, RoPE, KV append, attention
      at.X = qkvT; at.Y = attnT; at.W = kc + l * layer_stride + row_off; at.W2 = vc + l * layer_stride + row_off; at.aux = w.q_norm;
      at.aux2 = w.k_norm; at.N = dm.heads; at.K = dm.kv_heads; at.T = T; at.S = S; at.pos_add = pos_add;
      at.ldx = nh * 128; at.ldy = dm.heads * 128; at.flags = cp ? PF_CP : 0;
      pr.push_back(at);
      M2Phase o = blank(M2_GEMV);      // o_proj + residual: h1 = x + attn_out (un-rounded f32 slots)
      o.W = w.wo; o.N = dm.H; o.K = dm.heads * 128; o.T = T; o.X = attnT; o.ldx = dm.heads * 128; o.xf = XF_BF16T;
      o.epi = EPI_O_H1; o.R = xT; o.ldr = dm.H; o.rf = XF_BF16T; o.Y = h1T; o.ldy = dm.H; o.yf = XF_F32T;
      pick_small(o); pr.push_back(o);
      M2Phase gu = blank(M2_GEMV);     // post-attention RMSNorm -> SwiGLU(gate, up)
      gu.W = w.gate; gu.W2 = w.up; gu.N = dm.I; gu.K = dm.H; gu.T = T; gu.flags = PF_NORM | PF_DUAL; gu.aux = w.post_ln;
      gu.X = h1T; gu.ldx = dm.H; gu.xf = XF_F32T; gu.epi = EPI_SWIGLU; gu.Y = actT; gu.ldy = dm.I; gu.yf = XF_BF16T;
      pick_small(gu); pr.push_back(gu);
      M2Phase dn = blank(M2_GEMV);     // down_proj + residual -> x
      dn.W = w.down; dn.N = dm.H; dn.K = dm.I; dn.T = T; dn.X = actT; dn.ldx = dm.I; dn.xf = XF_BF16T;
      dn.epi = EPI_RESIDUAL; dn.R = h1T; dn.ldr = dm.H; dn.rf = XF_F32T; dn.Y = xT; dn.ldy = dm.H; dn.yf = XF_BF16T;
      pick_small(dn); pr.push_back(dn);
    }
  };
  {
    M2Phase p0 = blank(M2_PROLOGUE);
    p0.flags = PF_WAIT_ACQ | PF_ARRIVE_REL;
    pr.push_back(p0);
  }
  const int n_ac = d.groups - 1;
  if (do_cp) {
    for (int g = 0; g < n_ac; ++g) {
      // pass 0 carries two tokens per row (last hidden, semantic embedding): batches above 8 run it in row groups of 8
      // so that no phase exceeds MEGA_TMAX = 16 tokens; the later passes have one token per row.
      const int group = g == 0 ? std::min(B, MEGA_TMAX / 2) : B;
      for (int row0 = 0; row0 < B; row0 += group) {
        const int Bg = std::min(group, B - row0);
        const int T = g == 0 ? 2 * Bg : Bg, S = g == 0 ? 2 : 1;
        const bf16* emb = g == 0 ? m->codec_emb : m->cp_emb[g - 1];
        if (m->cp_proj_w) {
          M2Phase q = blank(M2_GEMV);
          q.W = m->cp_proj_w; q.N = C; q.K = H; q.T = T; q.xf = XF_GATHER; q.aux2 = emb; q.g = g; q.pos_add = row0;
          q.flags = PF_WAIT_ACQ | (g == 0 ? PF_CP0 : 0); q.aux = m->cp_proj_b; q.epi = EPI_BIAS; q.Y = xT; q.ldy = C;
          q.yf = XF_BF16T;
          pick_small(q); pr.push_back(q);
        } else {
          M2Phase q = blank(M2_GATHER);
          q.K = H; q.T = T; q.g = g; q.aux2 = emb; q.Y = xT; q.ldy = C; q.flags = PF_WAIT_ACQ; q.pos_add = row0;
          pr.push_back(q);
        }
        layers(m->cl, m->cdims(), T, S, true, g == 0 ? 0 : g + 1, s->cp_k.as<bf16>(), s->cp_v.as<bf16>(), d.cp_max_seq, row0);
        M2Phase h = blank(M2_GEMV);
        h.W = m->cp_head[samew ? 0 : g]; h.N = d.cp_vocab; h.K = C; h.T = Bg; h.flags = PF_NORM | PF_ARRIVE_REL; h.aux = m->cp_norm;
        h.xf = XF_BF16T;
        h.X = g == 0 ? (const void*)(reinterpret_cast<const char*>(xT) + (size_t)C * 4) : (const void*)xT;
        h.ldx = g == 0 ? 2 * C : C;
        h.epi = EPI_LOGITS; h.amax = amax_g + (size_t)g * B + row0;
        {
          // one-off programs pass their own buffer ([15][B] rows, r0 == 0); the full-frame program of a tapped run writes
          // into the session's tap buffer, whose leading dimension is the whole batch
          float* lg = cp_logits ? cp_logits : s->tap_cp_logits;
          const int ldb = cp_logits ? B : Btot;
          h.Yf = lg ? lg + ((size_t)g * ldb + (cp_logits ? 0 : r0) + row0) * d.cp_vocab : nullptr;
        }
        pick_small(h); pr.push_back(h);
      }
    }
  }
  if (do_finish) {
    M2Phase f = blank(M2_FINISH);
    f.Y = xT; f.flags = PF_WAIT_ACQ;
    pr.push_back(f);
  }
  if (do_talker) {
    if (!do_finish) {
      M2Phase c = blank(M2_COPYIN);
      c.X = ext_in; c.Y = xT; c.flags = PF_WAIT_ACQ;
      pr.push_back(c);
    }
    layers(m->tl, m->tdims(), B, 1, false, 0, s->tk_k.as<bf16>(), s->tk_v.as<bf16>(), s->max_seq, 0);
    M2Phase h = blank(M2_GEMV);
    h.W = m->codec_head; h.N = d.codec_vocab; h.K = H; h.T = B; h.flags = PF_NORM | PF_ARRIVE_REL; h.aux = m->t_norm;
    h.xf = XF_BF16T; h.X = xT; h.ldx = H; h.xn_out = s->fs.last_hidden + (size_t)r0 * H; h.epi = EPI_LOGITS;
    h.Yf = s->logits.as<float>() + (size_t)r0 * d.codec_vocab;
    pick_small(h); pr.push_back(h);
  }
  if (do_sample) {
    M2Phase sp = blank(M2_SAMPLE);
    sp.flags = PF_WAIT_ACQ | PF_ARRIVE_REL;
    pr.push_back(sp);
  }
  // generation 5 (mega5.cuh): the register-resident phases take their weights from the TMA ring
  if (s->mega_ver == 5)
    for (M2Phase& ph : pr)
      if (m5_ring_phase(ph.kind, ph.small, ph.K) && m5_region_bytes(ph.N, ph.K, ph.flags & PF_DUAL, G) <= (size_t)s->m5_slots) ph.flags |= PF_RING;
  // L2 prefetch plan (Q3_PREFETCH: 0 none, 1 the next skinny-GEMM phase's rows at the end of every GEMV phase,
  // 2 (default) additionally: the attention phase requests the gate/up rows two phases ahead at its start and the
  // o_proj phase that follows requests nothing)
  {
    const char* e2 = std::getenv("Q3_PREFETCH");
    const int mode = e2 ? std::atoi(e2) : 2;
    const int n = (int)pr.size();
    for (int i = 0; i < n; ++i) {
      pr[i].next_gemv = -1;
      if (mode == 0) continue;
      if (pr[i].kind == M2_GEMV) {
        int j = (i + 1) % n;
        if (pr[j].kind != M2_GEMV) j = (j + 1) % n;
        if (pr[j].kind == M2_GEMV) pr[i].next_gemv = j;
        if (mode == 2 && i > 0 && pr[i - 1].kind == M2_ATTN) pr[i].next_gemv = -1;
      } else if (pr[i].kind == M2_ATTN && mode == 2) {
        if (i + 2 < n && pr[i + 2].kind == M2_GEMV) pr[i].next_gemv = i + 2;
      }
    }
  }
  return pr;
}

// Split-KV attention (m2_attn_units, mega2.cuh): rows whose context has reached `thr` positions are cut over split_ns CTAs.
// The kernel takes the unit-mapped path only in launches in which some row CAN reach the threshold (contexts grow by one
// position per frame): frames_end = frames generated when the launch ends.  Whether a given row is split depends on its own
// context length only.  Q3_SPLIT_KV = threshold in positions (default 512), 0 = never.
static int split_kv_threshold(const q3_session* s, int frames_end) {
  const int thr = s->split_thr;                  // read from Q3_SPLIT_KV when the session was created
  if (thr <= M2_ATT_FAST_L || !s->m2_xchg.p) return 0;
  int max_len = 0;
  for (int b = 0; b < s->B; ++b) max_len = std::max(max_len, s->prefill_len[b]);
  return max_len + frames_end + 1 >= thr ? thr : 0;
}

static M2Args mega2_args(q3_session* s, int r0 = 0, int Bg = -1) {
  const q3_model* m = s->m;
  const q3_model_desc& d = m->d;
  M2Args a;
  memset(&a, 0, sizeof(a));
  if (Bg < 0) Bg = s->B;
  a.B = Bg; a.H = d.hidden; a.n_ac = d.groups - 1; a.eps = d.rms_eps;
  a.fs = s->fs;
  {
    // per-row state of the group: every array shifted to its first row
    FrameState& f = a.fs;
    const size_t V = d.codec_vocab, H = d.hidden;
    f.cur_tok += r0; f.done += r0; f.n_frames += r0; f.token_count += r0; f.offset += r0; f.frame_idx += r0; f.rng += r0;
    f.seen += (size_t)r0 * V; f.last_hidden += (size_t)r0 * H; f.trailing += (size_t)r0 * f.lt_max * H; f.lt += r0;
    f.codes += (size_t)r0 * f.frames_cap * 16; f.frame_codes += (size_t)r0 * 16;
    f.amax += (size_t)(r0 / MEGA_TMAX) * 15 * MEGA_TMAX;
  }
  a.cp_cos = m->cp_cos; a.cp_sin = m->cp_sin; a.t_cos = s->cos_tab.as<bf16>(); a.t_sin = s->sin_tab.as<bf16>();
  a.max_seq = s->max_seq; a.cp_max_seq = d.cp_max_seq;
  a.codec_emb = m->codec_emb;
  for (int i = 0; i < 15; ++i) a.cp_emb[i] = m->cp_emb[i];
  a.step_input = s->step_input.as<bf16>() + (size_t)r0 * d.hidden;
  SampleArgs sa = make_sample_args(s->cfg, d.codec_vocab, Bg);
  sa.logits = s->logits.as<float>() + (size_t)r0 * d.codec_vocab;
  sa.seen = a.fs.seen; sa.rng = a.fs.rng; sa.tok_out = a.fs.cur_tok; sa.token_count = a.fs.token_count;
  sa.done = a.fs.done; sa.offset = a.fs.offset; sa.frame_idx = a.fs.frame_idx; sa.host_flags = nullptr; sa.advance = 1;
  a.smp = sa;
  a.bar = s->bar.as<unsigned>();
  a.tag_ctr = s->m2_tag.as<unsigned>();
  a.err = s->host_flags_dev + 4;
  const char* e2 = std::getenv("Q3_PREFETCH");
  a.prefetch = e2 ? (std::atoi(e2) != 0) : 1;      // the plan itself is part of the program (m2_build_program)
  {
    // Q3_PF_SPLIT=1: the request is split over the 16 warps (m2_prefetch).  Measured 3 % SLOWER (2.99 vs 2.89 ms per frame,
    // 1.7B batch 8): the L2 prefetch does not hold its issuing thread the way a bulk copy into shared memory does.
    const char* e6 = std::getenv("Q3_PF_SPLIT");
    if (a.prefetch && e6 && std::atoi(e6) != 0) a.prefetch = 2;
  }
  const char* e3 = std::getenv("Q3_PF_SLEEP");
  a.pf_sleep = e3 ? std::atoi(e3) : 200;
  const char* e4 = std::getenv("Q3_RING_SHIFT");
  a.ring_shift = e4 ? std::min(3, std::max(0, std::atoi(e4))) : 3;
  a.m4_slots = s->mega_ver == 5 ? s->m5_slots : s->m4_slots; a.m4_red2 = s->m4_red2;
  a.xchg = s->m2_xchg.as<u64>();
  a.split_min_l = split_kv_threshold(s, s->frames_run + 17);      // callers that know the launch's last frame set it exactly
  a.split_ns = s->split_ns;
  return a;
}

static void mega2_check_watchdog(q3_session* s) {
  if (s->host_flags && s->host_flags[4] != 0) {
    const int code = s->host_flags[4];
    s->host_flags[4] = 0;
    throw Q3Error(Q3_ERR_CUDA, "persistent decode kernel watchdog fired (code " + std::to_string(code) + ")");
  }
}

static void mega2_launch(q3_session* s, M2Args& a, const DBuf& prog, int n_ph, bool force_v2 = false) {
  const unsigned magic = (unsigned)(((1ull << 24) + (unsigned)s->mega_grid - 1) / (unsigned)s->mega_grid);
  Q3_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_grid_magic, &magic, 4, 0, cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaMemsetAsync(s->bar.p, 0, 4, s->st));
  a.prog = prog.as<M2Phase>();
  a.n_ph = n_ph;
  if (s->prof.p) {
    a.prof = s->prof.as<unsigned long long>(); a.prof_cap = (int)(s->prof.bytes / 8);
    const char* pm = std::getenv("Q3_PROF_MODE");
    a.prof_mode = pm ? std::atoi(pm) : 0;
    if (a.prof_mode == 2 && (size_t)a.prof_cap < (size_t)(n_ph * 4 + 8) * s->mega_grid + 2048) a.prof_mode = 0;
  }
  void* params[] = {(void*)&a};
  if (s->mega_ver == 5 && !force_v2)
    Q3_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)decode_frames_mega5_kernel, dim3(s->mega_grid), dim3(MEGA_THREADS), params,
                                              s->m5_smem, s->st));
  else if (s->mega_ver == 4 && !force_v2)
    Q3_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)decode_frames_mega4_kernel, dim3(s->mega_grid), dim3(M4_THREADS), params,
                                              s->m4_smem, s->st));
  else if (s->mega_ver == 3 && !force_v2)
    Q3_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)decode_frames_mega3_kernel, dim3(s->mega_grid), dim3(M3_THREADS), params,
                                              s->m3_smem, s->st));
  else
    Q3_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)decode_frames_mega2_kernel, dim3(s->mega_grid), dim3(MEGA_THREADS), params,
                                              s->m2_smem, s->st));
  Q3_COUNT_LAUNCH();
}

// one-off program (per-op entry points): built, uploaded, launched
static void mega2_run_mode(q3_session* s, bool do_cp, bool do_finish, bool do_talker, bool do_sample, const bf16* ext_in,
                           float* cp_logits) {
  std::vector<M2Phase> pr = m2_build_program(s, do_cp, do_finish, do_talker, do_sample, ext_in, cp_logits);
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));        // the previous one-off program may still be in use
  s->m2_prog_tmp.ensure(pr.size() * sizeof(M2Phase));
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->m2_prog_tmp.p, pr.data(), pr.size() * sizeof(M2Phase), cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));        // pr is a local
  M2Args a = mega2_args(s);
  a.n_frames = 1; a.do_sample = do_sample;
  mega2_launch(s, a, s->m2_prog_tmp, (int)pr.size());
}

static void run_frames_mega2(q3_session* s, int n) {
  // row groups of at most MEGA_TMAX rows, each with its own cached phase program
  const int n_groups = (s->B + MEGA_TMAX - 1) / MEGA_TMAX;
  if (s->m2_n_ph == 0) {
    s->m2_progs.clear();
    s->m2_progs.resize(n_groups);
    s->m2_group_nph.assign(n_groups, 0);
    for (int gi = 0; gi < n_groups; ++gi) {
      const int r0 = gi * MEGA_TMAX, Bg = std::min(MEGA_TMAX, s->B - r0);
      std::vector<M2Phase> pr = m2_build_program(s, true, true, true, true, nullptr, nullptr, r0, Bg);
      s->m2_progs[gi].ensure(pr.size() * sizeof(M2Phase));
      Q3_CHECK_CUDA(cudaMemcpyAsync(s->m2_progs[gi].p, pr.data(), pr.size() * sizeof(M2Phase), cudaMemcpyHostToDevice, s->st));
      Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
      s->m2_group_nph[gi] = (int)pr.size();
    }
    s->m2_n_ph = s->m2_group_nph[0];
  }
  const int per_launch = 16;
  int done_frames = 0, blk = 0;
  bool stop = false;
  while (done_frames < n && !stop) {
    const int todo = std::min(per_launch, n - done_frames);
    for (int gi = 0; gi < n_groups; ++gi) {
      const int r0 = gi * MEGA_TMAX, Bg = std::min(MEGA_TMAX, s->B - r0);
      M2Args a = mega2_args(s, r0, Bg);
      a.n_frames = todo; a.do_sample = 1;
      a.split_min_l = split_kv_threshold(s, s->frames_run + done_frames + todo);
      mega2_launch(s, a, s->m2_progs[gi], s->m2_group_nph[gi]);
    }
    done_frames += todo;
    count_active_kernel<<<1, 32, 0, s->st>>>(s->fs.done, s->B, s->host_flags_dev + (blk & 1));
    Q3_COUNT_LAUNCH();
    Q3_CHECK_CUDA(cudaEventRecord(s->ev_poll[blk & 1], s->st));
    if (blk > 0) {
      Q3_CHECK_CUDA(cudaEventSynchronize(s->ev_poll[(blk - 1) & 1]));
      mega2_check_watchdog(s);
      if (s->host_flags[(blk - 1) & 1] == 0) stop = true;
    }
    ++blk;
  }
  s->frames_run += done_frames;
}

static void run_frames_mega(q3_session* s, int n) {
  const int per_launch = 16;
  int done_frames = 0, blk = 0;
  bool stop = false;
  while (done_frames < n && !stop) {
    const int todo = std::min(per_launch, n - done_frames);
    MegaArgs a = mega_args(s);
    a.n_frames = todo; a.do_cp = a.do_finish = a.do_talker = a.do_sample = 1;
    mega_launch(s, a);
    done_frames += todo;
    count_active_kernel<<<1, 32, 0, s->st>>>(s->fs.done, s->B, s->host_flags_dev + (blk & 1));
    Q3_COUNT_LAUNCH();
    Q3_CHECK_CUDA(cudaEventRecord(s->ev_poll[blk & 1], s->st));
    if (blk > 0) {
      Q3_CHECK_CUDA(cudaEventSynchronize(s->ev_poll[(blk - 1) & 1]));
      if (s->host_flags[(blk - 1) & 1] == 0) stop = true;
    }
    ++blk;
  }
  s->frames_run += done_frames;
}

static void run_frames(q3_session* s, int n) {
  if (n <= 0) return;
  NvtxRange nvtx("generate_frames");
  // KV overflow check (kv_cache.rs:293-300)
  int max_len = 0;
  for (int b = 0; b < s->B; ++b) max_len = std::max(max_len, s->prefill_len[b]);
  if (max_len + s->frames_run + n > s->max_seq)
    throw Q3Error(Q3_ERR_KV_OVERFLOW, "KV cache overflow: current=" + std::to_string(max_len + s->frames_run) +
                                          " + new=" + std::to_string(n) + " > max=" + std::to_string(s->max_seq));
  if (s->use_mega) {
    if (s->mega_ver >= 2) run_frames_mega2(s, n);
    else run_frames_mega(s, n);
    return;
  }
  int done_frames = 0;
  uint64_t per_frame = 0;
  if (!s->graph_ok) {
    // first frame eagerly (also performs every lazy cudaFuncSetAttribute), then capture one frame
    uint64_t before = g_q3_launches.load();
    frame_body(s);
    per_frame = g_q3_launches.load() - before;
    done_frames = 1;
    if (n > 1) {
      cudaGraph_t graph = nullptr;
      Q3_CHECK_CUDA(cudaStreamBeginCapture(s->st, cudaStreamCaptureModeThreadLocal));
      try {
        frame_body(s);
      } catch (...) {
        cudaStreamEndCapture(s->st, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      Q3_CHECK_CUDA(cudaStreamEndCapture(s->st, &graph));
      g_q3_launches.fetch_sub(per_frame);      // the captured launches did not execute
      Q3_CHECK_CUDA(cudaGraphInstantiate(&s->graph_exec, graph, 0));
      cudaGraphDestroy(graph);
      s->graph_ok = true;
      s->graph_launches = per_frame;
    }
  }
  const uint64_t graph_launches = s->graph_launches;
  const int poll_every = 16;
  int blk = 0;
  bool stop = false;
  while (done_frames < n && !stop) {
    int todo = std::min(poll_every, n - done_frames);
    for (int i = 0; i < todo; ++i) {
      Q3_CHECK_CUDA(cudaGraphLaunch(s->graph_exec, s->st));
      g_q3_launches.fetch_add(graph_launches);
    }
    done_frames += todo;
    count_active_kernel<<<1, 32, 0, s->st>>>(s->fs.done, s->B, s->host_flags_dev + (blk & 1));
    Q3_COUNT_LAUNCH();
    Q3_CHECK_CUDA(cudaEventRecord(s->ev_poll[blk & 1], s->st));
    if (blk > 0) {
      // look at the flag written one block ago: no pipeline drain
      Q3_CHECK_CUDA(cudaEventSynchronize(s->ev_poll[(blk - 1) & 1]));
      if (s->host_flags[(blk - 1) & 1] == 0) stop = true;
    }
    ++blk;
  }
  s->frames_run += done_frames;
}

static void sample_first_if_needed(q3_session* s) {
  if (s->first_sampled) return;
  launch_sampler(s, 0);                       // lib.rs:557-571, token_count = 0
  s->first_sampled = true;
}

// =================================================================================================
// ABI
#define Q3_API_BEGIN try {
#define Q3_API_END                                                       \
  }                                                                      \
  catch (const Q3Error& e) { g_last_error = e.what(); return e.code; }   \
  catch (const std::exception& e) { g_last_error = e.what(); return Q3_ERR_INVALID; } \
  catch (...) { g_last_error = "unknown error"; return Q3_ERR_INVALID; } \
  return Q3_OK;

extern "C" {

const char* q3_last_error(void) { return g_last_error.c_str(); }
int q3_abi_version(void) { return Q3_ABI_VERSION; }
uint64_t q3_kernel_launch_count(void) { return g_q3_launches.load(); }

q3_status q3_model_create(const q3_model_desc* desc, q3_model** out) {
  Q3_API_BEGIN
  Q3_REQUIRE(desc && out, Q3_ERR_INVALID, "null argument");
  int ndev = 0;
  Q3_CHECK_CUDA(cudaGetDeviceCount(&ndev));
  Q3_REQUIRE(ndev > 0 && desc->device < ndev, Q3_ERR_CUDA, "no usable CUDA device (there is no CPU fallback)");
  cudaDeviceProp prop;
  Q3_CHECK_CUDA(cudaGetDeviceProperties(&prop, desc->device));
  Q3_REQUIRE(prop.major == 10, Q3_ERR_CUDA, "libq3tts_b200 is built for sm_100a only; found sm_" +
                                                std::to_string(prop.major) + std::to_string(prop.minor));
  Q3_REQUIRE(desc->head_dim == 128, Q3_ERR_UNSUPPORTED, "head_dim must be 128");
  Q3_REQUIRE(desc->heads == 2 * desc->kv_heads && desc->cp_heads == 2 * desc->cp_kv_heads, Q3_ERR_UNSUPPORTED,
             "GQA group size must be 2");
  Q3_REQUIRE(desc->codec_vocab <= 4096 && desc->codec_vocab > 1024, Q3_ERR_UNSUPPORTED, "codec vocab must be in (1024, 4096]");
  Q3_REQUIRE(desc->groups == 16, Q3_ERR_UNSUPPORTED, "16 code groups expected");
  Q3_REQUIRE(desc->hidden % 8 == 0 && desc->cp_hidden % 8 == 0 && desc->inter % 8 == 0 && desc->cp_inter % 8 == 0 &&
                 desc->text_embed_dim % 8 == 0,
             Q3_ERR_UNSUPPORTED, "dimensions must be multiples of 8");
  Q3_CHECK_CUDA(cudaSetDevice(desc->device));
  auto* m = new q3_model();
  m->d = *desc;
  m->num_sms = prop.multiProcessorCount;
  *out = m;
  Q3_API_END
}

q3_status q3_model_set_tensor(q3_model* m, const char* hf_name, const void* data, q3_dtype dtype, const int64_t* shape,
                              int32_t ndim, int32_t on_device) {
  Q3_API_BEGIN
  Q3_REQUIRE(m && hf_name && data && shape && ndim > 0 && ndim <= 4, Q3_ERR_INVALID, "bad argument");
  Q3_REQUIRE(!m->finalized, Q3_ERR_STATE, "model already finalized");
  Q3_CHECK_CUDA(cudaSetDevice(m->d.device));
  std::string name(hf_name);
  RawTensor t;
  {
    size_t cnt = 1;        // refuse shapes whose element count would wrap around (headers come from files)
    for (int i = 0; i < ndim; ++i) {
      Q3_REQUIRE(shape[i] >= 0 && (shape[i] == 0 || cnt <= ((size_t)1 << 40) / (size_t)shape[i]), Q3_ERR_INVALID,
                 "tensor shape out of range");
      cnt *= (size_t)shape[i];
    }
  }
  t.shape.assign(shape, shape + ndim);
  const size_t n = t.numel();
  const bool is_voc = name.rfind("decoder.", 0) == 0 || name.rfind("speaker_encoder.", 0) == 0;   // the F32 parts of the model
  const q3_dtype want = is_voc ? Q3_F32 : Q3_BF16;
  const size_t src_bytes = n * (dtype == Q3_BF16 ? 2 : 4);
  DBuf src;
  const void* dsrc = data;
  if (!on_device) {
    src.alloc(src_bytes);
    Q3_CHECK_CUDA(cudaMemcpy(src.p, data, src_bytes, cudaMemcpyHostToDevice));
    dsrc = src.p;
  }
  t.dtype = want;
  t.buf.alloc(n * (want == Q3_BF16 ? 2 : 4));
  if (dtype == want) {
    Q3_CHECK_CUDA(cudaMemcpy(t.buf.p, dsrc, src_bytes, cudaMemcpyDeviceToDevice));
  } else if (want == Q3_BF16) {
    f32_to_bf16_kernel<<<1024, 256>>>((const float*)dsrc, t.buf.as<bf16>(), n);
    Q3_COUNT_LAUNCH();
    Q3_LAUNCH_CHECK();
    Q3_CHECK_CUDA(cudaDeviceSynchronize());
  } else {
    bf16_to_f32_kernel<<<1024, 256>>>((const bf16*)dsrc, t.buf.as<float>(), n);
    Q3_COUNT_LAUNCH();
    Q3_LAUNCH_CHECK();
    Q3_CHECK_CUDA(cudaDeviceSynchronize());
  }
  m->t[name] = std::move(t);
  Q3_API_END
}

static void check_shape(const RawTensor& t, std::initializer_list<int64_t> want, const std::string& name) {
  std::vector<int64_t> w(want);
  if (t.shape != w) {
    std::string s = "shape mismatch for " + name + ": got [";
    for (auto d : t.shape) s += std::to_string(d) + ",";
    s += "] want [";
    for (auto d : w) s += std::to_string(d) + ",";
    throw Q3Error(Q3_ERR_INVALID, s + "]");
  }
}

static void build_stack(q3_model* m, const std::string& prefix, const StackDims& dm, std::vector<LayerW>& out) {
  out.clear();
  const int qd = dm.heads * 128, kd = dm.kv_heads * 128;
  for (int l = 0; l < dm.layers; ++l) {
    const std::string p = prefix + ".layers." + std::to_string(l);
    LayerW w{};
    w.in_ln = needb(m, p + ".input_layernorm.weight");
    w.post_ln = needb(m, p + ".post_attention_layernorm.weight");
    w.q_norm = needb(m, p + ".self_attn.q_norm.weight");
    w.k_norm = needb(m, p + ".self_attn.k_norm.weight");
    const RawTensor& q = need_bf16(m, p + ".self_attn.q_proj.weight");
    const RawTensor& k = need_bf16(m, p + ".self_attn.k_proj.weight");
    const RawTensor& v = need_bf16(m, p + ".self_attn.v_proj.weight");
    check_shape(q, {qd, dm.H}, p + ".q_proj");
    check_shape(k, {kd, dm.H}, p + ".k_proj");
    check_shape(v, {kd, dm.H}, p + ".v_proj");
    // fused [q;k;v] weight so one pass over the activations yields all three projections
    DBuf fused;
    fused.alloc((size_t)(qd + 2 * kd) * dm.H * 2);
    Q3_CHECK_CUDA(cudaMemcpy(fused.p, q.buf.p, (size_t)qd * dm.H * 2, cudaMemcpyDeviceToDevice));
    Q3_CHECK_CUDA(cudaMemcpy((char*)fused.p + (size_t)qd * dm.H * 2, k.buf.p, (size_t)kd * dm.H * 2, cudaMemcpyDeviceToDevice));
    Q3_CHECK_CUDA(cudaMemcpy((char*)fused.p + (size_t)(qd + kd) * dm.H * 2, v.buf.p, (size_t)kd * dm.H * 2, cudaMemcpyDeviceToDevice));
    w.wqkv = fused.as<bf16>();
    m->owned.push_back(std::move(fused));
    m->t.erase(p + ".self_attn.q_proj.weight");
    m->t.erase(p + ".self_attn.k_proj.weight");
    m->t.erase(p + ".self_attn.v_proj.weight");
    const RawTensor& o = need_bf16(m, p + ".self_attn.o_proj.weight");
    check_shape(o, {dm.H, qd}, p + ".o_proj");
    w.wo = o.buf.as<bf16>();
    const RawTensor& g = need_bf16(m, p + ".mlp.gate_proj.weight");
    const RawTensor& u = need_bf16(m, p + ".mlp.up_proj.weight");
    const RawTensor& dn = need_bf16(m, p + ".mlp.down_proj.weight");
    check_shape(g, {dm.I, dm.H}, p + ".gate_proj");
    check_shape(u, {dm.I, dm.H}, p + ".up_proj");
    check_shape(dn, {dm.H, dm.I}, p + ".down_proj");
    w.gate = g.buf.as<bf16>(); w.up = u.buf.as<bf16>(); w.down = dn.buf.as<bf16>();
    out.push_back(w);
  }
}

// bf16 RoPE tables: angle = (float)pos * inv_freq (f32), cos/sin in f32, then rounded to the activation
// dtype before use (transformer.rs:48-57, 79-90, 133-175).
static void build_rope_table(int n_pos, float theta, DBuf& cos_out, DBuf& sin_out) {
  std::vector<uint16_t> c((size_t)n_pos * 64), s((size_t)n_pos * 64);
  float inv[64];
  for (int i = 0; i < 64; ++i) inv[i] = 1.0f / powf(theta, (float)(2 * i) / 128.0f);
  auto to_bf16 = [](float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    uint32_t r = u + 0x7fffu + ((u >> 16) & 1u);       // round to nearest even (finite inputs)
    return (uint16_t)(r >> 16);
  };
  for (int p = 0; p < n_pos; ++p)
    for (int i = 0; i < 64; ++i) {
      float ang = (float)p * inv[i];
      c[(size_t)p * 64 + i] = to_bf16(cosf(ang));
      s[(size_t)p * 64 + i] = to_bf16(sinf(ang));
    }
  cos_out.alloc(c.size() * 2);
  sin_out.alloc(s.size() * 2);
  Q3_CHECK_CUDA(cudaMemcpy(cos_out.p, c.data(), c.size() * 2, cudaMemcpyHostToDevice));
  Q3_CHECK_CUDA(cudaMemcpy(sin_out.p, s.data(), s.size() * 2, cudaMemcpyHostToDevice));
}

q3_status q3_model_finalize(q3_model* m) {
  Q3_API_BEGIN
  Q3_REQUIRE(m, Q3_ERR_INVALID, "null model");
  Q3_REQUIRE(!m->finalized, Q3_ERR_STATE, "model already finalized");
  Q3_CHECK_CUDA(cudaSetDevice(m->d.device));
  const q3_model_desc& d = m->d;
  const bool any_talker = m->t.count("talker.model.norm.weight") > 0;
  const bool any_voc = m->t.count("decoder.pre_conv.conv.weight") > 0;
  const bool any_spk = m->t.count("speaker_encoder.blocks.0.conv.weight") > 0;
  Q3_REQUIRE(any_talker || any_voc || any_spk, Q3_ERR_MISSING_WEIGHT,
             "Missing weight: no talker.*, decoder.* or speaker_encoder.* tensors were set");
  if (any_talker) {
    m->codec_emb = needb(m, "talker.model.codec_embedding.weight");
    check_shape(need_bf16(m, "talker.model.codec_embedding.weight"), {d.codec_vocab, d.hidden}, "codec_embedding");
    if (m->t.count("talker.model.text_embedding.weight")) {
      m->text_emb = needb(m, "talker.model.text_embedding.weight");
      m->fc1_w = needb(m, "talker.text_projection.linear_fc1.weight");
      m->fc1_b = needb(m, "talker.text_projection.linear_fc1.bias");
      m->fc2_w = needb(m, "talker.text_projection.linear_fc2.weight");
      m->fc2_b = needb(m, "talker.text_projection.linear_fc2.bias");
    }
    build_stack(m, "talker.model", m->tdims(), m->tl);
    m->t_norm = needb(m, "talker.model.norm.weight");
    m->codec_head = needb(m, "talker.codec_head.weight");
    check_shape(need_bf16(m, "talker.codec_head.weight"), {d.codec_vocab, d.hidden}, "codec_head");
    const std::string cp = "talker.code_predictor";
    if (d.hidden != d.cp_hidden) {
      m->cp_proj_w = needb(m, cp + ".small_to_mtp_projection.weight");
      m->cp_proj_b = needb(m, cp + ".small_to_mtp_projection.bias");
      check_shape(need_bf16(m, cp + ".small_to_mtp_projection.weight"), {d.cp_hidden, d.hidden}, "small_to_mtp_projection");
    }
    for (int g = 0; g < d.groups - 1; ++g) {
      m->cp_emb[g] = needb(m, cp + ".model.codec_embedding." + std::to_string(g) + ".weight");
      m->cp_head[g] = needb(m, cp + ".lm_head." + std::to_string(g) + ".weight");
      check_shape(need_bf16(m, cp + ".model.codec_embedding." + std::to_string(g) + ".weight"), {d.cp_vocab, d.hidden}, "cp codec_embedding");
      check_shape(need_bf16(m, cp + ".lm_head." + std::to_string(g) + ".weight"), {d.cp_vocab, d.cp_hidden}, "lm_head");
    }
    build_stack(m, cp + ".model", m->cdims(), m->cl);
    m->cp_norm = needb(m, cp + ".model.norm.weight");
    {
      DBuf t1, t2;
      t1.alloc(m->tl.size() * sizeof(LayerW));
      t2.alloc(m->cl.size() * sizeof(LayerW));
      Q3_CHECK_CUDA(cudaMemcpy(t1.p, m->tl.data(), m->tl.size() * sizeof(LayerW), cudaMemcpyHostToDevice));
      Q3_CHECK_CUDA(cudaMemcpy(t2.p, m->cl.data(), m->cl.size() * sizeof(LayerW), cudaMemcpyHostToDevice));
      m->tl_dev = t1.as<LayerW>();
      m->cl_dev = t2.as<LayerW>();
      m->owned.push_back(std::move(t1));
      m->owned.push_back(std::move(t2));
    }
    DBuf c, s;
    build_rope_table(d.cp_rope_positions, d.rope_theta, c, s);
    m->cp_cos = c.as<bf16>();
    m->cp_sin = s.as<bf16>();
    m->owned.push_back(std::move(c));
    m->owned.push_back(std::move(s));
    m->has_talker = true;
  }
  if (any_voc) vocoder_finalize(m);
  if (any_spk) speaker_finalize(m);
  Q3_CHECK_CUDA(cudaDeviceSynchronize());
  m->finalized = true;
  Q3_API_END
}

void q3_model_destroy(q3_model* m) { delete m; }

// -------------------------------------------------------------------------------------------------
static void reset_state(q3_session* s, const uint64_t* seeds) {
  const int B = s->B;
  std::vector<unsigned long long> st(B);
  for (int b = 0; b < B; ++b)   // SamplingContext::new (sampling.rs:36-38)
    st[b] = (unsigned long long)seeds[b] * 2685821657736338717ull + 1442695040888963407ull;
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->rng.p, st.data(), B * 8, cudaMemcpyHostToDevice, s->st));
  s->cur_tok.zero(s->st); s->done.zero(s->st); s->n_frames.zero(s->st); s->token_count.zero(s->st);
  s->offset.zero(s->st); s->frame_idx.zero(s->st); s->seen.zero(s->st); s->amax.zero(s->st);
  s->frame_codes.zero(s->st);
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  s->prefilled = false;
  s->first_sampled = false;
  s->frames_run = 0;
  if (s->m2_tag.p) {
    // Tags are a 32-bit counter that only ever grows inside a session (one per phase, ~550 per frame): a reused session
    // would wrap it after ~7.8 M frames.  A reset is a quiescent point (the stream was just synchronised), so restart the
    // counter here and clear the tagged buffers, whose stale tags could otherwise alias the restarted sequence.
    const unsigned one = 1;
    Q3_CHECK_CUDA(cudaMemcpyAsync(s->m2_tag.p, &one, 4, cudaMemcpyHostToDevice, s->st));
    s->m2_x.zero(s->st); s->m2_qkv.zero(s->st); s->m2_attn.zero(s->st); s->m2_h1.zero(s->st); s->m2_act.zero(s->st);
    if (s->m2_xchg.p) s->m2_xchg.zero(s->st);      // the exchange slots of m2_attn_units carry the same tags
    Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  }
  std::fill(s->prefill_len.begin(), s->prefill_len.end(), 0);
  std::fill(s->stream_emitted.begin(), s->stream_emitted.end(), 0);
  if (s->vstream) s->vstream->frames = 0;
}

q3_status q3_session_create(const q3_model* m, int32_t batch, int32_t max_seq, const q3_gen_config* cfg,
                            const uint64_t* seeds, q3_session** out) {
  Q3_API_BEGIN
  Q3_REQUIRE(m && cfg && seeds && out, Q3_ERR_INVALID, "null argument");
  Q3_REQUIRE(m->finalized && m->has_talker, Q3_ERR_STATE, "model has no finalized talker weights");
  Q3_REQUIRE(batch >= 1 && batch <= 256, Q3_ERR_INVALID, "batch must be in [1, 256]");
  Q3_REQUIRE(max_seq >= 8 && max_seq <= 12288, Q3_ERR_INVALID, "max_seq must be in [8, 12288]");
  Q3_REQUIRE(cfg->max_new_tokens >= 1, Q3_ERR_INVALID, "max_new_tokens must be >= 1");
  Q3_REQUIRE(cfg->eos_token_id < m->d.codec_vocab, Q3_ERR_INVALID, "eos_token_id out of range");
  Q3_CHECK_CUDA(cudaSetDevice(m->d.device));
  const q3_model_desc& d = m->d;
  std::unique_ptr<q3_session> s(new q3_session());
  s->m = m; s->B = batch; s->max_seq = max_seq; s->cfg = *cfg;
  s->frames_cap = std::min(cfg->max_new_tokens, max_seq);
  Q3_CHECK_CUDA(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
  Q3_CHECK_CUDA(cudaEventCreate(&s->ev0));
  Q3_CHECK_CUDA(cudaEventCreate(&s->ev1));
  for (auto& e : s->ev_poll) Q3_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  const int B = batch, V = d.codec_vocab, H = d.hidden;
  s->tk_k.alloc((size_t)d.layers * B * d.kv_heads * max_seq * 128 * 2);
  s->tk_v.alloc((size_t)d.layers * B * d.kv_heads * max_seq * 128 * 2);
  s->cp_k.alloc((size_t)d.cp_layers * B * d.cp_kv_heads * d.cp_max_seq * 128 * 2);
  s->cp_v.alloc((size_t)d.cp_layers * B * d.cp_kv_heads * d.cp_max_seq * 128 * 2);
  s->tk_k.zero(); s->tk_v.zero(); s->cp_k.zero(); s->cp_v.zero();       // PreAllocKVCache::new zero-fills
  build_rope_table(max_seq, d.rope_theta, s->cos_tab, s->sin_tab);
  s->cur_tok.alloc(B * 4); s->done.alloc(B * 4); s->n_frames.alloc(B * 4); s->token_count.alloc(B * 4);
  s->offset.alloc(B * 4); s->frame_idx.alloc(B * 4); s->rng.alloc(B * 8); s->seen.alloc((size_t)B * V);
  s->last_hidden.alloc((size_t)B * H * 2); s->tts_pad.alloc((size_t)H * 2); s->lt.alloc(B * 4);
  s->trailing.alloc((size_t)B * H * 2);
  s->codes.alloc((size_t)B * s->frames_cap * 16 * 4);
  s->amax.alloc((size_t)15 * std::max(B, MEGA_TMAX * ((B + MEGA_TMAX - 1) / MEGA_TMAX)) * 8); s->frame_codes.alloc((size_t)B * 16 * 4);
  s->logits.alloc((size_t)B * V * 4); s->step_input.alloc((size_t)B * H * 2);
  s->cp_x0.alloc((size_t)2 * B * H * 2); s->cp_xe.alloc((size_t)B * H * 2);
  s->lens_dev.alloc(B * 4);
  s->tts_pad.zero(); s->lt.zero(); s->trailing.zero(); s->codes.zero(); s->last_hidden.zero();
  Q3_CHECK_CUDA(cudaHostAlloc((void**)&s->host_flags, 64, cudaHostAllocMapped));
  s->host_flags[0] = s->host_flags[1] = B;
  Q3_CHECK_CUDA(cudaHostGetDevicePointer((void**)&s->host_flags_dev, s->host_flags, 0));
  s->prefill_len.assign(B, 0);
  s->stream_emitted.assign(B, 0);
  ensure_scratch(s.get(), 2 * B);
  if (attn_smem_bytes(max_seq) > 48 * 1024)
    Q3_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)attn_smem_bytes(max_seq)));
  {
    // persistent frame kernel: batch <= 8 (16 tokens in the CP prefill pass), needs one resident CTA per SM
    const char* env = std::getenv("Q3_MEGA");
    const bool want = !(env && env[0] == '0');
#ifndef Q3_ALL_GENERATIONS
    Q3_REQUIRE(!(env && (env[0] == '1' || env[0] == '3')), Q3_ERR_UNSUPPORTED,
               "Q3_MEGA=1 / 3 (the historical generations of the persistent kernel) are only built into libq3tts_b200_dev.so");
#endif
    s->mega_smem = mega_smem_bytes(d, B, max_seq, m->num_sms);
    s->bar.alloc(64);
    s->bar.zero();
    if (want && 2 * B <= MEGA_TMAX && d.layers + d.cp_layers <= MEGA_MAX_LAYERS && s->mega_smem > 0 && s->mega_smem <= 227 * 1024 && d.hidden % 32 == 0 && d.cp_hidden % 32 == 0 &&
        d.inter % 32 == 0 && d.cp_inter % 32 == 0 && d.codec_vocab % 16 == 0 && d.cp_vocab % 16 == 0) {
      Q3_CHECK_CUDA(cudaFuncSetAttribute(decode_frames_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->mega_smem));
      int per_sm = 0;
      Q3_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_frames_mega_kernel, MEGA_THREADS, s->mega_smem));
      if (per_sm >= 1) {
        s->use_mega = true;
        s->mega_grid = m->num_sms;
        s->ss.alloc((size_t)2 * s->mega_grid * MEGA_TMAX * 4);
        s->ss.zero();
      }
    }
    s->mega_ver = (env && env[0] == '1') ? 1 : 2;
    const bool v1_ok = s->use_mega;
    (void)v1_ok;
#ifndef Q3_ALL_GENERATIONS
    s->use_mega = false;             // the first generation is a stub here: only the dataflow / ring kernels below may enable it
#endif
    const bool dims_ok = d.layers + d.cp_layers <= MEGA_MAX_LAYERS && d.hidden % 32 == 0 && d.cp_hidden % 32 == 0 && d.inter % 32 == 0 &&
                         d.cp_inter % 32 == 0 && d.codec_vocab % 16 == 0 && d.cp_vocab % 16 == 0;
    // the dataflow generation also takes batches 9..16 (code-predictor pass 0 in two row groups)
    if (want && s->mega_ver == 2 && dims_ok) {           // batches above 16 run as row groups of 16
      s->use_mega = true;
      s->mega_grid = m->num_sms;
    }
    if (s->use_mega && s->mega_ver == 2) {
      // dataflow generation: tagged activation buffers (8-byte slots), the phase program, the session's tag counter
      const int n_ph_max = 3 + (d.groups - 1 + (B > 8 ? 1 : 0)) * (2 + 5 * d.cp_layers) + 1 + 5 * d.layers + 1;
      s->m2_smem = mega2_smem_bytes(d, std::min(B, (int)MEGA_TMAX), max_seq, m->num_sms, n_ph_max);
      int per_sm = 0;
      if (s->m2_smem > 0 && s->m2_smem <= 227 * 1024) {
        Q3_CHECK_CUDA(cudaFuncSetAttribute(decode_frames_mega2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->m2_smem));
        Q3_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_frames_mega2_kernel, MEGA_THREADS, s->m2_smem));
      }
      if (per_sm >= 1) {
        const size_t Hm = std::max(d.hidden, d.cp_hidden), Im = std::max(d.inter, d.cp_inter);
        const size_t nhm = (size_t)std::max(d.heads + 2 * d.kv_heads, d.cp_heads + 2 * d.cp_kv_heads) * 128;
        const size_t qdm = (size_t)std::max(d.heads, d.cp_heads) * 128;
        s->m2_x.alloc(MEGA_TMAX * Hm * 4); s->m2_qkv.alloc(MEGA_TMAX * nhm * 4); s->m2_attn.alloc(MEGA_TMAX * qdm * 4);
        s->m2_h1.alloc(MEGA_TMAX * Hm * 8); s->m2_act.alloc(MEGA_TMAX * Im * 4); s->m2_tag.alloc(64);
        s->m2_xchg.alloc((size_t)MEGA_TMAX * d.kv_heads * M2_SPLIT_NS_MAX * M2_XCHG_SLOTS * 8);
        s->m2_xchg.zero();
        {
          const char* e8 = std::getenv("Q3_SPLIT_KV");
          s->split_thr = e8 ? std::atoi(e8) : 512;
          // 4 splits by default; 8 (Q3_SPLIT_NS=8) suit a single long-form stream (+10 % at batch 1, -9 % at batch 8, 1536
          // frames).  The count is part of the numerics (summation order), so it is a session setting, never derived from
          // the batch: a row equals its batch-1 run under the same setting.
          const char* e9 = std::getenv("Q3_SPLIT_NS");
          s->split_ns = e9 ? std::max(2, std::min((int)M2_SPLIT_NS_MAX, std::atoi(e9))) : 4;
        }
        s->m2_x.zero(); s->m2_qkv.zero(); s->m2_attn.zero(); s->m2_h1.zero(); s->m2_act.zero();
        const unsigned one = 1;
        Q3_CHECK_CUDA(cudaMemcpy(s->m2_tag.p, &one, 4, cudaMemcpyHostToDevice));
        s->host_flags[4] = 0;
        // TMA weight ring (mega3.cuh): every skinny-GEMM phase must be one of its (K, format) combinations
        const bool want3 = env && env[0] == '3';     // experimental: slower than the dataflow kernel on B200
        if (want3 && B <= 8) {
          s->mega_grid = m->num_sms;
          std::vector<M2Phase> pr = m2_build_program(s.get(), true, true, true, true, nullptr, nullptr);
          bool ok3 = true;
          for (const M2Phase& ph : pr)
            if (ph.kind == M2_GEMV && !m3_gemv_supported(ph.K, ph.flags & PF_DUAL, ph.flags & PF_NORM, ph.xf, ph.T)) ok3 = false;
          s->m3_smem = ok3 ? mega3_smem_bytes(d, B, max_seq, m->num_sms) : 0;
          int per_sm3 = 0;
          if (s->m3_smem > 0 && s->m3_smem + 4096 <= 227 * 1024) {
            Q3_CHECK_CUDA(cudaFuncSetAttribute(decode_frames_mega3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->m3_smem));
            Q3_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm3, decode_frames_mega3_kernel, M3_THREADS, s->m3_smem));
          }
          if (per_sm3 >= 1) s->mega_ver = 3;
        }
        // TMA weight ring, generation 4 (mega4.cuh): opt-in (Q3_MEGA=4).  Measured on B200 (1.7B, batch 8) it is 1.47x
        // SLOWER than the dataflow kernel (4.30 vs 2.93 ms per frame, profiles/r2_mega4_vs_mega2.md), so the dataflow
        // kernel stays the default; models with a skinny-GEMM K that is not a multiple of 1024 cannot use it at all.
        const bool want4 = env && env[0] == '4';
        if (want4) {
          std::vector<M2Phase> pr = m2_build_program(s.get(), true, true, true, true, nullptr, nullptr, 0, std::min(B, (int)MEGA_TMAX));
          bool ok4 = true;
          for (const M2Phase& ph : pr)
            if (ph.kind == M2_GEMV && !m4_gemv_supported(ph.N, ph.K, ph.flags & PF_DUAL, ph.flags & PF_NORM, ph.xf)) ok4 = false;
          s->m4_smem = ok4 ? mega4_smem_bytes(d, std::min(B, (int)MEGA_TMAX), max_seq, &s->m4_slots, &s->m4_red2) : 0;
          {
            const char* e5 = std::getenv("Q3_M4_SLOTS");
            if (e5 && s->m4_smem > 0) {
              const int want_slots = std::max(2, std::min(s->m4_slots, std::atoi(e5)));
              s->m4_smem -= (size_t)(s->m4_slots - want_slots) * M4_SLOT_BYTES;
              s->m4_slots = want_slots;
            }
          }
          int per_sm4 = 0;
          if (s->m4_smem > 0) {
            Q3_CHECK_CUDA(cudaFuncSetAttribute(decode_frames_mega4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->m4_smem));
            Q3_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm4, decode_frames_mega4_kernel, M4_THREADS, s->m4_smem));
          }
          if (per_sm4 >= 1) s->mega_ver = 4;
        }
        // generation 5 (mega5.cuh): mega2's phases with the register-resident ones fed by a TMA ring; Q3_MEGA=5
        const bool want5 = env && env[0] == '5';
        if (want5) {
          size_t work5 = 0;
          const size_t cap5 = mega5_buffer_cap(d, std::min(B, (int)MEGA_TMAX), max_seq, m->num_sms, &work5);
          s->m5_slots = (int)cap5;                 // bytes of the prefetch buffer
          s->m5_smem = cap5 ? cap5 + work5 : 0;
          int per_sm5 = 0;
          if (s->m5_smem > 0) {
            Q3_CHECK_CUDA(cudaFuncSetAttribute(decode_frames_mega5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->m5_smem));
            Q3_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm5, decode_frames_mega5_kernel, MEGA_THREADS, s->m5_smem));
          }
          if (per_sm5 >= 1) s->mega_ver = 5;
        }
#ifdef Q3_ALL_GENERATIONS
      } else if (v1_ok) {
        s->mega_ver = 1;
#endif
      } else {
        s->use_mega = false;
      }
    }
  }
  FrameState& fs = s->fs;
  fs.cur_tok = s->cur_tok.as<uint32_t>(); fs.done = s->done.as<int>(); fs.n_frames = s->n_frames.as<int>();
  fs.token_count = s->token_count.as<int>(); fs.offset = s->offset.as<int>(); fs.frame_idx = s->frame_idx.as<int>();
  fs.rng = s->rng.as<unsigned long long>(); fs.seen = s->seen.as<uint8_t>(); fs.last_hidden = s->last_hidden.as<bf16>();
  fs.trailing = s->trailing.as<bf16>(); fs.lt = s->lt.as<int>(); fs.lt_max = 1; fs.tts_pad = s->tts_pad.as<bf16>();
  fs.codes = s->codes.as<uint32_t>(); fs.frames_cap = s->frames_cap; fs.amax = s->amax.as<unsigned long long>();
  fs.frame_codes = s->frame_codes.as<uint32_t>(); fs.host_flags = s->host_flags_dev;
  // the zero-fills above went to the legacy default stream, the session works on its own non-blocking stream, and
  // recycled buffers (DevPool) may hold a previous session's data: order them explicitly
  Q3_CHECK_CUDA(cudaStreamSynchronize(0));
  reset_state(s.get(), seeds);
  *out = s.release();
  Q3_API_END
}

q3_status q3_session_reset(q3_session* s, const uint64_t* seeds) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && seeds, Q3_ERR_INVALID, "null argument");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  reset_state(s, seeds);
  Q3_API_END
}

void q3_session_destroy(q3_session* s) {
  if (!s) return;
  cudaSetDevice(s->m->d.device);
  if (s->st) cudaStreamSynchronize(s->st);
  delete s;
}

void* q3_session_stream(q3_session* s) { return s ? (void*)s->st : nullptr; }

q3_status q3_session_synchronize(q3_session* s) {
  Q3_API_BEGIN
  Q3_REQUIRE(s, Q3_ERR_INVALID, "null session");
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  mega2_check_watchdog(s);
  Q3_API_END
}

// prefill over device-resident embeddings x [B][l_max][H] (in scratch x)
static void prefill_run(q3_session* s, const int32_t* lens, int l_max) {
  NvtxRange nvtx("prefill");
  const q3_model* m = s->m;
  const q3_model_desc& d = m->d;
  const int B = s->B, T = B * l_max;
  for (int b = 0; b < B; ++b) {
    Q3_REQUIRE(lens[b] >= 1 && lens[b] <= l_max, Q3_ERR_INVALID, "prefill length out of range");
    if (lens[b] > s->max_seq)
      throw Q3Error(Q3_ERR_KV_OVERFLOW, "KV cache overflow: current=0 + new=" + std::to_string(lens[b]) + " > max=" +
                                            std::to_string(s->max_seq));
  }
  Q3_REQUIRE(l_max <= s->max_seq, Q3_ERR_KV_OVERFLOW, "KV cache overflow: padded prefill exceeds max_seq");
  Q3_CHECK_CUDA(cudaEventRecord(s->ev0, s->st));
  bf16* x = s->sc.x.as<bf16>();
  layers_forward(s, m->tl, m->tdims(), x, T, l_max, nullptr, 0, s->tk_k.as<bf16>(), s->tk_v.as<bf16>(), s->max_seq,
                 s->cos_tab.as<bf16>(), s->sin_tab.as<bf16>());
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->lens_dev.p, lens, B * 4, cudaMemcpyHostToDevice, s->st));
  gather_last_kernel<<<B, 128, 0, s->st>>>(x, s->lens_dev.as<int>(), l_max, d.hidden, s->sc.o.as<bf16>());
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
  talker_head(s, s->sc.o.as<bf16>(), d.hidden);
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->offset.p, lens, B * 4, cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaEventRecord(s->ev1, s->st));
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  Q3_CHECK_CUDA(cudaEventElapsedTime(&s->timing.prefill_ms, s->ev0, s->ev1));
  for (int b = 0; b < B; ++b) s->prefill_len[b] = lens[b];
  s->prefilled = true;
  s->first_sampled = false;
  s->frames_run = 0;
}

q3_status q3_prefill_embeds(q3_session* s, const uint16_t* embeds, const int32_t* lens, int32_t l_max) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && embeds && lens && l_max >= 1, Q3_ERR_INVALID, "bad argument");
  Q3_REQUIRE(!s->prefilled, Q3_ERR_STATE, "session already prefilled; call q3_session_reset first");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  ensure_scratch(s, s->B * l_max);
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->sc.x.p, embeds, (size_t)s->B * l_max * s->m->d.hidden * 2, cudaMemcpyHostToDevice, s->st));
  prefill_run(s, lens, l_max);
  Q3_API_END
}

q3_status q3_prefill_ids(q3_session* s, const int32_t* text_ids, const int32_t* codec_ids, const int32_t* lens, int32_t l_max) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && text_ids && codec_ids && lens && l_max >= 1, Q3_ERR_INVALID, "bad argument");
  Q3_REQUIRE(!s->prefilled, Q3_ERR_STATE, "session already prefilled; call q3_session_reset first");
  Q3_REQUIRE(s->m->text_emb, Q3_ERR_MISSING_WEIGHT, "Missing weight: talker.model.text_embedding.weight");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  const q3_model_desc& d = s->m->d;
  const int T = s->B * l_max;
  for (int i = 0; i < T; ++i) {
    Q3_REQUIRE(text_ids[i] < d.text_vocab && codec_ids[i] < d.codec_vocab, Q3_ERR_INVALID, "token id out of range");
  }
  ensure_scratch(s, T);
  DBuf& tid = s->pf_tid;
  DBuf& cid = s->pf_cid;
  tid.ensure((size_t)T * 4); cid.ensure((size_t)T * 4);
  Q3_CHECK_CUDA(cudaMemcpyAsync(tid.p, text_ids, T * 4, cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaMemcpyAsync(cid.p, codec_ids, T * 4, cudaMemcpyHostToDevice, s->st));
  text_project(s, tid.as<int>(), T, s->sc.o.as<bf16>());
  assemble_embeds_kernel<<<T, 256, 0, s->st>>>(s->sc.o.as<bf16>(), tid.as<int>(), cid.as<int>(), s->m->codec_emb, d.hidden,
                                               s->sc.x.as<bf16>());
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
  prefill_run(s, lens, l_max);
  Q3_API_END
}

q3_status q3_prefill_voice_clone(q3_session* s, const int32_t* text_ids, const int32_t* codec_ids, const int32_t* lens, int32_t l_max,
                                 const uint16_t* speaker_embeds, const uint32_t* ref_codes, const int32_t* t_ref, int32_t t_ref_max) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && text_ids && codec_ids && lens && l_max >= 1 && t_ref_max >= 0, Q3_ERR_INVALID, "bad argument");
  Q3_REQUIRE(!s->prefilled, Q3_ERR_STATE, "session already prefilled; call q3_session_reset first");
  Q3_REQUIRE(s->m->text_emb, Q3_ERR_MISSING_WEIGHT, "Missing weight: talker.model.text_embedding.weight");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  const q3_model_desc& d = s->m->d;
  const int B = s->B, T = B * l_max, H = d.hidden;
  Q3_REQUIRE(d.groups == 16, Q3_ERR_UNSUPPORTED, "reference frames are 16 codes wide");
  for (int b = 0; b < B; ++b) {
    Q3_REQUIRE(lens[b] >= 1 && lens[b] <= l_max, Q3_ERR_INVALID, "prompt length out of range");
    const int tr = (t_ref && t_ref_max > 0) ? t_ref[b] : 0;
    Q3_REQUIRE(tr >= 0 && tr <= t_ref_max, Q3_ERR_INVALID, "reference length out of range");
    for (int p = 0; p < l_max; ++p) {
      const int ti = text_ids[(size_t)b * l_max + p], ci = codec_ids[(size_t)b * l_max + p];
      Q3_REQUIRE(ti < d.text_vocab && ci < d.codec_vocab, Q3_ERR_INVALID, "token id out of range");
      if (p >= lens[b]) continue;
      if (ci == -2) Q3_REQUIRE(speaker_embeds != nullptr, Q3_ERR_INVALID, "speaker position without speaker embeddings");
      else if (ci <= -16) Q3_REQUIRE(ref_codes != nullptr && -16 - ci < tr, Q3_ERR_INVALID, "reference frame index out of range");
      else Q3_REQUIRE(ci >= -1, Q3_ERR_INVALID, "unknown codec-part kind");
    }
    for (int f = 0; f < tr; ++f)
      for (int g = 0; g < 16; ++g) {
        const uint32_t c = ref_codes[((size_t)b * t_ref_max + f) * 16 + g];
        Q3_REQUIRE(c < (uint32_t)(g == 0 ? d.codec_vocab : d.cp_vocab), Q3_ERR_INVALID, "reference code out of range");
      }
  }
  ensure_scratch(s, T);
  DBuf& tid = s->pf_tid;
  DBuf& cid = s->pf_cid;
  tid.ensure((size_t)T * 4); cid.ensure((size_t)T * 4);
  Q3_CHECK_CUDA(cudaMemcpyAsync(tid.p, text_ids, T * 4, cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaMemcpyAsync(cid.p, codec_ids, T * 4, cudaMemcpyHostToDevice, s->st));
  s->pf_spk.ensure((size_t)B * H * 2);
  if (speaker_embeds) Q3_CHECK_CUDA(cudaMemcpyAsync(s->pf_spk.p, speaker_embeds, (size_t)B * H * 2, cudaMemcpyHostToDevice, s->st));
  s->pf_ref.ensure(std::max<size_t>(16, (size_t)B * t_ref_max * 16 * 4));
  if (ref_codes && t_ref_max > 0)
    Q3_CHECK_CUDA(cudaMemcpyAsync(s->pf_ref.p, ref_codes, (size_t)B * t_ref_max * 16 * 4, cudaMemcpyHostToDevice, s->st));
  text_project(s, tid.as<int>(), T, s->sc.o.as<bf16>());
  RefTables tab;
  tab.e[0] = s->m->codec_emb;
  for (int g = 1; g < 16; ++g) tab.e[g] = s->m->cp_emb[g - 1];
  assemble_embeds_ex_kernel<<<T, 256, 0, s->st>>>(s->sc.o.as<bf16>(), tid.as<int>(), cid.as<int>(), tab, s->pf_spk.as<bf16>(),
                                                  s->pf_ref.as<uint32_t>(), t_ref_max, l_max, H, s->sc.x.as<bf16>());
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
  prefill_run(s, lens, l_max);
  Q3_API_END
}

q3_status q3_set_trailing_text(q3_session* s, const uint16_t* trailing, const int32_t* lt, int32_t lt_max, const uint16_t* tts_pad) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && trailing && lt && tts_pad && lt_max >= 1, Q3_ERR_INVALID, "bad argument");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  const int H = s->m->d.hidden;
  for (int b = 0; b < s->B; ++b) Q3_REQUIRE(lt[b] >= 0 && lt[b] <= lt_max, Q3_ERR_INVALID, "trailing length out of range");
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  if (lt_max != s->lt_max || (size_t)s->B * lt_max * H * 2 > s->trailing.bytes) invalidate_graph(s);
  s->trailing.ensure((size_t)s->B * lt_max * H * 2);
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->trailing.p, trailing, (size_t)s->B * lt_max * H * 2, cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->lt.p, lt, s->B * 4, cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->tts_pad.p, tts_pad, (size_t)H * 2, cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  s->lt_max = lt_max;
  s->fs.trailing = s->trailing.as<bf16>();
  s->fs.lt_max = lt_max;
  Q3_API_END
}

q3_status q3_set_trailing_ids(q3_session* s, const int32_t* ids, const int32_t* n, int32_t n_max, int32_t tts_eos_id,
                              int32_t tts_pad_id) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && n && n_max >= 0 && (ids || n_max == 0), Q3_ERR_INVALID, "bad argument");
  Q3_REQUIRE(s->m->text_emb, Q3_ERR_MISSING_WEIGHT, "Missing weight: talker.model.text_embedding.weight");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  const q3_model_desc& d = s->m->d;
  const int B = s->B, H = d.hidden, lt_max = n_max + 1;
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  if (lt_max != s->lt_max || (size_t)B * lt_max * H * 2 > s->trailing.bytes) invalidate_graph(s);
  // rows: ids[b][0..n-1] ++ tts_eos (lib.rs:508-516); one extra row at the end for tts_pad
  std::vector<int> all((size_t)B * lt_max + 1, tts_pad_id), lt(B);
  for (int b = 0; b < B; ++b) {
    Q3_REQUIRE(n[b] >= -1 && n[b] <= n_max, Q3_ERR_INVALID, "trailing length out of range");
    if (n[b] < 0) { lt[b] = 0; continue; }     // no trailing rows at all: every frame adds tts_pad (ICL prompt that consumed the text)
    for (int i = 0; i < n[b]; ++i) {
      int id = ids[(size_t)b * n_max + i];
      Q3_REQUIRE(id >= 0 && id < d.text_vocab, Q3_ERR_INVALID, "text id out of range");
      all[(size_t)b * lt_max + i] = id;
    }
    all[(size_t)b * lt_max + n[b]] = tts_eos_id;
    lt[b] = n[b] + 1;
  }
  const int T = B * lt_max + 1;
  // session-owned staging buffers: a cudaMalloc/cudaFree pair per call cost 70-450 ms once the vocoder workspace existed
  DBuf& idd = s->tr_ids;
  DBuf& proj = s->tr_proj;
  idd.ensure((size_t)T * 4);
  proj.ensure((size_t)T * H * 2);
  Q3_CHECK_CUDA(cudaMemcpyAsync(idd.p, all.data(), T * 4, cudaMemcpyHostToDevice, s->st));
  text_project(s, idd.as<int>(), T, proj.as<bf16>());
  s->trailing.ensure((size_t)B * lt_max * H * 2);
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->trailing.p, proj.p, (size_t)B * lt_max * H * 2, cudaMemcpyDeviceToDevice, s->st));
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->tts_pad.p, proj.as<bf16>() + (size_t)B * lt_max * H, (size_t)H * 2, cudaMemcpyDeviceToDevice, s->st));
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->lt.p, lt.data(), B * 4, cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  s->lt_max = lt_max;
  s->fs.trailing = s->trailing.as<bf16>();
  s->fs.lt_max = lt_max;
  Q3_API_END
}

static int frames_budget(q3_session* s, int max_frames) {
  // the reference loop runs `for frame_idx in 0..max_new_tokens` (lib.rs:580)
  return std::max(0, std::min(max_frames, s->cfg.max_new_tokens - s->frames_run));
}

q3_status q3_generate_async(q3_session* s, int32_t max_frames) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && max_frames >= 0, Q3_ERR_INVALID, "bad argument");
  Q3_REQUIRE(s->prefilled, Q3_ERR_STATE, "q3_generate before prefill");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  Q3_CHECK_CUDA(cudaEventRecord(s->ev0, s->st));
  sample_first_if_needed(s);
  run_frames(s, frames_budget(s, max_frames));
  Q3_CHECK_CUDA(cudaEventRecord(s->ev1, s->st));
  Q3_API_END
}

q3_status q3_get_codes(q3_session* s, int32_t max_frames, uint32_t* codes, int32_t* n_frames) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && codes && n_frames && max_frames >= 0, Q3_ERR_INVALID, "bad argument");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  const int B = s->B;
  Q3_CHECK_CUDA(cudaMemcpyAsync(n_frames, s->n_frames.p, B * 4, cudaMemcpyDeviceToHost, s->st));
  const int take = std::min(max_frames, s->frames_cap);
  if (take > 0)
    Q3_CHECK_CUDA(cudaMemcpy2DAsync(codes, (size_t)max_frames * 64, s->codes.p, (size_t)s->frames_cap * 64, (size_t)take * 64, B,
                                    cudaMemcpyDeviceToHost, s->st));
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  mega2_check_watchdog(s);
  for (int b = 0; b < B; ++b) n_frames[b] = std::min(n_frames[b], max_frames);
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess) s->timing.generation_ms = ms;
  int mx = 0;
  for (int b = 0; b < B; ++b) mx = std::max(mx, n_frames[b]);
  s->timing.generation_frames = mx;
  Q3_API_END
}

q3_status q3_generate(q3_session* s, int32_t max_frames, uint32_t* codes, int32_t* n_frames) {
  q3_status st = q3_generate_async(s, max_frames);
  if (st != Q3_OK) return st;
  return q3_get_codes(s, max_frames, codes, n_frames);
}

// vocode frames [f0, f0+T) of every row (rows shorter than that are padded with their own frame 0.. and zeroed after).
// left_ctx > 0 (opt-in, q3_session_set_stream_context): frames [f0-c, f0) with c = min(left_ctx, f0) are decoded again as
// left context and their samples dropped -- every vocoder op is causal, so with c == f0 the kept samples are exactly the
// ones a single decode of the whole utterance produces.
static void vocode_rows(q3_session* s, int f0, int T, const std::vector<int>& row_len, float* pcm_host, size_t pcm_row_stride,
                        int left_ctx = 0) {
  NvtxRange nvtx("decode");
  const q3_model* m = s->m;
  const int B = s->B, up = vocoder_total_upsample(m);
  if (T <= 0) return;
  const int c = std::max(0, std::min(left_ctx, f0));
  const int f0c = f0 - c, Tc = T + c;
  if (!s->voc) s->voc = m->acquire_scratch();
  DBuf& voc_codes = s->voc->codes;
  DBuf& voc_pcm = s->voc->pcm;
  voc_codes.ensure((size_t)B * 16 * Tc * 8);
  voc_pcm.ensure((size_t)B * Tc * up * 4);
  // rows decode independently (the vocoder is causal per row), so one batched call over T frames and a
  // per-row truncation reproduces B separate Decoder12Hz::decode calls of length row_len[b].
  vocoder_codes_to_tensor(s->codes.as<uint32_t>(), s->frames_cap, f0c, Tc, B, voc_codes.as<long long>(), s->st);
  vocoder_run(m, s->voc->ws, voc_codes.as<long long>(), B, Tc, voc_pcm.as<float>(), s->st);
  if (pcm_host) {
    for (int b = 0; b < B; ++b) {
      const size_t n = (size_t)row_len[b] * up;
      if (n) Q3_CHECK_CUDA(cudaMemcpyAsync(pcm_host + b * pcm_row_stride, voc_pcm.as<float>() + ((size_t)b * Tc + c) * up, n * 4,
                                           cudaMemcpyDeviceToHost, s->st));
    }
    Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
    for (int b = 0; b < B; ++b) {
      const size_t n = (size_t)row_len[b] * up;
      if (n < (size_t)T * up) memset(pcm_host + b * pcm_row_stride + n, 0, ((size_t)T * up - n) * 4);
    }
  }
}

// Stateful streamed decode of frames [f0, f0+T) (q3_session_set_stream_context(sess, -1)): the pre-transformer's keys / values
// and the front half's outputs are carried in the session, so a chunk costs O(chunk + 10 frames) of vocoder work however long
// the utterance is, and the samples equal the ones a single decode of the whole utterance produces (every op is causal; the
// conv stack looks back 9.4 frames).  No reference counterpart: lib.rs:1755-1758 decodes chunks statelessly.
static void vocode_rows_stateful(q3_session* s, int f0, int T, const std::vector<int>& row_len, float* pcm_host, size_t pcm_row_stride) {
  NvtxRange nvtx("decode");
  const q3_model* m = s->m;
  const q3_model_desc& d = m->d;
  const int B = s->B, up = vocoder_total_upsample(m);
  if (T <= 0) return;
  if (!s->voc) s->voc = m->acquire_scratch();
  if (!s->vstream) {
    s->vstream.reset(new VocoderStreamState());
    VocoderStreamState& ss = *s->vstream;
    ss.cap = std::min(s->frames_cap, 3072);
    const size_t kv = (size_t)d.v_layers * B * d.v_heads * ss.cap * d.v_head_dim * sizeof(float);
    ss.kc.alloc(kv); ss.vc.alloc(kv);
    ss.front.alloc((size_t)B * d.v_latent_dim * ss.cap * sizeof(float));
  }
  VocoderStreamState& ss = *s->vstream;
  Q3_REQUIRE(f0 == ss.frames, Q3_ERR_STATE, "stateful streaming: chunks must be decoded in order");
  const int c0 = std::min((int)VOC_STREAM_FRONT_CTX, f0), Tw = c0 + T;
  const int cb = std::min((int)VOC_STREAM_BACK_CTX, f0), Tb = cb + T;
  DBuf& voc_codes = s->voc->codes;
  DBuf& voc_pcm = s->voc->pcm;
  voc_codes.ensure((size_t)B * 16 * Tw * 8);
  voc_pcm.ensure((size_t)B * Tb * up * 4);
  vocoder_codes_to_tensor(s->codes.as<uint32_t>(), s->frames_cap, f0 - c0, Tw, B, voc_codes.as<long long>(), s->st);
  vocoder_stream_chunk(m, s->voc->ws, ss, voc_codes.as<long long>(), B, f0, T, voc_pcm.as<float>(), s->st);
  if (pcm_host) {
    for (int b = 0; b < B; ++b) {
      const size_t n = (size_t)row_len[b] * up;
      if (n) Q3_CHECK_CUDA(cudaMemcpyAsync(pcm_host + b * pcm_row_stride, voc_pcm.as<float>() + ((size_t)b * Tb + cb) * up, n * 4,
                                           cudaMemcpyDeviceToHost, s->st));
    }
    Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
    for (int b = 0; b < B; ++b) {
      const size_t n = (size_t)row_len[b] * up;
      if (n < (size_t)T * up) memset(pcm_host + b * pcm_row_stride + n, 0, ((size_t)T * up - n) * 4);
    }
  }
}

q3_status q3_session_set_stream_context(q3_session* s, int32_t left_context_frames) {
  Q3_API_BEGIN
  Q3_REQUIRE(s, Q3_ERR_INVALID, "null session");
  s->stream_left_ctx = left_context_frames < 0 ? INT32_MAX : left_context_frames;
  Q3_API_END
}

q3_status q3_session_set_first_chunk(q3_session* s, int32_t first_chunk_frames) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && first_chunk_frames >= 0, Q3_ERR_INVALID, "bad argument");
  s->stream_first = first_chunk_frames;
  Q3_API_END
}

q3_status q3_vocode_session(q3_session* s, int32_t max_frames, float* pcm) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && max_frames >= 0, Q3_ERR_INVALID, "bad argument");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  std::vector<int> nf(s->B);
  Q3_CHECK_CUDA(cudaMemcpyAsync(nf.data(), s->n_frames.p, s->B * 4, cudaMemcpyDeviceToHost, s->st));
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  int T = 0;
  for (int b = 0; b < s->B; ++b) { nf[b] = std::min(nf[b], max_frames); T = std::max(T, nf[b]); }
  Q3_CHECK_CUDA(cudaEventRecord(s->ev0, s->st));
  const int up = vocoder_total_upsample(s->m);
  if (pcm && T < max_frames)
    for (int b = 0; b < s->B; ++b) memset(pcm + (size_t)b * max_frames * up, 0, (size_t)max_frames * up * 4);
  vocode_rows(s, 0, T, nf, pcm, (size_t)max_frames * up);
  Q3_CHECK_CUDA(cudaEventRecord(s->ev1, s->st));
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  Q3_CHECK_CUDA(cudaEventElapsedTime(&s->timing.decode_ms, s->ev0, s->ev1));
  Q3_API_END
}

q3_status q3_stream_next(q3_session* s, uint32_t* codes, float* pcm, int32_t* n_frames, int32_t* done) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && codes && pcm && n_frames && done, Q3_ERR_INVALID, "null argument");
  Q3_REQUIRE(s->prefilled, Q3_ERR_STATE, "q3_stream_next before prefill");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  const int B = s->B, chunk = std::max(1, s->cfg.chunk_frames), up = vocoder_total_upsample(s->m);
  sample_first_if_needed(s);
  // opt-in (q3_session_set_first_chunk): the first chunk of the stream is shorter, for a low time to first audio
  const int gen = (s->stream_first > 0 && s->frames_run == 0) ? std::min(chunk, s->stream_first) : chunk;
  run_frames(s, frames_budget(s, gen));
  std::vector<int> nf(B), dn(B);
  Q3_CHECK_CUDA(cudaMemcpyAsync(nf.data(), s->n_frames.p, B * 4, cudaMemcpyDeviceToHost, s->st));
  Q3_CHECK_CUDA(cudaMemcpyAsync(dn.data(), s->done.p, B * 4, cudaMemcpyDeviceToHost, s->st));
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  // every row emits the frames it produced since the last chunk; all rows share the same f0 only while
  // none has finished, so decode per distinct start offset.
  bool all_done = true;
  std::vector<int> len(B);
  int T = 0;
  for (int b = 0; b < B; ++b) {
    len[b] = std::min(nf[b] - s->stream_emitted[b], chunk);
    n_frames[b] = len[b];
    T = std::max(T, len[b]);
    const bool row_done = dn[b] || nf[b] >= s->cfg.max_new_tokens || s->frames_run >= s->cfg.max_new_tokens;
    all_done = all_done && row_done;
  }
  memset(codes, 0, (size_t)B * chunk * 64);
  memset(pcm, 0, (size_t)B * chunk * up * 4);
  if (T > 0) {
    // rows that are still running all have stream_emitted == frames emitted before this call
    int f0 = -1;
    bool same = true;
    for (int b = 0; b < B; ++b)
      if (len[b] > 0) { if (f0 < 0) f0 = s->stream_emitted[b]; else same = same && (f0 == s->stream_emitted[b]); }
    Q3_REQUIRE(same, Q3_ERR_STATE, "streaming rows out of step");
    if (s->stream_left_ctx == INT32_MAX && s->frames_cap <= 3072) vocode_rows_stateful(s, f0, T, len, pcm, (size_t)chunk * up);
    else vocode_rows(s, f0, T, len, pcm, (size_t)chunk * up, s->stream_left_ctx);
    for (int b = 0; b < B; ++b)
      if (len[b] > 0)
        Q3_CHECK_CUDA(cudaMemcpyAsync(codes + (size_t)b * chunk * 16, s->codes.as<uint32_t>() + ((size_t)b * s->frames_cap + f0) * 16,
                                      (size_t)len[b] * 64, cudaMemcpyDeviceToHost, s->st));
    Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
    for (int b = 0; b < B; ++b) s->stream_emitted[b] += len[b];
  }
  *done = all_done ? 1 : 0;
  Q3_API_END
}

q3_status q3_vocoder_decode(const q3_model* m, const int64_t* codes, int32_t batch, int32_t t, float* pcm) {
  Q3_API_BEGIN
  Q3_REQUIRE(m && codes && pcm && batch >= 0 && t >= 0, Q3_ERR_INVALID, "bad argument");
  Q3_REQUIRE(m->finalized && m->has_vocoder, Q3_ERR_STATE, "model has no finalized vocoder weights");
  if (batch == 0 || t == 0) return Q3_OK;       // codes_to_tensor of zero frames -> empty waveform
  {
    // the reference's index_select fails on an out-of-range code (decoder_12hz.rs:429, 443): a caller bug must not become
    // plausible-sounding audio.  Semantic codes are reduced modulo the codebook size (decoder_12hz.rs:423-427), so only a
    // negative one is an error there (Rust's % keeps the sign and index_select then fails).
    const int nq = m->d.v_quantizers;
    const int64_t cb = m->d.v_codebook_size;
    for (int b = 0; b < batch; ++b)
      for (int q = 0; q < nq; ++q) {
        const int64_t* row = codes + ((size_t)b * nq + q) * t;
        for (int f = 0; f < t; ++f) {
          const int64_t c = row[f];
          if (c < 0 || (q > 0 && c >= cb))
            throw Q3Error(Q3_ERR_INVALID, "code out of range: batch " + std::to_string(b) + " quantizer " + std::to_string(q) +
                                              " frame " + std::to_string(f) + " value " + std::to_string((long long)c) +
                                              " (codebook size " + std::to_string((long long)cb) + ")");
        }
      }
  }
  Q3_CHECK_CUDA(cudaSetDevice(m->d.device));
  std::lock_guard<std::mutex> lock(m->voc_mutex);
  const int up = vocoder_total_upsample(m);
  DBuf dc, dp;
  dc.alloc((size_t)batch * m->d.v_quantizers * t * 8);
  dp.alloc((size_t)batch * t * up * 4);
  Q3_CHECK_CUDA(cudaMemcpy(dc.p, codes, dc.bytes, cudaMemcpyHostToDevice));
  vocoder_run(m, m->voc_ws, dc.as<long long>(), batch, t, dp.as<float>(), 0);
  Q3_CHECK_CUDA(cudaMemcpy(pcm, dp.p, (size_t)batch * t * up * 4, cudaMemcpyDeviceToHost));
  Q3_API_END
}

int32_t q3_speaker_embed_dim(const q3_model* m) { return (m && m->finalized && m->has_speaker) ? m->spk.enc_dim : 0; }

q3_status q3_speaker_encode(const q3_model* m, const float* mel, int32_t batch, int32_t t, float* embed_out) {
  Q3_API_BEGIN
  Q3_REQUIRE(m && mel && embed_out && batch >= 0 && t >= 1, Q3_ERR_INVALID, "bad argument");
  Q3_REQUIRE(m->finalized && m->has_speaker, Q3_ERR_STATE, "model has no finalized speaker-encoder weights");
  Q3_CHECK_CUDA(cudaSetDevice(m->d.device));
  std::lock_guard<std::mutex> lock(m->voc_mutex);
  const size_t n_mel = (size_t)m->spk.mel * t;
  DBuf dm, de;
  dm.alloc(n_mel * 4);
  de.alloc((size_t)m->spk.enc_dim * 4);
  for (int b = 0; b < batch; ++b) {          // one utterance at a time, as SpeakerEncoder::encode is called (speaker.rs:436-445)
    Q3_CHECK_CUDA(cudaMemcpy(dm.p, mel + (size_t)b * n_mel, n_mel * 4, cudaMemcpyHostToDevice));
    speaker_run(m, dm.as<float>(), t, de.as<float>(), 0);
    Q3_CHECK_CUDA(cudaMemcpy(embed_out + (size_t)b * m->spk.enc_dim, de.p, (size_t)m->spk.enc_dim * 4, cudaMemcpyDeviceToHost));
  }
  Q3_API_END
}

// ---- fine-grained entry points ------------------------------------------------------------------
q3_status q3_talker_step(q3_session* s, const uint16_t* step_input, uint16_t* hidden_out, float* logits_out) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && step_input, Q3_ERR_INVALID, "null argument");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  const q3_model_desc& d = s->m->d;
  int max_len = 0;
  for (int b = 0; b < s->B; ++b) max_len = std::max(max_len, s->prefill_len[b]);
  if (max_len + s->frames_run + 1 > s->max_seq)
    throw Q3Error(Q3_ERR_KV_OVERFLOW, "KV cache overflow: current=" + std::to_string(max_len + s->frames_run) +
                                          " + new=1 > max=" + std::to_string(s->max_seq));
  ensure_scratch(s, 2 * s->B);
  bf16* x = s->sc.x.as<bf16>();
  if (s->use_mega && s->B <= MEGA_TMAX) {
    Q3_CHECK_CUDA(cudaMemcpyAsync(s->step_input.p, step_input, (size_t)s->B * d.hidden * 2, cudaMemcpyHostToDevice, s->st));
    if (s->mega_ver >= 2) {
      mega2_run_mode(s, false, false, true, false, s->step_input.as<bf16>(), nullptr);
    } else {
      MegaArgs a = mega_args(s);
      a.n_frames = 1; a.do_talker = 1; a.ext_step_input = s->step_input.as<bf16>();
      mega_launch(s, a);
    }
  } else {
    Q3_CHECK_CUDA(cudaMemcpyAsync(x, step_input, (size_t)s->B * d.hidden * 2, cudaMemcpyHostToDevice, s->st));
    layers_forward(s, s->m->tl, s->m->tdims(), x, s->B, 1, s->fs.offset, 0, s->tk_k.as<bf16>(), s->tk_v.as<bf16>(), s->max_seq,
                   s->cos_tab.as<bf16>(), s->sin_tab.as<bf16>());
    talker_head(s, x, d.hidden);
  }
  add_int_kernel<<<1, 256, 0, s->st>>>(s->fs.offset, s->B, 1);
  Q3_COUNT_LAUNCH();
  s->frames_run += 1;
  if (hidden_out) Q3_CHECK_CUDA(cudaMemcpyAsync(hidden_out, s->last_hidden.p, (size_t)s->B * d.hidden * 2, cudaMemcpyDeviceToHost, s->st));
  if (logits_out) Q3_CHECK_CUDA(cudaMemcpyAsync(logits_out, s->logits.p, (size_t)s->B * d.codec_vocab * 4, cudaMemcpyDeviceToHost, s->st));
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  mega2_check_watchdog(s);
  Q3_API_END
}

q3_status q3_code_predictor_frame(q3_session* s, const uint16_t* last_hidden, const uint32_t* sem_tokens, uint32_t* codes_out,
                                  float* logits_out) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && last_hidden && sem_tokens && codes_out, Q3_ERR_INVALID, "null argument");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  const q3_model_desc& d = s->m->d;
  const int B = s->B, n_ac = d.groups - 1;
  for (int b = 0; b < B; ++b) Q3_REQUIRE((int)sem_tokens[b] < d.codec_vocab, Q3_ERR_INVALID, "semantic token out of range");
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->last_hidden.p, last_hidden, (size_t)B * d.hidden * 2, cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->cur_tok.p, sem_tokens, B * 4, cudaMemcpyHostToDevice, s->st));
  if (logits_out) s->cp_logits.ensure((size_t)n_ac * B * d.cp_vocab * 4);
  if (s->use_mega && s->B <= MEGA_TMAX) {
    ensure_scratch(s, 2 * B);
    if (s->mega_ver >= 2) {
      mega2_run_mode(s, true, false, false, false, nullptr, logits_out ? s->cp_logits.as<float>() : nullptr);
    } else {
      MegaArgs a = mega_args(s);
      a.n_frames = 1; a.do_cp = 1; a.cp_logits = logits_out ? s->cp_logits.as<float>() : nullptr;
      mega_launch(s, a);
    }
  } else {
    cp_frame(s, logits_out ? s->cp_logits.as<float>() : nullptr);
  }
  std::vector<unsigned long long> keys((size_t)n_ac * B);
  Q3_CHECK_CUDA(cudaMemcpyAsync(keys.data(), s->amax.p, keys.size() * 8, cudaMemcpyDeviceToHost, s->st));
  std::vector<float> lg;
  if (logits_out) {
    lg.resize((size_t)n_ac * B * d.cp_vocab);
    Q3_CHECK_CUDA(cudaMemcpyAsync(lg.data(), s->cp_logits.p, lg.size() * 4, cudaMemcpyDeviceToHost, s->st));
  }
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  for (int b = 0; b < B; ++b)
    for (int g = 0; g < n_ac; ++g) {
      codes_out[b * n_ac + g] = argmax_key_index(keys[(size_t)g * B + b]);
      if (logits_out)
        memcpy(logits_out + ((size_t)b * n_ac + g) * d.cp_vocab, lg.data() + ((size_t)g * B + b) * d.cp_vocab, (size_t)d.cp_vocab * 4);
    }
  Q3_API_END
}

q3_status q3_sample(const q3_model* m, const float* logits, int32_t batch, int32_t vocab, const q3_gen_config* cfg,
                    uint64_t* rng_states, uint8_t* seen_mask, int32_t token_count, uint32_t* tokens_out) {
  Q3_API_BEGIN
  Q3_REQUIRE(logits && cfg && rng_states && seen_mask && tokens_out && batch >= 1, Q3_ERR_INVALID, "bad argument");
  Q3_REQUIRE(vocab > 1024 && vocab <= 4096, Q3_ERR_UNSUPPORTED, "vocab must be in (1024, 4096]");
  if (m) Q3_CHECK_CUDA(cudaSetDevice(m->d.device));
  DBuf dl, ds, dr, dt;
  dl.alloc((size_t)batch * vocab * 4); ds.alloc((size_t)batch * vocab); dr.alloc(batch * 8); dt.alloc(batch * 4);
  Q3_CHECK_CUDA(cudaMemcpy(dl.p, logits, dl.bytes, cudaMemcpyHostToDevice));
  Q3_CHECK_CUDA(cudaMemcpy(ds.p, seen_mask, (size_t)batch * vocab, cudaMemcpyHostToDevice));
  Q3_CHECK_CUDA(cudaMemcpy(dr.p, rng_states, batch * 8, cudaMemcpyHostToDevice));
  SampleArgs a = make_sample_args(*cfg, vocab, batch);
  a.logits = dl.as<float>(); a.seen = ds.as<uint8_t>(); a.rng = dr.as<unsigned long long>(); a.tok_out = dt.as<uint32_t>();
  a.token_count = nullptr; a.token_count_imm = token_count; a.done = nullptr; a.offset = nullptr; a.frame_idx = nullptr;
  sample_kernel<<<batch, 1024>>>(a);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
  Q3_CHECK_CUDA(cudaMemcpy(tokens_out, dt.p, batch * 4, cudaMemcpyDeviceToHost));
  Q3_CHECK_CUDA(cudaMemcpy(seen_mask, ds.p, (size_t)batch * vocab, cudaMemcpyDeviceToHost));
  Q3_CHECK_CUDA(cudaMemcpy(rng_states, dr.p, batch * 8, cudaMemcpyDeviceToHost));
  Q3_API_END
}

q3_status q3_fused_residual_rmsnorm(const void* x, const void* r, const void* w, void* out_normed, void* out_sum, int32_t rows,
                                    int32_t cols, float eps, q3_dtype dtype, void* stream) {
  Q3_API_BEGIN
  Q3_REQUIRE(x && r && w && out_normed && out_sum && rows >= 0 && cols >= 1, Q3_ERR_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == Q3_BF16)
    fused_residual_rmsnorm_launch<bf16>((const bf16*)x, (const bf16*)r, (const bf16*)w, (bf16*)out_normed, (bf16*)out_sum, rows, cols, eps, st);
  else if (dtype == Q3_F32)
    fused_residual_rmsnorm_launch<float>((const float*)x, (const float*)r, (const float*)w, (float*)out_normed, (float*)out_sum, rows, cols, eps, st);
  else
    throw Q3Error(Q3_ERR_UNSUPPORTED, "fused-residual-rmsnorm unsupported dtype");
  Q3_API_END
}

q3_status q3_fused_residual_rmsnorm_host(const void* x, const void* r, const void* w, void* out_normed, void* out_sum,
                                         int32_t rows, int32_t cols, float eps, q3_dtype dtype, int32_t device) {
  Q3_API_BEGIN
  Q3_REQUIRE(x && r && w && out_normed && out_sum && rows >= 0 && cols >= 1, Q3_ERR_INVALID, "bad argument");
  Q3_CHECK_CUDA(cudaSetDevice(device));
  const size_t es = dtype == Q3_BF16 ? 2 : 4, n = (size_t)rows * cols * es;
  DBuf dx, dr, dw, dn, dsum;
  dx.alloc(n); dr.alloc(n); dw.alloc(cols * es); dn.alloc(n); dsum.alloc(n);
  Q3_CHECK_CUDA(cudaMemcpy(dx.p, x, n, cudaMemcpyHostToDevice));
  Q3_CHECK_CUDA(cudaMemcpy(dr.p, r, n, cudaMemcpyHostToDevice));
  Q3_CHECK_CUDA(cudaMemcpy(dw.p, w, cols * es, cudaMemcpyHostToDevice));
  q3_status st = q3_fused_residual_rmsnorm(dx.p, dr.p, dw.p, dn.p, dsum.p, rows, cols, eps, dtype, nullptr);
  if (st != Q3_OK) return st;
  Q3_CHECK_CUDA(cudaDeviceSynchronize());
  Q3_CHECK_CUDA(cudaMemcpy(out_normed, dn.p, n, cudaMemcpyDeviceToHost));
  Q3_CHECK_CUDA(cudaMemcpy(out_sum, dsum.p, n, cudaMemcpyDeviceToHost));
  Q3_API_END
}

// Test aid (not part of the drop-in boundary): q3_generate with the loop's decision inputs tapped every frame, so a test can
// replay the reference's sampler on the very logits this path sampled from and compare every tensor of a FREE-RUNNING run
// with an oracle that follows the emitted codes (oracle/generate.py follow()).  Same kernels and phase program as
// q3_generate; frames run one per launch so that the taps can be read back between them (the codes are bit-identical to
// an untapped run -- tests/test_gpu_parity.py asserts it).  Host buffers, any of them may be NULL:
//   first_logits f32 [B][V]            prefill logits token 0 is sampled from (lib.rs:557-571)
//   logits       f32 [F][B][V]         raw talker logits of frame f (lib.rs:627-631), before penalties
//   cp_logits    f32 [F][15][B][cpV]   code-predictor logits of every pass (code_predictor.rs:357-413)
//   rng          u64 [F+2][B]          PCG state before the first draw, then after every draw
//   step_input   bf16 [F][B][H]        talker input of frame f (lib.rs:612-622)
q3_status q3_debug_generate_tapped(q3_session* s, int32_t max_frames, uint32_t* codes, int32_t* n_frames, float* first_logits,
                                   float* logits, float* cp_logits, uint64_t* rng, uint16_t* step_input) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && codes && n_frames && max_frames >= 0, Q3_ERR_INVALID, "bad argument");
  Q3_REQUIRE(s->prefilled && !s->first_sampled, Q3_ERR_STATE, "tapped generation needs a freshly prefilled session");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  const q3_model_desc& d = s->m->d;
  const size_t B = s->B, V = d.codec_vocab, cpV = d.cp_vocab, H = d.hidden, n_ac = d.groups - 1;
  if (cp_logits) {
    s->cp_logits.ensure(n_ac * B * cpV * 4);
    s->tap_cp_logits = s->cp_logits.as<float>();
    s->m2_n_ph = 0;                 // the cached frame program does not write logits: rebuild with the tap
    invalidate_graph(s);
  }
  struct Untap {
    q3_session* s;
    ~Untap() { if (s->tap_cp_logits) { s->tap_cp_logits = nullptr; s->m2_n_ph = 0; invalidate_graph(s); } }
  } untap{s};
  if (first_logits) Q3_CHECK_CUDA(cudaMemcpyAsync(first_logits, s->logits.p, B * V * 4, cudaMemcpyDeviceToHost, s->st));
  if (rng) Q3_CHECK_CUDA(cudaMemcpyAsync(rng, s->rng.p, B * 8, cudaMemcpyDeviceToHost, s->st));
  sample_first_if_needed(s);
  if (rng) Q3_CHECK_CUDA(cudaMemcpyAsync(rng + B, s->rng.p, B * 8, cudaMemcpyDeviceToHost, s->st));
  const int budget = frames_budget(s, max_frames);
  std::vector<int> dn(B);
  for (int f = 0; f < budget; ++f) {
    Q3_CHECK_CUDA(cudaMemcpyAsync(dn.data(), s->done.p, B * 4, cudaMemcpyDeviceToHost, s->st));
    Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
    bool all_done = true;
    for (size_t b = 0; b < B; ++b) all_done = all_done && dn[b] != 0;
    if (all_done) break;
    run_frames(s, 1);
    if (logits) Q3_CHECK_CUDA(cudaMemcpyAsync(logits + (size_t)f * B * V, s->logits.p, B * V * 4, cudaMemcpyDeviceToHost, s->st));
    if (cp_logits)
      Q3_CHECK_CUDA(cudaMemcpyAsync(cp_logits + (size_t)f * n_ac * B * cpV, s->cp_logits.p, n_ac * B * cpV * 4, cudaMemcpyDeviceToHost, s->st));
    if (rng) Q3_CHECK_CUDA(cudaMemcpyAsync(rng + (size_t)(f + 2) * B, s->rng.p, B * 8, cudaMemcpyDeviceToHost, s->st));
    if (step_input)
      Q3_CHECK_CUDA(cudaMemcpyAsync(step_input + (size_t)f * B * H, s->step_input.p, B * H * 2, cudaMemcpyDeviceToHost, s->st));
  }
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  mega2_check_watchdog(s);
  Q3_CHECK_CUDA(cudaEventRecord(s->ev0, s->st));
  Q3_CHECK_CUDA(cudaEventRecord(s->ev1, s->st));
  return q3_get_codes(s, max_frames, codes, n_frames);
  Q3_API_END
}

// Test aid: the session's trailing-text rows as the DEVICE built them (q3_set_trailing_ids: text projection GEMM) -- the
// second operand of the talker-input add (lib.rs:617-621) -- so a test can check that add bit-exactly without inheriting the
// rounding noise of the projection.  trailing: bf16 [B][cap][H] (rows past lt[b] untouched), lt: [B], tts_pad: bf16 [H].
q3_status q3_debug_get_trailing(q3_session* s, uint16_t* trailing, int32_t cap, int32_t* lt, uint16_t* tts_pad) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && trailing && lt && tts_pad && cap >= 1, Q3_ERR_INVALID, "bad argument");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  const size_t H = s->m->d.hidden;
  const int rows = std::min(cap, s->lt_max);
  Q3_CHECK_CUDA(cudaMemcpyAsync(lt, s->lt.p, s->B * 4, cudaMemcpyDeviceToHost, s->st));
  Q3_CHECK_CUDA(cudaMemcpyAsync(tts_pad, s->tts_pad.p, H * 2, cudaMemcpyDeviceToHost, s->st));
  Q3_CHECK_CUDA(cudaMemcpy2DAsync(trailing, (size_t)cap * H * 2, s->fs.trailing, (size_t)s->lt_max * H * 2, (size_t)rows * H * 2, s->B,
                                  cudaMemcpyDeviceToHost, s->st));
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  Q3_API_END
}

// Debug aid: which decode engine the session runs on: 0 = multi-kernel CUDA graph, 1 / 2 / 3 / 4 = generation of the persistent
// frame kernel (mega.cuh, mega2.cuh, mega3.cuh, mega4.cuh).  Tests use it to make sure a requested generation was not silently
// replaced by a fallback.
int q3_debug_decode_generation(q3_session* s) { return (s && s->use_mega) ? s->mega_ver : 0; }

// Debug aid (race hunting): the first `n_ph` phases of the code-predictor program of one frame, then the raw tagged
// activation buffers (8-byte slots {payload, tag}) copied back.  sizes in bytes: x 16*H*4, qkv 16*nh*4, attn 16*qd*4,
// h1 16*H*8, act 16*I*4 with H/I the larger of the talker / code-predictor dimensions (see q3_session_create).
q3_status q3_debug_cp_prefix(q3_session* s, const uint16_t* last_hidden, const uint32_t* sem_tokens, int32_t n_ph, void* x_out,
                             void* qkv_out, void* attn_out, void* h1_out, void* act_out) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && last_hidden && sem_tokens && s->use_mega && s->mega_ver >= 2 && s->B <= MEGA_TMAX, Q3_ERR_STATE, "needs the persistent path");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  const q3_model_desc& d = s->m->d;
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->last_hidden.p, last_hidden, (size_t)s->B * d.hidden * 2, cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->cur_tok.p, sem_tokens, s->B * 4, cudaMemcpyHostToDevice, s->st));
  ensure_scratch(s, 2 * s->B);
  std::vector<M2Phase> pr = m2_build_program(s, true, false, false, false, nullptr, nullptr);
  if (n_ph > 0 && n_ph < (int)pr.size()) pr.resize(n_ph);
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  s->m2_prog_tmp.ensure(pr.size() * sizeof(M2Phase));
  Q3_CHECK_CUDA(cudaMemcpyAsync(s->m2_prog_tmp.p, pr.data(), pr.size() * sizeof(M2Phase), cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  M2Args a = mega2_args(s);
  a.n_frames = 1; a.do_sample = 0;
  mega2_launch(s, a, s->m2_prog_tmp, (int)pr.size());
  if (x_out) Q3_CHECK_CUDA(cudaMemcpyAsync(x_out, s->m2_x.p, s->m2_x.bytes, cudaMemcpyDeviceToHost, s->st));
  if (qkv_out) Q3_CHECK_CUDA(cudaMemcpyAsync(qkv_out, s->m2_qkv.p, s->m2_qkv.bytes, cudaMemcpyDeviceToHost, s->st));
  if (attn_out) Q3_CHECK_CUDA(cudaMemcpyAsync(attn_out, s->m2_attn.p, s->m2_attn.bytes, cudaMemcpyDeviceToHost, s->st));
  if (h1_out) Q3_CHECK_CUDA(cudaMemcpyAsync(h1_out, s->m2_h1.p, s->m2_h1.bytes, cudaMemcpyDeviceToHost, s->st));
  if (act_out) Q3_CHECK_CUDA(cudaMemcpyAsync(act_out, s->m2_act.p, s->m2_act.bytes, cudaMemcpyDeviceToHost, s->st));
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  mega2_check_watchdog(s);
  Q3_API_END
}

// Debug aid (not part of the drop-in boundary): run `frames` frames with %globaltimer stamps of block 0 and return
// them as (time_ns << 8 | tag).  Used by tools/profile_mega.py only.
q3_status q3_debug_profile(q3_session* s, int32_t frames, uint64_t* stamps, int32_t cap, int32_t* n_out) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && stamps && n_out && s->use_mega && s->prefilled, Q3_ERR_STATE, "profile needs a prefilled session on the persistent path");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  sample_first_if_needed(s);
  s->prof.alloc((size_t)cap * 8);
  s->prof.zero(s->st);
  unsigned zero = 0;
  Q3_CHECK_CUDA(cudaMemcpyToSymbolAsync(g_prof_idx, &zero, 4, 0, cudaMemcpyHostToDevice, s->st));
  Q3_CHECK_CUDA(cudaMemcpyToSymbolAsync(g_prof2_idx, &zero, 4, 0, cudaMemcpyHostToDevice, s->st));
  run_frames(s, frames);
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  unsigned n = 0;
  if (s->mega_ver >= 2) Q3_CHECK_CUDA(cudaMemcpyFromSymbol(&n, g_prof2_idx, 4));
  else Q3_CHECK_CUDA(cudaMemcpyFromSymbol(&n, g_prof_idx, 4));
  {
    const char* pm = std::getenv("Q3_PROF_MODE");
    if (pm && std::atoi(pm) == 2) n = (unsigned)cap;
  }
  *n_out = (int)std::min<unsigned>(n, (unsigned)cap);
  Q3_CHECK_CUDA(cudaMemcpy(stamps, s->prof.p, (size_t)(*n_out) * 8, cudaMemcpyDeviceToHost));
  s->prof.release();
  Q3_API_END
}

// Debug aid: time `n` back-to-back grid barriers of the persistent kernel (ms for the whole launch).
q3_status q3_debug_barrier_bench(q3_session* s, int32_t n, float* ms_out) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && ms_out && s->use_mega, Q3_ERR_STATE, "needs the persistent path");
  Q3_CHECK_CUDA(cudaSetDevice(s->m->d.device));
  if (s->mega_ver >= 2) {
    // odd n: release/acquire barriers, even n: relaxed (hint) barriers
    std::vector<M2Phase> pr = m2_build_program(s, false, false, false, false, nullptr, nullptr);
    s->m2_prog_tmp.ensure(pr.size() * sizeof(M2Phase));
    Q3_CHECK_CUDA(cudaMemcpy(s->m2_prog_tmp.p, pr.data(), pr.size() * sizeof(M2Phase), cudaMemcpyHostToDevice));
    M2Args a = mega2_args(s);
    a.bench_barriers = n;
    mega2_launch(s, a, s->m2_prog_tmp, (int)pr.size(), true);
    Q3_CHECK_CUDA(cudaEventRecord(s->ev0, s->st));
    mega2_launch(s, a, s->m2_prog_tmp, (int)pr.size(), true);
    Q3_CHECK_CUDA(cudaEventRecord(s->ev1, s->st));
    Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
    Q3_CHECK_CUDA(cudaEventElapsedTime(ms_out, s->ev0, s->ev1));
    return Q3_OK;
  }
  MegaArgs a = mega_args(s);
  a.bench_barriers = n;
  mega_launch(s, a);
  Q3_CHECK_CUDA(cudaEventRecord(s->ev0, s->st));
  mega_launch(s, a);
  Q3_CHECK_CUDA(cudaEventRecord(s->ev1, s->st));
  Q3_CHECK_CUDA(cudaStreamSynchronize(s->st));
  Q3_CHECK_CUDA(cudaEventElapsedTime(ms_out, s->ev0, s->ev1));
  Q3_API_END
}

q3_status q3_session_timing(q3_session* s, q3_timing* out) {
  Q3_API_BEGIN
  Q3_REQUIRE(s && out, Q3_ERR_INVALID, "null argument");
  *out = s->timing;
  Q3_API_END
}

}  // extern "C"
