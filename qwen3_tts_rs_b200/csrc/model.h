// Host-side model / session objects behind the opaque C handles.
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

// Size-keyed cache of freed device buffers (per device).  A session allocates ~40 buffers (0.5 GB of KV cache at
// batch 8 / 512 positions) and the reference-style API creates one session per synthesize call; cudaMalloc/cudaFree
// of those cost 100-400 ms per call and serialise the device.  Buffers come back in an undefined state: callers that
// need zeros call zero() (they already did).  A buffer is only released after its session's stream was synchronised.
struct DevPool {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> free_;
  size_t cached = 0;
  static constexpr size_t kLimit = (size_t)24 << 30;
  static DevPool& get() { static DevPool p; return p; }
  void* take(int dev, size_t n) {
    std::lock_guard<std::mutex> lock(mu);
    auto it = free_.find({dev, n});
    if (it == free_.end()) return nullptr;
    void* p = it->second;
    free_.erase(it);
    cached -= n;
    return p;
  }
  bool give(int dev, void* p, size_t n) {
    std::lock_guard<std::mutex> lock(mu);
    if (cached + n > kLimit) return false;
    free_.emplace(std::make_pair(dev, n), p);
    cached += n;
    return true;
  }
};

struct DBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int dev = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), bytes(o.bytes), dev(o.dev) { o.p = nullptr; o.bytes = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; bytes = o.bytes; dev = o.dev; o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  void alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    Q3_CHECK_CUDA(cudaGetDevice(&dev));
    p = DevPool::get().take(dev, n);
    if (p == nullptr) Q3_CHECK_CUDA(cudaMalloc(&p, n));
    bytes = n;
  }
  void ensure(size_t n) { if (n > bytes) alloc(n); }
  void zero(cudaStream_t st = 0) { if (p) Q3_CHECK_CUDA(cudaMemsetAsync(p, 0, bytes, st)); }
  void release() {
    if (p && !DevPool::get().give(dev, p, bytes)) cudaFree(p);
    p = nullptr; bytes = 0;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct RawTensor {
  DBuf buf;
  std::vector<int64_t> shape;
  q3_dtype dtype;
  size_t numel() const { size_t n = 1; for (auto d : shape) n *= (size_t)d; return n; }
};

// ---- talker / code predictor ---------------------------------------------------------------------
struct LayerW {
  const bf16 *in_ln, *wqkv, *wo, *q_norm, *k_norm, *post_ln, *gate, *up, *down;
};
struct StackDims { int H, I, heads, kv_heads, layers; };

// ---- vocoder ----------------------------------------------------------------------------------------
struct VConv {
  const float* w = nullptr;       // SIMT layout [Cin*k][Cout]
  const float* b = nullptr;
  int cin = 0, cout = 0, k = 1;
  const bf16 *w_hi = nullptr, *w_lo = nullptr;   // tensor-core layout: bf16 hi/lo split, [k][chunks][Cout_pad][32]
  const bf16* w_um = nullptr;                    // tcgen05 layout (vocoder_umma.cuh): 16 KB shared-memory image per (tap, chunk, 128 rows)
  int cout_pad = 0, chunks = 0;
  int um_rows_pad = 0;            // rows of the tcgen05 image (cout_pad; cout * stride padded for a transposed conv)
};
struct VSnake { const float* ea = nullptr; const float* ib = nullptr; };
struct VLayer { const float *in_ln, *post_ln, *attn_scale, *mlp_scale; VConv q, k, v, o, gate, up, down; };
struct VConvNext { const float *dw_w, *dw_b, *ln_w, *ln_b, *gamma; VConv pw1, pw2; int C; };
struct VResUnit { VSnake a1, a2; VConv c1, c2; int dil; };
struct VBlock { VSnake s; VConv up; int rate; VResUnit ru[3]; };
struct VUpsample { VConv tconv; int ratio; VConvNext cn; };
struct VocoderW {
  const float* first_cb = nullptr;   // [size][vq]
  const float* rest_cb = nullptr;    // [nq-1][size][vq]
  VConv first_proj, rest_proj, pre_conv, in_proj, out_proj, init_conv, final_conv;
  std::vector<VLayer> layers;
  const float* final_norm = nullptr;
  std::vector<VUpsample> ups;
  std::vector<VBlock> blocks;
  VSnake final_snake;
};

// ECAPA-TDNN speaker encoder (speaker.rs:352-434).  k = 1 convs run on the vocoder's tensor-core conv kernels (VConv);
// the reflect-padded k > 1 convs keep the checkpoint layout [Cout][Cin][k].
struct SpkConv { const float* w = nullptr; const float* b = nullptr; int cout = 0, cin = 0, k = 1; };
struct SpkBlock { VConv tdnn1, tdnn2, se1, se2; std::vector<SpkConv> branches; int dil = 1; };
struct SpeakerW {
  SpkConv init;
  SpkBlock blk[3];
  VConv mfa, asp_tdnn, asp_conv, fc;
  int mel = 0, enc_dim = 0;
};

struct VocoderWorkspace {
  DBuf codes, e_first, e_rest, a, b, c, d, qh, kh, vh;
};

// Carried state of a STATEFUL streamed decode (q3_session_set_stream_context(sess, -1)): keys and values
// of the pre-transformer for every frame decoded so far, and the front half's output (the input of the causal conv stack),
// of which the back half needs the last 10 frames as left context (its look-back is 9.4 frames).
struct VocoderStreamState {
  DBuf kc, vc;        // [layers][B][heads][cap][head_dim] f32
  DBuf front;         // [B][latent][cap] f32, channel-major
  DBuf win;           // [B][latent][<= 10 + chunk] staging of the back half's input window
  int cap = 0, frames = 0;
};

// Vocoder scratch of one session (multi-GB at 256 frames x 8 rows).  Sessions borrow it from the model's pool and hand
// it back when they are destroyed: a cudaMalloc + cudaFree of these buffers per synthesize call cost ~0.7 s.
struct VocoderScratch {
  VocoderWorkspace ws;
  DBuf codes, pcm;
};

struct q3_model {
  q3_model_desc d;
  int num_sms = 148;
  bool finalized = false;
  std::map<std::string, RawTensor> t;
  std::vector<DBuf> owned;               // re-packed buffers
  // talker
  const bf16 *text_emb = nullptr, *codec_emb = nullptr, *fc1_w = nullptr, *fc1_b = nullptr, *fc2_w = nullptr,
             *fc2_b = nullptr, *t_norm = nullptr, *codec_head = nullptr;
  std::vector<LayerW> tl;
  // code predictor
  const bf16 *cp_proj_w = nullptr, *cp_proj_b = nullptr, *cp_norm = nullptr;
  const bf16* cp_emb[15] = {nullptr};
  const bf16* cp_head[15] = {nullptr};
  std::vector<LayerW> cl;
  const LayerW *tl_dev = nullptr, *cl_dev = nullptr;   // device copies of the layer tables (persistent kernel)
  const bf16 *cp_cos = nullptr, *cp_sin = nullptr;   // [cp_rope_positions][64]
  bool has_talker = false, has_vocoder = false, has_speaker = false;
  VocoderW voc;
  SpeakerW spk;
  mutable std::mutex voc_mutex;          // guards voc_ws for the session-less q3_vocoder_decode
  mutable VocoderWorkspace voc_ws;
  mutable std::mutex pool_mutex;
  mutable std::vector<std::unique_ptr<VocoderScratch>> voc_pool;   // idle scratch objects (at most 4 are kept)
  std::unique_ptr<VocoderScratch> acquire_scratch() const {
    std::lock_guard<std::mutex> lock(pool_mutex);
    if (!voc_pool.empty()) {
      std::unique_ptr<VocoderScratch> r = std::move(voc_pool.back());
      voc_pool.pop_back();
      return r;
    }
    return std::unique_ptr<VocoderScratch>(new VocoderScratch());
  }
  void release_scratch(std::unique_ptr<VocoderScratch> v) const {
    if (!v) return;
    std::lock_guard<std::mutex> lock(pool_mutex);
    if (voc_pool.size() < 4) voc_pool.push_back(std::move(v));
  }
  StackDims tdims() const { return {d.hidden, d.inter, d.heads, d.kv_heads, d.layers}; }
  StackDims cdims() const { return {d.cp_hidden, d.cp_inter, d.cp_heads, d.cp_kv_heads, d.cp_layers}; }
};

void vocoder_finalize(q3_model* m);
void speaker_finalize(q3_model* m);
// mel: device f32 [mel_dim][T] (one utterance); out: device f32 [enc_dim].  ref: SpeakerEncoder::forward (speaker.rs:448-476)
void speaker_run(const q3_model* m, const float* mel, int T, float* out, cudaStream_t st);
// codes: device i64 [B][nq][T]; pcm: device f32 [B][T*upsample]
void vocoder_run(const q3_model* m, VocoderWorkspace& ws, const long long* codes, int B, int T, float* pcm,
                 cudaStream_t st);
// Streamed chunk [f0, f0 + T) of every row, stateful: front half over the new frames with the carried keys / values (exact:
// the pre-transformer attends to the whole history), back half over the chunk and the min(10, f0) frames before it.
// codes_win: device i64 [B][nq][c0 + T] for frames [f0 - c0, f0 + T), c0 = min(2, f0) (left context of the k = 3 pre-conv).
// pcm: device f32 [B][(cb + T) * upsample] with cb = min(10, f0): the caller drops the first cb * upsample samples of a row.
void vocoder_stream_chunk(const q3_model* m, VocoderWorkspace& ws, VocoderStreamState& ss, const long long* codes_win, int B,
                          int f0, int T, float* pcm, cudaStream_t st);
constexpr int VOC_STREAM_BACK_CTX = 10;
constexpr int VOC_STREAM_FRONT_CTX = 2;
int vocoder_total_upsample(const q3_model* m);
// u32 frame-major codes [B][frames_cap][16] -> i64 [B][16][T] starting at frame f0 (codes_to_tensor, lib.rs:1417-1431)
void vocoder_codes_to_tensor(const uint32_t* frames, int frames_cap, int f0, int T, int B, long long* out, cudaStream_t st);
