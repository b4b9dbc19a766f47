// Vocoder (Decoder12Hz) host orchestration.  ref: Decoder12Hz::{from_weights, decode}
// (src/models/codec/decoder_12hz.rs:185-505).
#include <algorithm>
#include <cmath>

#include <cstdlib>

#include "model.h"
#include "vocoder_kernels.cuh"
#include "vocoder_mma.cuh"
#include "vocoder_umma.cuh"
#include "speaker.cuh"

namespace {

const RawTensor& need(const q3_model* m, const std::string& name) {
  auto it = m->t.find(name);
  if (it == m->t.end()) throw Q3Error(Q3_ERR_MISSING_WEIGHT, "Missing weight: " + name);
  if (it->second.dtype != Q3_F32) throw Q3Error(Q3_ERR_INVALID, "vocoder weight must be F32: " + name);
  return it->second;
}
const float* needp(const q3_model* m, const std::string& name) { return need(m, name).buf.as<float>(); }

// [A][Bd][k] -> [(a or b major)...]: conv weights [Cout][Cin][k] -> [Cin*k][Cout];
// transposed-conv weights [Cin][Cout][k] -> [Cin*k][Cout].
__global__ void repack_conv_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int k,
                                   int transposed) {
  size_t n = (size_t)Cout * Cin * k;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int co = (int)(i % Cout);
    size_t r = i / Cout;
    int j = (int)(r % k), ci = (int)(r / k);
    size_t src = transposed ? (((size_t)ci * Cout + co) * k + j) : (((size_t)co * Cin + ci) * k + j);
    out[i] = w[src];
  }
}

VConv make_conv(q3_model* m, const std::string& wname, const std::string& bname, bool transposed, int tconv_stride = 0) {
  const RawTensor& w = need(m, wname);
  VConv c;
  if (w.shape.size() == 2) {               // Linear [out][in] == 1x1 conv
    c.cout = (int)w.shape[0]; c.cin = (int)w.shape[1]; c.k = 1;
  } else {
    Q3_REQUIRE(w.shape.size() == 3, Q3_ERR_INVALID, "conv weight must be 3-D: " + wname);
    c.k = (int)w.shape[2];
    if (transposed) { c.cin = (int)w.shape[0]; c.cout = (int)w.shape[1]; }
    else { c.cout = (int)w.shape[0]; c.cin = (int)w.shape[1]; }
  }
  DBuf packed;
  packed.alloc(w.numel() * sizeof(float));
  repack_conv_kernel<<<256, 256>>>(w.buf.as<float>(), packed.as<float>(), c.cout, c.cin, c.k, transposed ? 1 : 0);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
  c.w = packed.as<float>();
  m->owned.push_back(std::move(packed));
  if (!bname.empty()) c.b = needp(m, bname);
  // tensor-core layout (bf16 hi/lo split)
  c.cout_pad = ceil_div(c.cout, MC_BM) * MC_BM;
  c.chunks = ceil_div(c.cin, MC_BK);
  const size_t np = (size_t)c.k * c.chunks * c.cout_pad * MC_BK;
  DBuf hi, lo;
  hi.alloc(np * sizeof(bf16));
  lo.alloc(np * sizeof(bf16));
  voc_pack_mma_weights_kernel<<<512, 256>>>(w.buf.as<float>(), hi.as<bf16>(), lo.as<bf16>(), c.cout, c.cin, c.k, c.cout_pad, c.chunks,
                                            transposed ? 1 : 0);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
  c.w_hi = hi.as<bf16>();
  c.w_lo = lo.as<bf16>();
  m->owned.push_back(std::move(hi));
  m->owned.push_back(std::move(lo));
  if (c.cin % MC_BK == 0) {
    // tcgen05 layout: the shared-memory image of every (tap, chunk, 128-row tile).  A transposed conv of stride s is packed
    // as ONE GEMM whose rows are (output channel, phase r < s) pairs, row = co * s + r, with k / s taps (launch_tconv)
    const int s = transposed ? tconv_stride : 1;
    const bool tc = transposed && s > 0 && (c.k == s || c.k == 2 * s);
    c.um_rows_pad = tc ? ceil_div(c.cout * s, MC_BM) * MC_BM : c.cout_pad;
    const int ntp = tc ? c.k / s : c.k;
    const size_t nu = (size_t)ntp * c.chunks * (c.um_rows_pad / MC_BM) * UC_A_ELEMS;
    DBuf um;
    um.alloc(nu * sizeof(bf16));
    voc_pack_umma_weights_kernel<<<512, 256>>>(w.buf.as<float>(), um.as<bf16>(), c.cout, c.cin, c.k, c.um_rows_pad, c.chunks,
                                               transposed ? 1 : 0, tc ? s : 1, ntp);
    Q3_COUNT_LAUNCH();
    Q3_LAUNCH_CHECK();
    c.w_um = um.as<bf16>();
    m->owned.push_back(std::move(um));
  }
  return c;
}

VSnake make_snake(q3_model* m, const std::string& prefix) {
  const RawTensor& a = need(m, prefix + ".alpha");
  const RawTensor& b = need(m, prefix + ".beta");
  int n = (int)a.numel();
  DBuf ea, ib;
  ea.alloc(n * sizeof(float));
  ib.alloc(n * sizeof(float));
  voc_prep_snake_kernel<<<ceil_div(n, 256), 256>>>(a.buf.as<float>(), b.buf.as<float>(), ea.as<float>(), ib.as<float>(), n);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
  VSnake s{ea.as<float>(), ib.as<float>()};
  m->owned.push_back(std::move(ea));
  m->owned.push_back(std::move(ib));
  return s;
}

bool use_mma_path() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("Q3_VOC_SIMT");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

// Which convolutions run on the tcgen05 kernel: Q3_VOC_UMMA = bit mask (1: 1x1, 2: k > 1 undilated, 4: dilated, 8: transposed
// phases); default all, 0 = the mma.sync kernel everywhere (A/B runs, tools/voc_ab.py).
int umma_mask() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("Q3_VOC_UMMA");
    v = e ? std::atoi(e) : 15;
  }
  return v;
}

void launch_mma(MmaConvArgs& a, int grid_q, int Cout_pad, int z, cudaStream_t st, int kind_bit) {
  const int halo = a.max_shift - a.min_shift;
  if ((umma_mask() & kind_bit) && a.w_um != nullptr && a.Cin % MC_BK == 0 && MC_BN + halo <= UC_MAX_BROWS) {
    const size_t smem_u = umma_conv_smem_bytes(halo);
    {
      static std::mutex mu;
      static bool configured[64] = {};
      int dev = 0;
      Q3_CHECK_CUDA(cudaGetDevice(&dev));
      std::lock_guard<std::mutex> lock(mu);
      if (dev < 0 || dev >= 64 || !configured[dev]) {
        Q3_CHECK_CUDA(cudaFuncSetAttribute(voc_conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)umma_conv_smem_bytes(UC_MAX_BROWS - MC_BN)));
        if (dev >= 0 && dev < 64) configured[dev] = true;
      }
    }
    voc_conv_umma_kernel<<<dim3(grid_q, Cout_pad / MC_BM, z), UC_THREADS, smem_u, st>>>(a);
    Q3_COUNT_LAUNCH();
    Q3_LAUNCH_CHECK();
    return;
  }
  const size_t smem = mma_conv_smem_bytes(a.max_shift - a.min_shift);
  {
    // per-DEVICE function attribute (see gemm_tc_launch)
    static std::mutex mu;
    static bool configured[64] = {};
    int dev = 0;
    Q3_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      Q3_CHECK_CUDA(cudaFuncSetAttribute(voc_conv_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
  }
  Q3_REQUIRE(smem <= 100 * 1024, Q3_ERR_UNSUPPORTED, "conv window too large for the staged tile");
  Q3_REQUIRE((MC_BK / 2) * (MC_BN + a.max_shift - a.min_shift) <= 12 * 256, Q3_ERR_UNSUPPORTED,
             "conv window too large for the register-prefetched staging");
  voc_conv_mma_kernel<<<dim3(grid_q, Cout_pad / MC_BM, z), 256, smem, st>>>(a);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
}

void launch_conv(const VConv& c, const float* x, float* y, int B, int T, int dil, const VSnake* snake, const float* res,
                 const float* scale, int epi, cudaStream_t st) {
  if (use_mma_path() && c.cout >= 16 && c.k <= MC_MAX_TAPS) {
    MmaConvArgs m{};
    m.x = x; m.w_hi = c.w_hi; m.w_lo = c.w_lo; m.w_um = c.w_um; m.bias = c.b;
    m.snake_a = snake ? snake->ea : nullptr; m.snake_ib = snake ? snake->ib : nullptr;
    m.res = res; m.scale = scale; m.y = y;
    m.B = B; m.Cin = c.cin; m.Cout = c.cout; m.Cout_pad = c.cout_pad; m.Tin = T; m.Tout = T; m.Q = T;
    m.ntaps = c.k;
    for (int j = 0; j < c.k; ++j) { m.tap_w[j] = j; m.tap_shift[j] = -(c.k - 1 - j) * dil; }
    m.min_shift = -(c.k - 1) * dil; m.max_shift = 0;
    m.out_stride = 1; m.out_off = 0; m.epi = epi; m.phases = 1; m.phase_tap_step = 0;
    launch_mma(m, ceil_div(T, MC_BN), c.cout_pad, B, st, c.k == 1 ? 1 : (dil == 1 ? 2 : 4));
    return;
  }
  if (c.cout == 1 && dil == 1 && res == nullptr && scale == nullptr && (epi == CEPI_NONE || epi == CEPI_CLAMP)) {
    // SIMT layout [Cin*k][Cout] with Cout == 1 is exactly [Cin][k]
    const size_t smem = (size_t)(32 * (256 + c.k - 1) + 32 * c.k) * sizeof(float);
    voc_conv_cout1_kernel<<<dim3(ceil_div(T, 256), B), 256, smem, st>>>(x, c.w, c.b, snake ? snake->ea : nullptr,
                                                                        snake ? snake->ib : nullptr, y, c.cin, T, c.k,
                                                                        epi == CEPI_CLAMP ? 1 : 0);
    Q3_COUNT_LAUNCH();
    Q3_LAUNCH_CHECK();
    return;
  }
  ConvArgs a;
  a.x = x; a.w = c.w; a.bias = c.b;
  a.snake_a = snake ? snake->ea : nullptr; a.snake_ib = snake ? snake->ib : nullptr;
  a.res = res; a.scale = scale; a.y = y;
  a.B = B; a.Cin = c.cin; a.Cout = c.cout; a.T = T; a.k = c.k; a.dil = dil; a.epi = epi;
  dim3 grid(ceil_div(T, CV_BN), ceil_div(c.cout, CV_BM), B);
  size_t smem = conv_smem_bytes(c.k, dil);
  Q3_REQUIRE(smem <= 48 * 1024, Q3_ERR_UNSUPPORTED, "conv kernel/dilation too large for the staged tile");
  voc_conv1d_kernel<<<grid, 256, smem, st>>>(a);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
}

void launch_tconv(const VConv& c, int stride, const float* x, float* y, int B, int T, const VSnake* snake, cudaStream_t st) {
  Q3_REQUIRE(c.k <= 2 * stride && c.k >= stride, Q3_ERR_UNSUPPORTED, "transposed conv needs stride <= k <= 2*stride");
  if (use_mma_path() && c.cout >= 16 && (c.k == stride || c.k == 2 * stride) && (umma_mask() & 8) && c.w_um != nullptr &&
      c.um_rows_pad == ceil_div(c.cout * stride, MC_BM) * MC_BM) {
    // tcgen05 kernel: ONE GEMM whose rows are (output channel, phase) pairs (voc_pack_umma_weights_kernel), so a CTA owns
    // all `stride` phases of its channels and writes whole runs of consecutive samples (the phase-per-CTA form below writes
    // every stride-th float of a sector from a different CTA: 6.0 ms for the last block's 192 -> 96 x3 upsample, ncu)
    MmaConvArgs m{};
    m.x = x; m.w_um = c.w_um; m.bias = c.b;
    m.snake_a = snake ? snake->ea : nullptr; m.snake_ib = snake ? snake->ib : nullptr;
    m.y = y; m.B = B; m.Cin = c.cin; m.Cout = c.cout; m.Cout_pad = c.um_rows_pad; m.Tin = T; m.Tout = T * stride; m.Q = T;
    m.ntaps = c.k / stride;
    for (int j = 0; j < m.ntaps; ++j) { m.tap_w[j] = j; m.tap_shift[j] = j - (m.ntaps - 1); }
    m.min_shift = -(m.ntaps - 1); m.max_shift = 0;
    m.out_stride = stride; m.out_off = 0; m.epi = CEPI_NONE; m.phases = 1; m.phase_tap_step = 0; m.rdiv = stride;
    launch_mma(m, ceil_div(T, MC_BN), c.um_rows_pad, B, st, 8);
    return;
  }
  if (use_mma_path() && c.cout >= 16 && (c.k == stride || c.k == 2 * stride)) {
    // phase r in [0, stride): y[co][stride*q + r] = b + sum_ci x[ci][q] w[ci][co][r] (+ x[ci][q-1] w[ci][co][r+stride])
    MmaConvArgs m{};
    m.x = x; m.w_hi = c.w_hi; m.w_lo = c.w_lo; m.w_um = c.w_um; m.bias = c.b;
    m.snake_a = snake ? snake->ea : nullptr; m.snake_ib = snake ? snake->ib : nullptr;
    m.y = y; m.B = B; m.Cin = c.cin; m.Cout = c.cout; m.Cout_pad = c.cout_pad; m.Tin = T; m.Tout = T * stride; m.Q = T;
    if (c.k == 2 * stride) {
      m.ntaps = 2;
      m.tap_w[0] = stride; m.tap_shift[0] = -1;
      m.tap_w[1] = 0; m.tap_shift[1] = 0;
      m.min_shift = -1;
    } else {
      m.ntaps = 1;
      m.tap_w[0] = 0; m.tap_shift[0] = 0;
      m.min_shift = 0;
    }
    m.max_shift = 0; m.out_stride = stride; m.out_off = 0; m.epi = CEPI_NONE; m.phases = stride; m.phase_tap_step = 1;
    launch_mma(m, ceil_div(T, MC_BN), c.cout_pad, B * stride, st, 0);
    return;
  }
  TConvArgs a;
  a.x = x; a.w = c.w; a.bias = c.b;
  a.snake_a = snake ? snake->ea : nullptr; a.snake_ib = snake ? snake->ib : nullptr;
  a.y = y; a.B = B; a.Cin = c.cin; a.Cout = c.cout; a.T = T; a.k = c.k; a.stride = stride;
  dim3 grid(ceil_div(T * stride, TC_BN), ceil_div(c.cout, TC_BM), B);
  size_t smem = tconv_smem_bytes(c.k);
  Q3_REQUIRE(smem <= 48 * 1024, Q3_ERR_UNSUPPORTED, "transposed conv kernel too large for the staged tile");
  voc_tconv1d_kernel<<<grid, 256, smem, st>>>(a);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
}

void launch_norm(const float* x, const float* w, const float* b, float* y, int B, int C, int T, float eps, int mode,
                 cudaStream_t st) {
  dim3 grid(ceil_div(T, 32), B), block(32, 8);
  voc_channel_norm_kernel<<<grid, block, 0, st>>>(x, w, b, y, C, T, eps, mode);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
}

}  // namespace

void vocoder_codes_to_tensor(const uint32_t* frames, int frames_cap, int f0, int T, int B, long long* out, cudaStream_t st) {
  voc_codes_to_tensor_kernel<<<dim3(ceil_div(16 * T, 256), B), 256, 0, st>>>(frames, frames_cap, f0, T, out);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
}

int vocoder_total_upsample(const q3_model* m) {
  int n = 1;
  for (int i = 0; i < m->d.v_n_upsampling; ++i) n *= m->d.v_upsampling[i];
  for (int i = 0; i < m->d.v_n_rates; ++i) n *= m->d.v_rates[i];
  return n;
}

void vocoder_finalize(q3_model* m) {
  const q3_model_desc& d = m->d;
  VocoderW& v = m->voc;
  const std::string q = "decoder.quantizer";
  // codebooks: embedding_sum / clamp(cluster_usage, 1e-7)   (decoder_12hz.rs:199-225)
  {
    const int rows = d.v_codebook_size, dim = d.v_vq_dim;
    DBuf first, rest;
    first.alloc((size_t)rows * dim * sizeof(float));
    rest.alloc((size_t)(d.v_quantizers - 1) * rows * dim * sizeof(float));
    auto prep = [&](const std::string& pre, float* out) {
      const RawTensor& es = need(m, pre + "._codebook.embedding_sum");
      const RawTensor& cu = need(m, pre + "._codebook.cluster_usage");
      Q3_REQUIRE((int)es.numel() == rows * dim && (int)cu.numel() == rows, Q3_ERR_INVALID, "codebook shape: " + pre);
      voc_prep_codebook_kernel<<<ceil_div(rows * dim, 256), 256>>>(es.buf.as<float>(), cu.buf.as<float>(), out, rows, dim);
      Q3_COUNT_LAUNCH();
      Q3_LAUNCH_CHECK();
    };
    prep(q + ".rvq_first.vq.layers.0", first.as<float>());
    for (int i = 0; i < d.v_quantizers - 1; ++i)
      prep(q + ".rvq_rest.vq.layers." + std::to_string(i), rest.as<float>() + (size_t)i * rows * dim);
    v.first_cb = first.as<float>();
    v.rest_cb = rest.as<float>();
    m->owned.push_back(std::move(first));
    m->owned.push_back(std::move(rest));
  }
  v.first_proj = make_conv(m, q + ".rvq_first.output_proj.weight", "", false);
  v.rest_proj = make_conv(m, q + ".rvq_rest.output_proj.weight", "", false);
  v.pre_conv = make_conv(m, "decoder.pre_conv.conv.weight", "decoder.pre_conv.conv.bias", false);
  const std::string t = "decoder.pre_transformer";
  v.in_proj = make_conv(m, t + ".input_proj.weight", t + ".input_proj.bias", false);
  v.out_proj = make_conv(m, t + ".output_proj.weight", t + ".output_proj.bias", false);
  v.layers.clear();
  for (int l = 0; l < d.v_layers; ++l) {
    const std::string p = t + ".layers." + std::to_string(l);
    VLayer L;
    L.in_ln = needp(m, p + ".input_layernorm.weight");
    L.post_ln = needp(m, p + ".post_attention_layernorm.weight");
    L.attn_scale = needp(m, p + ".self_attn_layer_scale.scale");
    L.mlp_scale = needp(m, p + ".mlp_layer_scale.scale");
    L.q = make_conv(m, p + ".self_attn.q_proj.weight", "", false);
    L.k = make_conv(m, p + ".self_attn.k_proj.weight", "", false);
    L.v = make_conv(m, p + ".self_attn.v_proj.weight", "", false);
    L.o = make_conv(m, p + ".self_attn.o_proj.weight", "", false);
    L.gate = make_conv(m, p + ".mlp.gate_proj.weight", "", false);
    L.up = make_conv(m, p + ".mlp.up_proj.weight", "", false);
    L.down = make_conv(m, p + ".mlp.down_proj.weight", "", false);
    v.layers.push_back(L);
  }
  v.final_norm = needp(m, t + ".norm.weight");
  v.ups.clear();
  for (int s = 0; s < d.v_n_upsampling; ++s) {
    const std::string p = "decoder.upsample." + std::to_string(s);
    VUpsample u;
    u.ratio = d.v_upsampling[s];
    u.tconv = make_conv(m, p + ".0.conv.weight", p + ".0.conv.bias", true, u.ratio);
    u.cn.dw_w = needp(m, p + ".1.dwconv.conv.weight");
    u.cn.dw_b = needp(m, p + ".1.dwconv.conv.bias");
    u.cn.ln_w = needp(m, p + ".1.norm.weight");
    u.cn.ln_b = needp(m, p + ".1.norm.bias");
    u.cn.gamma = needp(m, p + ".1.gamma");
    u.cn.pw1 = make_conv(m, p + ".1.pwconv1.weight", p + ".1.pwconv1.bias", false);
    u.cn.pw2 = make_conv(m, p + ".1.pwconv2.weight", p + ".1.pwconv2.bias", false);
    u.cn.C = u.tconv.cout;
    v.ups.push_back(u);
  }
  v.init_conv = make_conv(m, "decoder.decoder.0.conv.weight", "decoder.decoder.0.conv.bias", false);
  v.blocks.clear();
  for (int b = 0; b < d.v_n_rates; ++b) {
    const std::string bp = "decoder.decoder." + std::to_string(b + 1) + ".block";
    VBlock blk;
    blk.rate = d.v_rates[b];
    blk.s = make_snake(m, bp + ".0");
    blk.up = make_conv(m, bp + ".1.conv.weight", bp + ".1.conv.bias", true, blk.rate);
    const int dils[3] = {1, 3, 9};
    for (int u = 0; u < 3; ++u) {
      const std::string up = bp + "." + std::to_string(u + 2);
      blk.ru[u].a1 = make_snake(m, up + ".act1");
      blk.ru[u].c1 = make_conv(m, up + ".conv1.conv.weight", up + ".conv1.conv.bias", false);
      blk.ru[u].a2 = make_snake(m, up + ".act2");
      blk.ru[u].c2 = make_conv(m, up + ".conv2.conv.weight", up + ".conv2.conv.bias", false);
      blk.ru[u].dil = dils[u];
    }
    v.blocks.push_back(blk);
  }
  const std::string fs = "decoder.decoder." + std::to_string(d.v_n_rates + 1);
  const std::string fc = "decoder.decoder." + std::to_string(d.v_n_rates + 2);
  v.final_snake = make_snake(m, fs);
  v.final_conv = make_conv(m, fc + ".conv.weight", fc + ".conv.bias", false);
  Q3_CHECK_CUDA(cudaDeviceSynchronize());
  m->has_vocoder = true;
}

// Largest [C][T'] activation (floats per batch row) for T frames.
static size_t vocoder_max_act(const q3_model* m, int T) {
  const q3_model_desc& d = m->d;
  size_t mx = (size_t)std::max(std::max(std::max(d.v_latent_dim, d.v_codebook_dim), d.v_heads * d.v_head_dim), std::max(d.v_inter, d.v_hidden)) * T;
  int len = T;
  int c = d.v_latent_dim;
  for (int s = 0; s < d.v_n_upsampling; ++s) {
    len *= d.v_upsampling[s];
    mx = std::max(mx, (size_t)4 * c * len);       // ConvNeXt expansion
  }
  c = d.v_decoder_dim;
  mx = std::max(mx, (size_t)c * len);
  for (int b = 0; b < d.v_n_rates; ++b) {
    len *= d.v_rates[b];
    c /= 2;
    mx = std::max(mx, (size_t)c * len);
  }
  return mx;
}

namespace {

struct VocBufs { float *A, *Bf, *C, *D; };

VocBufs voc_ensure(const q3_model* m, VocoderWorkspace& ws, int B, int T) {
  const q3_model_desc& d = m->d;
  const size_t act = vocoder_max_act(m, T) * (size_t)B * sizeof(float);
  ws.a.ensure(act); ws.b.ensure(act); ws.c.ensure(act); ws.d.ensure(act);
  const int AD = d.v_heads * d.v_head_dim;
  ws.e_first.ensure((size_t)B * d.v_vq_dim * T * sizeof(float));
  ws.e_rest.ensure((size_t)B * d.v_vq_dim * T * sizeof(float));
  ws.qh.ensure((size_t)B * AD * T * sizeof(float));
  ws.kh.ensure((size_t)B * AD * T * sizeof(float));
  ws.vh.ensure((size_t)B * AD * T * sizeof(float));
  return {ws.a.as<float>(), ws.b.as<float>(), ws.c.as<float>(), ws.d.as<float>()};
}

// RVQ decode + pre_conv + input projection over T frames: codes -> out [B][hidden][T]   (decoder_12hz.rs:420-470)
void voc_embed(const q3_model* m, VocoderWorkspace& ws, const VocBufs& w, const long long* codes, int B, int T, float* out,
               cudaStream_t st) {
  const q3_model_desc& d = m->d;
  const VocoderW& v = m->voc;
  // 1. RVQ decode: first_proj(E_first) + rest_proj(sum E_rest)
  voc_rvq_gather_kernel<<<dim3(T, B), 128, 0, st>>>(codes, v.first_cb, v.rest_cb, d.v_quantizers, d.v_codebook_size,
                                                   d.v_vq_dim, T, ws.e_first.as<float>(), ws.e_rest.as<float>());
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
  launch_conv(v.first_proj, ws.e_first.as<float>(), w.A, B, T, 1, nullptr, nullptr, nullptr, CEPI_NONE, st);
  launch_conv(v.rest_proj, ws.e_rest.as<float>(), w.Bf, B, T, 1, nullptr, w.A, nullptr, CEPI_NONE, st);   // Bf = A + rest
  // 2. pre_conv (k3) -> C ; 3. input_proj -> out (hidden 512)
  launch_conv(v.pre_conv, w.Bf, w.C, B, T, 1, nullptr, nullptr, nullptr, CEPI_NONE, st);
  launch_conv(v.in_proj, w.C, out, B, T, 1, nullptr, nullptr, nullptr, CEPI_NONE, st);
}

// Pre-transformer (decoder_12hz.rs:586-672) over T frames whose first sits at position pos0; hidden lives in A and the result
// of the output projection ([B][latent][T]) ends up in A.  kc / vc == nullptr: keys / values of this call only (whole
// utterance, pos0 == 0); else the session's caches [layers][B][heads][cap][head_dim], to which this call appends.
void voc_transformer(const q3_model* m, VocoderWorkspace& ws, const VocBufs& w, int B, int T, int pos0, float* kc, float* vc,
                     int cap, cudaStream_t st) {
  const q3_model_desc& d = m->d;
  const VocoderW& v = m->voc;
  float *A = w.A, *Bf = w.Bf, *C = w.C, *D = w.D;
  const float scale = 1.0f / sqrtf((float)d.v_head_dim);
  const int Ltot = pos0 + T;
  Q3_REQUIRE(voc_attn_smem_floats(Ltot, d.v_head_dim) * sizeof(float) <= 48 * 1024, Q3_ERR_UNSUPPORTED,
             "vocoder attention: more than ~2400 frames in one utterance");
  const size_t layer_stride = (size_t)B * d.v_heads * cap * d.v_head_dim;
  int li = 0;
  for (const VLayer& L : v.layers) {
    float* kdst = kc ? kc + (size_t)li * layer_stride : ws.kh.as<float>();
    float* vdst = vc ? vc + (size_t)li * layer_stride : ws.vh.as<float>();
    const int Tk = kc ? cap : T, to0 = kc ? pos0 : 0;
    ++li;
    launch_norm(A, L.in_ln, nullptr, Bf, B, d.v_hidden, T, d.v_rms_eps, 0, st);
    launch_conv(L.q, Bf, C, B, T, 1, nullptr, nullptr, nullptr, CEPI_NONE, st);
    voc_rope_relayout_kernel<<<dim3(T, d.v_heads, B), 64, 0, st>>>(C, ws.qh.as<float>(), d.v_heads, d.v_head_dim, T, d.v_rope_theta, 1, pos0, T, 0);
    Q3_COUNT_LAUNCH();
    launch_conv(L.k, Bf, C, B, T, 1, nullptr, nullptr, nullptr, CEPI_NONE, st);
    voc_rope_relayout_kernel<<<dim3(T, d.v_heads, B), 64, 0, st>>>(C, kdst, d.v_heads, d.v_head_dim, T, d.v_rope_theta, 1, pos0, Tk, to0);
    Q3_COUNT_LAUNCH();
    launch_conv(L.v, Bf, C, B, T, 1, nullptr, nullptr, nullptr, CEPI_NONE, st);
    voc_rope_relayout_kernel<<<dim3(T, d.v_heads, B), 64, 0, st>>>(C, vdst, d.v_heads, d.v_head_dim, T, d.v_rope_theta, 0, pos0, Tk, to0);
    Q3_COUNT_LAUNCH();
    voc_attn_kernel<<<dim3(ceil_div(T, 4), d.v_heads, B), 128, voc_attn_smem_floats(Ltot, d.v_head_dim) * sizeof(float), st>>>(
        ws.qh.as<float>(), kdst, vdst, C, d.v_heads, d.v_head_dim, T, scale, Tk, pos0);
    Q3_COUNT_LAUNCH();
    Q3_LAUNCH_CHECK();
    launch_conv(L.o, C, D, B, T, 1, nullptr, A, L.attn_scale, CEPI_NONE, st);        // D = A + scale*o_proj
    launch_norm(D, L.post_ln, nullptr, Bf, B, d.v_hidden, T, d.v_rms_eps, 0, st);
    launch_conv(L.gate, Bf, C, B, T, 1, nullptr, nullptr, nullptr, CEPI_NONE, st);
    launch_conv(L.up, Bf, A, B, T, 1, nullptr, nullptr, nullptr, CEPI_NONE, st);
    {
      size_t n = (size_t)B * d.v_inter * T;
      voc_silu_mul_kernel<<<(int)std::min<size_t>(4096, (n + 255) / 256), 256, 0, st>>>(C, A, C, n);
      Q3_COUNT_LAUNCH();
      Q3_LAUNCH_CHECK();
    }
    launch_conv(L.down, C, A, B, T, 1, nullptr, D, L.mlp_scale, CEPI_NONE, st);      // A = D + scale*down
  }
  launch_norm(A, v.final_norm, nullptr, Bf, B, d.v_hidden, T, d.v_rms_eps, 0, st);
  launch_conv(v.out_proj, Bf, A, B, T, 1, nullptr, nullptr, nullptr, CEPI_NONE, st);   // [B][latent][T]
}

// Back half: upsample stages (transposed conv + ConvNeXt), decoder blocks, final conv; input w.A [B][latent][T] (consumed),
// output pcm [B][T * upsample].  Every op is a causal convolution: look-back 9.4 frames in total.
void voc_back(const q3_model* m, const VocBufs& w, int B, int T, float* pcm, cudaStream_t st) {
  const VocoderW& v = m->voc;
  int len = T;
  float* cur = w.A;
  float* o1 = w.Bf;
  float* o2 = w.C;
  float* o3 = w.D;
  for (const VUpsample& u : v.ups) {
    launch_tconv(u.tconv, u.ratio, cur, o1, B, len, nullptr, st);
    len *= u.ratio;
    const int Cn = u.cn.C;
    voc_dwconv_kernel<<<dim3(ceil_div(len, 256), B * Cn), 256, 0, st>>>(o1, u.cn.dw_w, u.cn.dw_b, o2, Cn, len, 7);
    Q3_COUNT_LAUNCH();
    Q3_LAUNCH_CHECK();
    launch_norm(o2, u.cn.ln_w, u.cn.ln_b, o3, B, Cn, len, 1e-6f, 1, st);
    launch_conv(u.cn.pw1, o3, o2, B, len, 1, nullptr, nullptr, nullptr, CEPI_GELU, st);
    launch_conv(u.cn.pw2, o2, cur, B, len, 1, nullptr, o1, u.cn.gamma, CEPI_NONE, st);   // cur = o1 + gamma*pw2
  }
  launch_conv(v.init_conv, cur, o1, B, len, 1, nullptr, nullptr, nullptr, CEPI_NONE, st);
  std::swap(cur, o1);
  for (const VBlock& blk : v.blocks) {
    launch_tconv(blk.up, blk.rate, cur, o1, B, len, &blk.s, st);
    len *= blk.rate;
    std::swap(cur, o1);
    for (int r = 0; r < 3; ++r) {
      const VResUnit& ru = blk.ru[r];
      launch_conv(ru.c1, cur, o1, B, len, ru.dil, &ru.a1, nullptr, nullptr, CEPI_NONE, st);
      launch_conv(ru.c2, o1, o2, B, len, 1, &ru.a2, cur, nullptr, CEPI_NONE, st);        // o2 = cur + conv2
      std::swap(cur, o2);
    }
  }
  launch_conv(v.final_conv, cur, pcm, B, len, 1, &v.final_snake, nullptr, nullptr, CEPI_CLAMP, st);
}

}  // namespace

void vocoder_run(const q3_model* m, VocoderWorkspace& ws, const long long* codes, int B, int T, float* pcm, cudaStream_t st) {
  Q3_REQUIRE(m->has_vocoder, Q3_ERR_STATE, "vocoder weights were not loaded");
  if (B <= 0 || T <= 0) return;
  Q3_REQUIRE(T <= 3072, Q3_ERR_UNSUPPORTED, "vocoder: at most 3072 frames per call");
  const VocBufs w = voc_ensure(m, ws, B, T);
  voc_embed(m, ws, w, codes, B, T, w.A, st);
  voc_transformer(m, ws, w, B, T, 0, nullptr, nullptr, 0, st);
  voc_back(m, w, B, T, pcm, st);
}

void vocoder_stream_chunk(const q3_model* m, VocoderWorkspace& ws, VocoderStreamState& ss, const long long* codes_win, int B,
                          int f0, int T, float* pcm, cudaStream_t st) {
  Q3_REQUIRE(m->has_vocoder, Q3_ERR_STATE, "vocoder weights were not loaded");
  if (B <= 0 || T <= 0) return;
  const q3_model_desc& d = m->d;
  Q3_REQUIRE(ss.cap > 0 && f0 == ss.frames && f0 + T <= ss.cap, Q3_ERR_STATE, "vocoder stream state out of step");
  Q3_REQUIRE(f0 + T <= 3072, Q3_ERR_UNSUPPORTED, "vocoder: at most 3072 frames per utterance");
  const int c0 = std::min(VOC_STREAM_FRONT_CTX, f0), Tw = c0 + T;
  const int cb = std::min(VOC_STREAM_BACK_CTX, f0), Tb = cb + T;
  const VocBufs w = voc_ensure(m, ws, B, std::max(Tw, Tb));
  const int Hd = d.v_hidden, Lt = d.v_latent_dim;
  // ---- front half over the new frames ----
  // embed [f0 - c0, f0 + T): the k = 3 pre-conv sees its true left context, the c0 leading outputs are dropped
  voc_embed(m, ws, w, codes_win, B, Tw, w.D, st);
  Q3_CHECK_CUDA(cudaMemcpy2DAsync(w.A, (size_t)T * sizeof(float), w.D + c0, (size_t)Tw * sizeof(float), (size_t)T * sizeof(float),
                                  (size_t)B * Hd, cudaMemcpyDeviceToDevice, st));
  voc_transformer(m, ws, w, B, T, f0, ss.kc.as<float>(), ss.vc.as<float>(), ss.cap, st);
  // keep the front half's output of every frame: the back half of the NEXT chunks needs the last 10 as left context
  Q3_CHECK_CUDA(cudaMemcpy2DAsync(ss.front.as<float>() + f0, (size_t)ss.cap * sizeof(float), w.A, (size_t)T * sizeof(float),
                                  (size_t)T * sizeof(float), (size_t)B * Lt, cudaMemcpyDeviceToDevice, st));
  // ---- back half over [f0 - cb, f0 + T) ----
  Q3_CHECK_CUDA(cudaMemcpy2DAsync(w.A, (size_t)Tb * sizeof(float), ss.front.as<float>() + (f0 - cb), (size_t)ss.cap * sizeof(float),
                                  (size_t)Tb * sizeof(float), (size_t)B * Lt, cudaMemcpyDeviceToDevice, st));
  voc_back(m, w, B, Tb, pcm, st);
  ss.frames = f0 + T;
}

// ---- ECAPA-TDNN speaker encoder (voice-clone front end) ------------------------------------------------
// ref: SpeakerEncoder::{new, forward} (src/models/speaker.rs:362-476).  Dilations are the reference's config defaults
// (config.rs:144-146: 1, 2, 3, 4, 1 -- the published checkpoints use them; q3_model_desc carries no speaker section), every
// other dimension is read from the weight shapes.
namespace {
SpkConv make_spk_conv(q3_model* m, const std::string& name) {
  const RawTensor& w = need(m, name + ".weight");
  Q3_REQUIRE(w.shape.size() == 3, Q3_ERR_INVALID, "conv weight must be 3-D: " + name);
  SpkConv c;
  c.w = w.buf.as<float>();
  c.b = needp(m, name + ".bias");
  c.cout = (int)w.shape[0]; c.cin = (int)w.shape[1]; c.k = (int)w.shape[2];
  return c;
}
void launch_reflect_conv(const SpkConv& c, const float* xa, const float* xb, float* y, int dil, int T, cudaStream_t st) {
  Q3_REQUIRE(dil * (c.k - 1) / 2 < T && dil * (c.k - 1) - dil * (c.k - 1) / 2 < T, Q3_ERR_INVALID,
             "speaker encoder: the mel spectrogram is shorter than the reflect padding");
  const size_t smem = (size_t)c.cin * c.k * sizeof(float);
  Q3_REQUIRE(smem <= 48 * 1024, Q3_ERR_UNSUPPORTED, "speaker encoder: conv too wide for the staged weights");
  spk_reflect_conv_relu_kernel<<<dim3(ceil_div(T, 128), c.cout), 128, smem, st>>>(xa, xb, c.w, c.b, y, c.cin, c.k, dil, T);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
}
}  // namespace

void speaker_finalize(q3_model* m) {
  const std::string p = "speaker_encoder";
  SpeakerW& s = m->spk;
  const int dils[5] = {1, 2, 3, 4, 1};
  s.init = make_spk_conv(m, p + ".blocks.0.conv");
  s.mel = s.init.cin;
  for (int i = 0; i < 3; ++i) {
    const std::string b = p + ".blocks." + std::to_string(i + 1);
    SpkBlock& k = s.blk[i];
    k.dil = dils[i + 1];
    k.tdnn1 = make_conv(m, b + ".tdnn1.conv.weight", b + ".tdnn1.conv.bias", false);
    k.tdnn2 = make_conv(m, b + ".tdnn2.conv.weight", b + ".tdnn2.conv.bias", false);
    k.se1 = make_conv(m, b + ".se_block.conv1.weight", b + ".se_block.conv1.bias", false);
    k.se2 = make_conv(m, b + ".se_block.conv2.weight", b + ".se_block.conv2.bias", false);
    Q3_REQUIRE(k.tdnn1.k == 1 && k.tdnn2.k == 1 && k.se1.k == 1 && k.se2.k == 1, Q3_ERR_UNSUPPORTED, "speaker encoder: 1x1 convs expected in " + b);
    k.branches.clear();
    for (int j = 0;; ++j) {
      const std::string n = b + ".res2net_block.blocks." + std::to_string(j) + ".conv";
      if (!m->t.count(n + ".weight")) break;
      k.branches.push_back(make_spk_conv(m, n));
    }
    Q3_REQUIRE(!k.branches.empty() && k.branches[0].cout * ((int)k.branches.size() + 1) == k.tdnn1.cout, Q3_ERR_INVALID,
               "speaker encoder: Res2Net branches do not tile the block's channels: " + b);
  }
  s.mfa = make_conv(m, p + ".mfa.conv.weight", p + ".mfa.conv.bias", false);
  s.asp_tdnn = make_conv(m, p + ".asp.tdnn.conv.weight", p + ".asp.tdnn.conv.bias", false);
  s.asp_conv = make_conv(m, p + ".asp.conv.weight", p + ".asp.conv.bias", false);
  s.fc = make_conv(m, p + ".fc.weight", p + ".fc.bias", false);
  Q3_REQUIRE(s.mfa.k == 1 && s.asp_tdnn.k == 1 && s.asp_conv.k == 1 && s.fc.k == 1, Q3_ERR_UNSUPPORTED, "speaker encoder: 1x1 convs expected (mfa / asp / fc)");
  Q3_REQUIRE(s.mfa.cin == s.blk[0].tdnn1.cout + s.blk[1].tdnn1.cout + s.blk[2].tdnn1.cout && s.asp_tdnn.cin == 3 * s.mfa.cout &&
                 s.asp_conv.cout == s.mfa.cout && s.fc.cin == 2 * s.mfa.cout, Q3_ERR_INVALID, "speaker encoder: inconsistent shapes");
  s.enc_dim = s.fc.cout;
  m->has_speaker = true;
}

void speaker_run(const q3_model* m, const float* mel, int T, float* out, cudaStream_t st) {
  const SpeakerW& s = m->spk;
  const int C0 = s.init.cout, Cm = s.mfa.cout, Ccat = s.mfa.cin;
  int Cmax = C0;
  for (int i = 0; i < 3; ++i) Cmax = std::max(Cmax, s.blk[i].tdnn1.cout);
  DBuf h0, cat, t1, r2, t2, stat, se_a, se_b, hm, attn_in, a1, a2, pooled;
  h0.alloc((size_t)C0 * T * 4); cat.alloc((size_t)Ccat * T * 4);
  t1.alloc((size_t)Cmax * T * 4); r2.alloc((size_t)Cmax * T * 4); t2.alloc((size_t)Cmax * T * 4);
  stat.alloc((size_t)2 * std::max(Cmax, Cm) * 4); se_a.alloc((size_t)Cmax * 4); se_b.alloc((size_t)Cmax * 4);
  hm.alloc((size_t)Cm * T * 4); attn_in.alloc((size_t)3 * Cm * T * 4);
  a1.alloc((size_t)s.asp_tdnn.cout * T * 4); a2.alloc((size_t)Cm * T * 4); pooled.alloc((size_t)2 * Cm * 4);
  // blocks[0]: initial TDNN (speaker.rs:450)
  launch_reflect_conv(s.init, mel, nullptr, h0.as<float>(), 1, T, st);
  const float* x = h0.as<float>();
  size_t cat_off = 0;
  for (int i = 0; i < 3; ++i) {
    const SpkBlock& k = s.blk[i];
    const int C = k.tdnn1.cout, cs = k.branches[0].cout;
    Q3_REQUIRE(k.tdnn1.cin == (i == 0 ? C0 : s.blk[i - 1].tdnn1.cout) && k.tdnn1.cin == C, Q3_ERR_UNSUPPORTED,
               "speaker encoder: the residual connection needs equal channel counts");
    launch_conv(k.tdnn1, x, t1.as<float>(), 1, T, 1, nullptr, nullptr, nullptr, CEPI_RELU, st);
    // Res2Net (speaker.rs:170-189): chunk 0 passes through; branch j reads chunk j + 1 (+ the previous branch's output)
    Q3_CHECK_CUDA(cudaMemcpyAsync(r2.p, t1.p, (size_t)cs * T * 4, cudaMemcpyDeviceToDevice, st));
    for (int j = 0; j < (int)k.branches.size(); ++j)
      launch_reflect_conv(k.branches[j], t1.as<float>() + (size_t)(j + 1) * cs * T, j > 0 ? r2.as<float>() + (size_t)j * cs * T : nullptr,
                          r2.as<float>() + (size_t)(j + 1) * cs * T, k.dil, T, st);
    launch_conv(k.tdnn2, r2.as<float>(), t2.as<float>(), 1, T, 1, nullptr, nullptr, nullptr, CEPI_RELU, st);
    // squeeze-excitation (speaker.rs:214-221) + residual (:266), written straight into its slice of the MFA concatenation
    spk_channel_stats_kernel<<<C, 256, 0, st>>>(t2.as<float>(), stat.as<float>(), nullptr, T);
    Q3_COUNT_LAUNCH();
    launch_conv(k.se1, stat.as<float>(), se_a.as<float>(), 1, 1, 1, nullptr, nullptr, nullptr, CEPI_RELU, st);
    launch_conv(k.se2, se_a.as<float>(), se_b.as<float>(), 1, 1, 1, nullptr, nullptr, nullptr, CEPI_NONE, st);
    float* o = cat.as<float>() + cat_off;
    spk_se_apply_kernel<<<std::min(1024, ceil_div(C * T, 256)), 256, 0, st>>>(t2.as<float>(), se_b.as<float>(), x, o, C, T);
    Q3_COUNT_LAUNCH();
    Q3_LAUNCH_CHECK();
    x = o;
    cat_off += (size_t)C * T;
  }
  // MFA (speaker.rs:460-464), ASP (speaker.rs:294-346), FC (:470)
  launch_conv(s.mfa, cat.as<float>(), hm.as<float>(), 1, T, 1, nullptr, nullptr, nullptr, CEPI_RELU, st);
  spk_channel_stats_kernel<<<Cm, 256, 0, st>>>(hm.as<float>(), stat.as<float>(), stat.as<float>() + Cm, T);
  Q3_COUNT_LAUNCH();
  Q3_CHECK_CUDA(cudaMemcpyAsync(attn_in.p, hm.p, (size_t)Cm * T * 4, cudaMemcpyDeviceToDevice, st));
  spk_broadcast_stats_kernel<<<std::min(1024, ceil_div(2 * Cm * T, 256)), 256, 0, st>>>(stat.as<float>(), stat.as<float>() + Cm,
                                                                                        attn_in.as<float>() + (size_t)Cm * T, Cm, T);
  Q3_COUNT_LAUNCH();
  launch_conv(s.asp_tdnn, attn_in.as<float>(), a1.as<float>(), 1, T, 1, nullptr, nullptr, nullptr, CEPI_RELU, st);
  spk_tanh_kernel<<<std::min(1024, ceil_div(s.asp_tdnn.cout * T, 256)), 256, 0, st>>>(a1.as<float>(), (size_t)s.asp_tdnn.cout * T);
  Q3_COUNT_LAUNCH();
  launch_conv(s.asp_conv, a1.as<float>(), a2.as<float>(), 1, T, 1, nullptr, nullptr, nullptr, CEPI_NONE, st);
  spk_asp_pool_kernel<<<Cm, 256, 0, st>>>(hm.as<float>(), a2.as<float>(), pooled.as<float>(), Cm, T);
  Q3_COUNT_LAUNCH();
  Q3_LAUNCH_CHECK();
  launch_conv(s.fc, pooled.as<float>(), out, 1, 1, 1, nullptr, nullptr, nullptr, CEPI_NONE, st);
  Q3_CHECK_CUDA(cudaStreamSynchronize(st));       // the scratch buffers above are freed on return
}
