/*
 * q3tts.hpp -- header-only C++17 host side above the C ABI (q3tts.h), mirroring the reference's Rust API for the
 * decode path so that a compiled caller reads like the reference's own code.  The reference is a Rust crate and this
 * image has no Rust toolchain; this is the compiled-language host mirror (the Python one is qwen3_tts_rs_b200/api.py,
 * and the two are tested against each other).  Nothing here computes on the CPU: every model operation is a call
 * into libq3tts_b200.so and throws q3tts::Error when the library reports a failure (no GPU => Q3_ERR_CUDA).
 *
 *   q3tts::Qwen3TTS::from_pretrained      ref: src/lib.rs:183-262 (config.json, model.safetensors, speech_tokenizer/)
 *   Qwen3TTS::synthesize_with_voice       ref: src/lib.rs:718-784   (token ids in: tokenisation is outside the path)
 *   Qwen3TTS::synthesize_voice_design     ref: src/lib.rs:802-870
 *   Qwen3TTS::generate_codes              ref: src/lib.rs:530-656
 *   Qwen3TTS::synthesize_voice_clone      ref: src/lib.rs:895-1060  (speaker embedding / reference codes in: the encoders are outside the path)
 *   Qwen3TTS::decode_codes                ref: src/lib.rs:881-890
 *   Qwen3TTS::synthesize_streaming        ref: src/lib.rs:1070-1093 -> StreamingSession (src/lib.rs:1484-1782)
 *   SynthesisOptions                      ref: src/lib.rs:1786-1836
 *   AudioBuffer (+ save/load/normalize)   ref: src/audio/io.rs:27-165
 *   codes_to_tensor                       ref: src/lib.rs:1417-1431
 *   Speaker / Language ids                ref: src/models/talker.rs:96-105, 147-156
 *   ParsedModelConfig                     ref: src/models/config.rs:205-353
 *   save_codes_binary / save_audio_binary / compare_with_reference   ref: src/bin/generate_audio.rs:788-920
 *   SafeTensorsFile                       ref: candle_core::safetensors::load (src/lib.rs:1390-1396)
 */
#ifndef Q3TTS_HPP
#define Q3TTS_HPP

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "q3tts.h"

namespace q3tts {

// ---- errors ------------------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
  q3_status code;
  Error(q3_status c, const std::string& msg) : std::runtime_error(msg), code(c) {}
};
inline void check(q3_status st) {
  if (st != Q3_OK) throw Error(st, q3_last_error() ? q3_last_error() : "unknown error");
}

// ---- token tables (ref: src/models/talker.rs:31-54, 96-105, 147-156) -------------------------------------------------
namespace tok {
constexpr int32_t IM_START = 151644, ASSISTANT = 77091, NEWLINE = 198, TTS_PAD = 151671, TTS_BOS = 151672, TTS_EOS = 151673;
constexpr int32_t CODEC_PAD = 2148, CODEC_BOS = 2149, CODEC_EOS = 2150, CODEC_THINK = 2154, CODEC_NOTHINK = 2155,
                  CODEC_THINK_BOS = 2156, CODEC_THINK_EOS = 2157;
}  // namespace tok

enum class Language : int32_t {
  Chinese = 2055, English = 2050, Japanese = 2058, Korean = 2064, German = 2053, French = 2061, Russian = 2069,
  Portuguese = 2071, Spanish = 2054, Italian = 2070
};
enum class Speaker : int32_t {
  Serena = 3066, Vivian = 3065, UncleFu = 3010, Ryan = 3061, Aiden = 2861, OnoAnna = 2873, Sohee = 2864, Eric = 2875,
  Dylan = 2878
};
inline std::optional<Speaker> speaker_from_name(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
  static const std::map<std::string, Speaker> m = {
      {"serena", Speaker::Serena}, {"vivian", Speaker::Vivian}, {"uncle_fu", Speaker::UncleFu}, {"ryan", Speaker::Ryan},
      {"aiden", Speaker::Aiden},   {"ono_anna", Speaker::OnoAnna}, {"sohee", Speaker::Sohee},   {"eric", Speaker::Eric},
      {"dylan", Speaker::Dylan},   {"unclefu", Speaker::UncleFu},  {"onoanna", Speaker::OnoAnna}};  // talker.rs:121-137
  auto it = m.find(s);
  return it == m.end() ? std::nullopt : std::optional<Speaker>(it->second);
}
inline std::optional<Language> language_from_name(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
  static const std::map<std::string, Language> m = {
      {"chinese", Language::Chinese}, {"english", Language::English}, {"japanese", Language::Japanese},
      {"korean", Language::Korean},   {"german", Language::German},   {"french", Language::French},
      {"russian", Language::Russian}, {"portuguese", Language::Portuguese}, {"spanish", Language::Spanish},
      {"italian", Language::Italian}, {"en", Language::English}, {"zh", Language::Chinese}, {"ja", Language::Japanese},
      {"ko", Language::Korean}, {"de", Language::German}, {"fr", Language::French}, {"ru", Language::Russian},
      {"pt", Language::Portuguese}, {"es", Language::Spanish}, {"it", Language::Italian}};  // talker.rs:71-90
  auto it = m.find(s);
  return it == m.end() ? std::nullopt : std::optional<Language>(it->second);
}

using FrameCodes = std::vector<std::vector<uint32_t>>;  // [n_frames][16]  (ref: src/lib.rs:119-121)

// ---- options (ref: src/lib.rs:1786-1836, identical defaults) -----------------------------------------------------------
struct SynthesisOptions {
  int32_t max_length = 2048;
  double temperature = 0.9;
  int32_t top_k = 50;
  double top_p = 0.9;
  double repetition_penalty = 1.05;
  std::optional<int32_t> eos_token_id = Q3_CODEC_EOS;
  int32_t chunk_frames = 10;
  int32_t min_new_tokens = 2;
  std::optional<uint64_t> seed;
  /* not in the reference (opt-in, see q3_session_set_stream_context): 0 = the reference's stateless chunks, < 0 = stateful
     streaming (carried vocoder state, streamed PCM == non-streamed PCM), c > 0 = c left-context frames decoded again */
  int32_t stream_left_context = 0;
  /* not in the reference (opt-in, see q3_session_set_first_chunk): frames of the first streamed chunk, 0 = chunk_frames */
  int32_t stream_first_chunk = 0;

  q3_gen_config to_gen_config() const {
    q3_gen_config g{};
    g.max_new_tokens = max_length;
    g.temperature = temperature;
    g.top_k = top_k;
    g.top_p = top_p;
    g.repetition_penalty = repetition_penalty;
    g.eos_token_id = eos_token_id ? *eos_token_id : -1;
    g.min_new_tokens = min_new_tokens;
    g.chunk_frames = chunk_frames;
    return g;
  }
};

struct SynthesisTiming {  // ref: src/lib.rs:136-147
  double prefill_ms = 0, generation_ms = 0, decode_ms = 0;
  int32_t generation_frames = 0;
};

// ---- byte formats ----------------------------------------------------------------------------------------------------
/* ref: src/lib.rs:1417-1431 -- [n_frames][16] u32 -> i64 [1,16,T], data[q*T + f] = codes[f][q]. */
inline std::vector<int64_t> codes_to_tensor(const FrameCodes& codes) {
  const size_t t = codes.size();
  std::vector<int64_t> out(16 * t, 0);
  for (size_t f = 0; f < t; ++f) {
    if (codes[f].size() != 16) throw Error(Q3_ERR_INVALID, "codes_to_tensor: a frame must hold 16 codes");
    for (size_t q = 0; q < 16; ++q) out[q * t + f] = (int64_t)codes[f][q];
  }
  return out;
}

inline void write_file(const std::string& path, const void* data, size_t n) {
  std::ofstream f(path, std::ios::binary);
  if (!f) throw Error(Q3_ERR_INVALID, "cannot create " + path);
  f.write((const char*)data, (std::streamsize)n);
  if (!f) throw Error(Q3_ERR_INVALID, "short write to " + path);
}
inline std::vector<uint8_t> read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error(Q3_ERR_INVALID, "cannot open " + path);
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
inline bool file_exists(const std::string& p) {
  struct stat st;
  return ::stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
template <class T>
inline void put_le(std::vector<uint8_t>& b, T v) {
  for (size_t i = 0; i < sizeof(T); ++i) b.push_back((uint8_t)((uint64_t)v >> (8 * i)));
}
template <class T>
inline T get_le(const uint8_t* p) {
  uint64_t v = 0;
  for (size_t i = 0; i < sizeof(T); ++i) v |= (uint64_t)p[i] << (8 * i);
  return (T)v;
}

/* ref: generate_audio.rs:788-801 -- i64 little-endian, frame-major. */
inline void save_codes_binary(const FrameCodes& codes, const std::string& path) {
  std::vector<uint8_t> b;
  b.reserve(codes.size() * 16 * 8);
  for (const auto& fr : codes)
    for (uint32_t c : fr) put_le<uint64_t>(b, (uint64_t)(int64_t)c);
  write_file(path, b.data(), b.size());
}
inline FrameCodes load_codes_binary(const std::string& path) {
  auto raw = read_file(path);
  if (raw.size() % (16 * 8)) throw Error(Q3_ERR_INVALID, path + ": not a whole number of 16-code i64 frames");
  FrameCodes out(raw.size() / 128, std::vector<uint32_t>(16));
  for (size_t i = 0; i < raw.size() / 8; ++i) out[i / 16][i % 16] = (uint32_t)get_le<uint64_t>(raw.data() + 8 * i);
  return out;
}
/* ref: generate_audio.rs:803-813 -- f32 little-endian. */
inline void save_audio_binary(const std::vector<float>& s, const std::string& path) {
  std::vector<uint8_t> b;
  b.reserve(s.size() * 4);
  for (float x : s) {
    uint32_t u;
    std::memcpy(&u, &x, 4);
    put_le<uint32_t>(b, u);
  }
  write_file(path, b.data(), b.size());
}
inline std::vector<float> load_audio_binary(const std::string& path) {
  auto raw = read_file(path);
  std::vector<float> out(raw.size() / 4);  // chunks_exact(4)
  for (size_t i = 0; i < out.size(); ++i) {
    uint32_t u = get_le<uint32_t>(raw.data() + 4 * i);
    std::memcpy(&out[i], &u, 4);
  }
  return out;
}

/* ref: src/audio/io.rs:155-160 -- `(sample.clamp(-1.0, 1.0) * 32767.0) as i16`: f32 multiply, truncation, NaN -> 0. */
inline int16_t pcm_f32_to_i16(float x) {
  if (std::isnan(x)) return 0;
  float c = x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x);
  return (int16_t)(c * 32767.0f);
}

/* f32 -> bf16 bits, round to nearest even (what Tensor::to_dtype(BF16) does to the speaker embedding, lib.rs:930) */
inline uint16_t f32_to_bf16_bits(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);   // NaN stays NaN
  return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

struct AudioBuffer {  // ref: src/audio/io.rs:27-33
  std::vector<float> samples;
  uint32_t sample_rate = 24000;
  AudioBuffer() = default;
  AudioBuffer(std::vector<float> s, uint32_t rate) : samples(std::move(s)), sample_rate(rate) {}
  float duration() const { return (float)samples.size() / (float)sample_rate; }
  size_t len() const { return samples.size(); }
  bool is_empty() const { return samples.empty(); }

  /* ref: io.rs:143-165 (hound WavWriter, PCM16 mono): canonical 44-byte header. */
  void save(const std::string& path) const {
    std::vector<uint8_t> b;
    const uint32_t nbytes = (uint32_t)(samples.size() * 2);
    b.insert(b.end(), {'R', 'I', 'F', 'F'});
    put_le<uint32_t>(b, 36 + nbytes);
    b.insert(b.end(), {'W', 'A', 'V', 'E', 'f', 'm', 't', ' '});
    put_le<uint32_t>(b, 16);
    put_le<uint16_t>(b, 1);
    put_le<uint16_t>(b, 1);
    put_le<uint32_t>(b, sample_rate);
    put_le<uint32_t>(b, sample_rate * 2);
    put_le<uint16_t>(b, 2);
    put_le<uint16_t>(b, 16);
    b.insert(b.end(), {'d', 'a', 't', 'a'});
    put_le<uint32_t>(b, nbytes);
    for (float x : samples) put_le<uint16_t>(b, (uint16_t)pcm_f32_to_i16(x));
    write_file(path, b.data(), b.size());
  }

  /* ref: io.rs:110-141: int PCM / 2^(bits-1), float as is, channels averaged to mono. */
  static AudioBuffer load(const std::string& path) {
    auto raw = read_file(path);
    if (raw.size() < 12 || std::memcmp(raw.data(), "RIFF", 4) || std::memcmp(raw.data() + 8, "WAVE", 4))
      throw Error(Q3_ERR_INVALID, "Failed to open WAV file: " + path + ": not a RIFF/WAVE file");
    size_t pos = 12, fmt_at = 0, fmt_len = 0, data_at = 0, data_len = 0;
    while (pos + 8 <= raw.size()) {
      uint32_t sz = get_le<uint32_t>(raw.data() + pos + 4);
      size_t avail = std::min<size_t>(sz, raw.size() - pos - 8);
      if (!std::memcmp(raw.data() + pos, "fmt ", 4)) { fmt_at = pos + 8; fmt_len = avail; }
      if (!std::memcmp(raw.data() + pos, "data", 4)) { data_at = pos + 8; data_len = avail; break; }
      pos += 8 + (size_t)sz + (sz & 1);
    }
    if (!fmt_at || !data_at || fmt_len < 16) throw Error(Q3_ERR_INVALID, "Failed to open WAV file: " + path + ": missing fmt or data chunk");
    const uint8_t* f = raw.data() + fmt_at;
    uint16_t tag = get_le<uint16_t>(f), channels = get_le<uint16_t>(f + 2), bits = get_le<uint16_t>(f + 14);
    uint32_t rate = get_le<uint32_t>(f + 4);
    if (tag == 0xFFFE && fmt_len >= 26) tag = get_le<uint16_t>(f + 24);
    const uint8_t* d = raw.data() + data_at;
    std::vector<float> x;
    if (tag == 3 && bits == 32) {
      x.resize(data_len / 4);
      for (size_t i = 0; i < x.size(); ++i) { uint32_t u = get_le<uint32_t>(d + 4 * i); std::memcpy(&x[i], &u, 4); }
    } else if (tag == 1 && (bits == 8 || bits == 16 || bits == 24 || bits == 32)) {
      const size_t bytes = bits / 8;
      const float max_val = (float)(1ll << (bits - 1));
      x.resize(data_len / bytes);
      for (size_t i = 0; i < x.size(); ++i) {
        int64_t v;
        if (bits == 8) v = (int64_t)d[i] - 128;
        else if (bits == 16) v = (int16_t)get_le<uint16_t>(d + 2 * i);
        else if (bits == 24) { v = d[3 * i] | (d[3 * i + 1] << 8) | (d[3 * i + 2] << 16); if (v >= (1 << 23)) v -= (1 << 24); }
        else v = (int32_t)get_le<uint32_t>(d + 4 * i);
        x[i] = (float)v / max_val;
      }
    } else {
      throw Error(Q3_ERR_UNSUPPORTED, path + ": unsupported WAV sample format");
    }
    if (channels > 1) {
      std::vector<float> mono(x.size() / channels);
      for (size_t i = 0; i < mono.size(); ++i) {
        float s = 0.f;
        for (uint16_t c = 0; c < channels; ++c) s += x[i * channels + c];
        mono[i] = s / (float)channels;
      }
      x.swap(mono);
    }
    return AudioBuffer(std::move(x), rate);
  }

  void normalize() {  // ref: io.rs:83-91
    float m = 0.f;
    for (float s : samples) m = std::max(m, std::fabs(s));
    if (m > 0.f && m != 1.0f)
      for (float& s : samples) s /= m;
  }
  void normalize_db(float target_db) {  // ref: io.rs:94-103
    float m = 0.f;
    for (float s : samples) m = std::max(m, std::fabs(s));
    if (m > 0.f) {
      const float scale = std::pow(10.0f, target_db / 20.0f) / m;
      for (float& s : samples) s *= scale;
    }
  }
};

struct CompareReport {  // ref: generate_audio.rs:816-920, as values
  bool codes_found = false, codes_match = false, audio_found = false;
  size_t n_ref_codes = 0, n_our_codes = 0, n_code_diffs = 0, n_audio_compared = 0;
  float max_diff = 0.f;
  double mean_diff = 0.0, rmse = 0.0;
};
inline CompareReport compare_with_reference(const std::string& dir, uint64_t seed, size_t num_frames, const FrameCodes& codes,
                                            const std::vector<float>& audio) {
  CompareReport r;
  const std::string stem = "_seed" + std::to_string(seed) + "_frames" + std::to_string(num_frames) + ".bin";
  if (file_exists(dir + "/codes" + stem)) {
    r.codes_found = true;
    auto raw = read_file(dir + "/codes" + stem);
    r.n_ref_codes = raw.size() / 8;
    r.n_our_codes = codes.size() * 16;
    const size_t m = std::min(r.n_ref_codes, r.n_our_codes);
    for (size_t i = 0; i < m; ++i)
      if ((int64_t)get_le<uint64_t>(raw.data() + 8 * i) != (int64_t)codes[i / 16][i % 16]) ++r.n_code_diffs;
    r.codes_match = r.n_ref_codes == r.n_our_codes && r.n_code_diffs == 0;
  }
  if (file_exists(dir + "/audio" + stem)) {
    r.audio_found = true;
    auto ref = load_audio_binary(dir + "/audio" + stem);
    const size_t m = std::min(ref.size(), audio.size());
    r.n_audio_compared = m;
    double sum = 0, sq = 0;
    for (size_t i = 0; i < m; ++i) {
      float d = std::fabs(ref[i] - audio[i]);
      r.max_diff = std::max(r.max_diff, d);
      sum += (double)d;
      sq += (double)(d * d);
    }
    if (m) { r.mean_diff = sum / (double)m; r.rmse = std::sqrt(sq / (double)m); }
  }
  return r;
}

// ---- minimal JSON (config.json, safetensors headers) -------------------------------------------------------------------
namespace json {
struct Value;
using Object = std::vector<std::pair<std::string, Value>>;
struct Value {
  enum Kind { Null, Bool, Int, Float, String, Array, Obj } kind = Null;
  bool b = false;
  int64_t i = 0;
  double d = 0.0;
  std::string s;
  std::vector<Value> a;
  Object o;
  const Value* get(const std::string& key) const {
    if (kind != Obj) return nullptr;
    for (const auto& kv : o)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  /* serde_json `v[key]` on a non-object or a missing key is Null */
  const Value& operator[](const std::string& key) const {
    static const Value null_value;
    const Value* v = get(key);
    return v ? *v : null_value;
  }
  std::optional<uint64_t> as_u64() const { return kind == Int && i >= 0 ? std::optional<uint64_t>((uint64_t)i) : std::nullopt; }
  std::optional<double> as_f64() const {
    if (kind == Int) return (double)i;
    if (kind == Float) return d;
    return std::nullopt;
  }
  std::optional<std::string> as_str() const { return kind == String ? std::optional<std::string>(s) : std::nullopt; }
};
class Parser {
 public:
  explicit Parser(const std::string& t) : t_(t) {}
  Value parse() {
    Value v = value();
    ws();
    if (p_ != t_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& t_;
  size_t p_ = 0;
  [[noreturn]] void fail(const std::string& m) const { throw Error(Q3_ERR_INVALID, "JSON: " + m + " at byte " + std::to_string(p_)); }
  void ws() { while (p_ < t_.size() && (t_[p_] == ' ' || t_[p_] == '\n' || t_[p_] == '\t' || t_[p_] == '\r')) ++p_; }
  bool lit(const char* s) {
    size_t n = std::strlen(s);
    if (t_.compare(p_, n, s) == 0) { p_ += n; return true; }
    return false;
  }
  Value value() {
    ws();
    if (p_ >= t_.size()) fail("unexpected end");
    char c = t_[p_];
    Value v;
    if (c == '{') {
      v.kind = Value::Obj;
      ++p_; ws();
      if (p_ < t_.size() && t_[p_] == '}') { ++p_; return v; }
      for (;;) {
        ws();
        if (p_ >= t_.size() || t_[p_] != '"') fail("expected a key");
        std::string k = string();
        ws();
        if (p_ >= t_.size() || t_[p_] != ':') fail("expected ':'");
        ++p_;
        v.o.emplace_back(std::move(k), value());
        ws();
        if (p_ < t_.size() && t_[p_] == ',') { ++p_; continue; }
        if (p_ < t_.size() && t_[p_] == '}') { ++p_; return v; }
        fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      v.kind = Value::Array;
      ++p_; ws();
      if (p_ < t_.size() && t_[p_] == ']') { ++p_; return v; }
      for (;;) {
        v.a.push_back(value());
        ws();
        if (p_ < t_.size() && t_[p_] == ',') { ++p_; continue; }
        if (p_ < t_.size() && t_[p_] == ']') { ++p_; return v; }
        fail("expected ',' or ']'");
      }
    }
    if (c == '"') { v.kind = Value::String; v.s = string(); return v; }
    if (lit("true")) { v.kind = Value::Bool; v.b = true; return v; }
    if (lit("false")) { v.kind = Value::Bool; return v; }
    if (lit("null")) return v;
    if (c == '-' || (c >= '0' && c <= '9')) {
      size_t s = p_;
      bool is_float = false;
      if (t_[p_] == '-') ++p_;
      while (p_ < t_.size() && ((t_[p_] >= '0' && t_[p_] <= '9') || t_[p_] == '.' || t_[p_] == 'e' || t_[p_] == 'E' || t_[p_] == '+' || t_[p_] == '-')) {
        if (t_[p_] == '.' || t_[p_] == 'e' || t_[p_] == 'E') is_float = true;
        ++p_;
      }
      const std::string num = t_.substr(s, p_ - s);
      try {
        if (is_float) { v.kind = Value::Float; v.d = std::stod(num); }
        else { v.kind = Value::Int; v.i = std::stoll(num); }
      } catch (const std::exception&) { fail("bad number"); }
      return v;
    }
    fail("unexpected character");
  }
  std::string string() {
    std::string out;
    ++p_;
    while (p_ < t_.size() && t_[p_] != '"') {
      char c = t_[p_++];
      if (c != '\\') { out.push_back(c); continue; }
      if (p_ >= t_.size()) fail("bad escape");
      char e = t_[p_++];
      switch (e) {
        case 'n': out.push_back('\n'); break;
        case 't': out.push_back('\t'); break;
        case 'r': out.push_back('\r'); break;
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'u': {
          auto hex4 = [&]() -> unsigned {
            if (p_ + 4 > t_.size()) fail("bad \\u escape");
            unsigned v = 0;
            for (int i = 0; i < 4; ++i) {
              const char h = t_[p_++];
              v <<= 4;
              if (h >= '0' && h <= '9') v |= (unsigned)(h - '0');
              else if (h >= 'a' && h <= 'f') v |= (unsigned)(h - 'a' + 10);
              else if (h >= 'A' && h <= 'F') v |= (unsigned)(h - 'A' + 10);
              else fail("bad \\u escape");
            }
            return v;
          };
          unsigned cp = hex4();
          if (cp >= 0xD800 && cp <= 0xDBFF && p_ + 1 < t_.size() && t_[p_] == '\\' && t_[p_ + 1] == 'u') {  // surrogate pair
            const size_t save = p_;
            p_ += 2;
            const unsigned lo = hex4();
            if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            else p_ = save;
          }
          if (cp < 0x80) out.push_back((char)cp);
          else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
          else if (cp < 0x10000) { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
          else { out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
          break;
        }
        case '"': case '\\': case '/': out.push_back(e); break;
        default: fail("bad escape");
      }
    }
    if (p_ >= t_.size()) fail("unterminated string");
    ++p_;
    return out;
  }
};
inline Value parse(const std::string& text) { return Parser(text).parse(); }
}  // namespace json

// ---- config.json (ref: src/models/config.rs:236-353) -------------------------------------------------------------------
enum class ModelType { Base, CustomVoice, VoiceDesign };

struct ParsedModelConfig {
  ModelType model_type = ModelType::Base;
  std::string model_size = "unknown";
  size_t talker_hidden_size = 1024, talker_intermediate_size = 3072, talker_num_hidden_layers = 28,
         talker_num_attention_heads = 16, talker_num_key_value_heads = 8, talker_head_dim = 128, talker_vocab_size = 3072,
         talker_text_vocab_size = 151936, talker_text_hidden_size = 2048, talker_max_position_embeddings = 32768;
  double talker_rms_norm_eps = 1e-6, talker_rope_theta = 1000000.0;
  std::optional<std::array<size_t, 3>> mrope_section;
  size_t cp_hidden_size = 1024, cp_intermediate_size = 3072, cp_num_hidden_layers = 5, cp_num_attention_heads = 16,
         cp_num_key_value_heads = 8, cp_head_dim = 128, cp_vocab_size = 2048, cp_num_code_groups = 16;
  double cp_rms_norm_eps = 1e-6, cp_rope_theta = 1000000.0;
  std::optional<size_t> speaker_enc_dim;

  static ParsedModelConfig from_json(const std::string& text) {
    const json::Value v = json::parse(text);
    ParsedModelConfig c;
    const std::string mt = v["tts_model_type"].as_str().value_or("base");
    c.model_type = mt == "custom_voice" ? ModelType::CustomVoice : mt == "voice_design" ? ModelType::VoiceDesign : ModelType::Base;
    c.model_size = v["tts_model_size"].as_str().value_or("unknown");
    const json::Value& t = v["talker_config"];
    const json::Value& cp = t["code_predictor_config"];
    auto u = [](const json::Value& n, const char* k, size_t d) { return (size_t)n[k].as_u64().value_or(d); };
    auto f = [](const json::Value& n, const char* k, double d) { return n[k].as_f64().value_or(d); };
    c.talker_hidden_size = u(t, "hidden_size", 1024);
    c.talker_intermediate_size = u(t, "intermediate_size", 3072);
    c.talker_num_hidden_layers = u(t, "num_hidden_layers", 28);
    c.talker_num_attention_heads = u(t, "num_attention_heads", 16);
    c.talker_num_key_value_heads = u(t, "num_key_value_heads", 8);
    c.talker_head_dim = u(t, "head_dim", 128);
    c.talker_vocab_size = u(t, "vocab_size", 3072);
    c.talker_text_vocab_size = u(t, "text_vocab_size", 151936);
    c.talker_text_hidden_size = u(t, "text_hidden_size", 2048);
    c.talker_rms_norm_eps = f(t, "rms_norm_eps", 1e-6);
    c.talker_rope_theta = f(t, "rope_theta", 1000000.0);
    c.talker_max_position_embeddings = u(t, "max_position_embeddings", 32768);
    const json::Value& sec = t["rope_scaling"]["mrope_section"];
    if (sec.kind == json::Value::Array && sec.a.size() == 3 && sec.a[0].as_u64() && sec.a[1].as_u64() && sec.a[2].as_u64())
      c.mrope_section = std::array<size_t, 3>{(size_t)*sec.a[0].as_u64(), (size_t)*sec.a[1].as_u64(), (size_t)*sec.a[2].as_u64()};
    c.cp_hidden_size = u(cp, "hidden_size", 1024);
    c.cp_intermediate_size = u(cp, "intermediate_size", 3072);
    c.cp_num_hidden_layers = u(cp, "num_hidden_layers", 5);
    c.cp_num_attention_heads = u(cp, "num_attention_heads", 16);
    c.cp_num_key_value_heads = u(cp, "num_key_value_heads", 8);
    c.cp_head_dim = u(cp, "head_dim", 128);
    c.cp_vocab_size = u(cp, "vocab_size", 2048);
    c.cp_num_code_groups = u(cp, "num_code_groups", 16);
    c.cp_rms_norm_eps = f(cp, "rms_norm_eps", 1e-6);
    c.cp_rope_theta = f(cp, "rope_theta", 1000000.0);
    if (v["speaker_encoder_config"].kind == json::Value::Obj) c.speaker_enc_dim = u(v["speaker_encoder_config"], "enc_dim", 1024);
    return c;
  }
  static ParsedModelConfig from_file(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw Error(Q3_ERR_INVALID, "Failed to read config from " + path);
    std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    try {
      return from_json(text);
    } catch (const Error& e) {
      throw Error(Q3_ERR_INVALID, "Failed to parse config from " + path + ": " + e.what());
    }
  }
  std::string label() const {  // ref: config.rs:339-351
    const std::string size = model_size == "0b6" ? "0.6B" : model_size == "1b7" ? "1.7B" : model_size;
    const char* variant = model_type == ModelType::Base ? "Base" : model_type == ModelType::CustomVoice ? "CustomVoice" : "VoiceDesign";
    return size + " " + variant;
  }
};

/* Decoder12HzConfig::default (ref: src/models/codec/decoder_12hz.rs:47-67) + the talker / CP tables. */
inline q3_model_desc default_desc(bool large_1p7b, int device = 0) {
  q3_model_desc d{};
  d.hidden = large_1p7b ? 2048 : 1024;
  d.inter = large_1p7b ? 6144 : 3072;
  d.layers = 28; d.heads = 16; d.kv_heads = 8; d.head_dim = 128;
  d.codec_vocab = 3072; d.text_vocab = 151936; d.text_embed_dim = 2048;
  d.rope_theta = 1000000.0f; d.rms_eps = 1e-6f;
  d.cp_hidden = 1024; d.cp_inter = 3072; d.cp_layers = 5; d.cp_heads = 16; d.cp_kv_heads = 8; d.cp_vocab = 2048; d.groups = 16;
  d.cp_rope_positions = 1024; d.cp_max_seq = 17;
  d.v_codebook_dim = 512; d.v_vq_dim = 256; d.v_latent_dim = 1024; d.v_hidden = 512; d.v_layers = 8; d.v_heads = 16;
  d.v_head_dim = 64; d.v_inter = 1024; d.v_quantizers = 16; d.v_codebook_size = 2048; d.v_decoder_dim = 1536;
  d.v_n_upsampling = 2; d.v_upsampling[0] = 2; d.v_upsampling[1] = 2;
  d.v_n_rates = 4; d.v_rates[0] = 8; d.v_rates[1] = 5; d.v_rates[2] = 4; d.v_rates[3] = 3;
  d.v_rms_eps = 1e-5f; d.v_rope_theta = 10000.0f;
  d.device = device;
  return d;
}

/* TalkerConfig::from_parsed + CodePredictorConfig::from_parsed (ref: talker.rs:237-254, code_predictor.rs:72-91). */
inline q3_model_desc desc_from_config(const ParsedModelConfig& c, int device = 0) {
  if (c.talker_head_dim != 128 || c.cp_head_dim != 128) throw Error(Q3_ERR_UNSUPPORTED, "the decode kernels are built for head_dim 128");
  if (c.cp_rope_theta != c.talker_rope_theta || c.cp_rms_norm_eps != c.talker_rms_norm_eps)
    throw Error(Q3_ERR_UNSUPPORTED, "code predictor rope_theta / rms_norm_eps differ from the talker's");
  q3_model_desc d = default_desc(false, device);
  d.hidden = (int32_t)c.talker_hidden_size; d.inter = (int32_t)c.talker_intermediate_size;
  d.layers = (int32_t)c.talker_num_hidden_layers; d.heads = (int32_t)c.talker_num_attention_heads;
  d.kv_heads = (int32_t)c.talker_num_key_value_heads; d.codec_vocab = (int32_t)c.talker_vocab_size;
  d.text_vocab = (int32_t)c.talker_text_vocab_size; d.text_embed_dim = (int32_t)c.talker_text_hidden_size;
  d.rope_theta = (float)c.talker_rope_theta; d.rms_eps = (float)c.talker_rms_norm_eps;
  d.cp_hidden = (int32_t)c.cp_hidden_size; d.cp_inter = (int32_t)c.cp_intermediate_size; d.cp_layers = (int32_t)c.cp_num_hidden_layers;
  d.cp_heads = (int32_t)c.cp_num_attention_heads; d.cp_kv_heads = (int32_t)c.cp_num_key_value_heads;
  d.cp_vocab = (int32_t)c.cp_vocab_size; d.groups = (int32_t)c.cp_num_code_groups;
  return d;
}

/* speech_tokenizer/config.json "decoder_config" (an extension the reference never reads; see formats.py). */
inline void apply_vocoder_config(q3_model_desc& d, const std::string& text) {
  const json::Value v = json::parse(text);
  const json::Value& c = v["decoder_config"];
  auto u = [&](const char* k, int32_t& dst) { if (auto x = c[k].as_u64()) dst = (int32_t)*x; };
  u("codebook_dim", d.v_codebook_dim); u("vq_dim", d.v_vq_dim); u("latent_dim", d.v_latent_dim); u("hidden_size", d.v_hidden);
  u("num_hidden_layers", d.v_layers); u("num_attention_heads", d.v_heads); u("head_dim", d.v_head_dim);
  u("intermediate_size", d.v_inter); u("num_quantizers", d.v_quantizers); u("codebook_size", d.v_codebook_size);
  u("decoder_dim", d.v_decoder_dim);
  if (auto x = c["rms_norm_eps"].as_f64()) d.v_rms_eps = (float)*x;
  if (auto x = c["rope_theta"].as_f64()) d.v_rope_theta = (float)*x;
  const json::Value& up = c["upsampling_ratios"];
  if (up.kind == json::Value::Array && up.a.size() <= 4) {
    d.v_n_upsampling = (int32_t)up.a.size();
    for (size_t i = 0; i < up.a.size(); ++i) d.v_upsampling[i] = (int32_t)up.a[i].as_u64().value_or(1);
  }
  const json::Value& rt = c["upsample_rates"];
  if (rt.kind == json::Value::Array && rt.a.size() <= 8) {
    d.v_n_rates = (int32_t)rt.a.size();
    for (size_t i = 0; i < rt.a.size(); ++i) d.v_rates[i] = (int32_t)rt.a[i].as_u64().value_or(1);
  }
}

// ---- safetensors, memory-mapped ------------------------------------------------------------------------------------------
struct TensorEntry {
  std::string dtype;
  std::vector<int64_t> shape;
  size_t begin = 0, end = 0;
  size_t numel() const { size_t n = 1; for (auto s : shape) n *= (size_t)s; return n; }
};

class SafeTensorsFile {
 public:
  explicit SafeTensorsFile(const std::string& path) : path_(path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) throw Error(Q3_ERR_INVALID, "cannot open " + path);
    struct stat st;
    if (::fstat(fd_, &st) != 0 || st.st_size < 8) { ::close(fd_); throw Error(Q3_ERR_INVALID, path + ": too short for a safetensors header"); }
    size_ = (size_t)st.st_size;
    map_ = (const uint8_t*)::mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (map_ == MAP_FAILED) { ::close(fd_); throw Error(Q3_ERR_INVALID, "mmap failed for " + path); }
    try {
      const uint64_t n = get_le<uint64_t>(map_);
      if (n > size_ - 8 || n > 100000000ull) throw Error(Q3_ERR_INVALID, path + ": header length is out of range");
      base_ = 8 + (size_t)n;
      const json::Value hdr = json::parse(std::string((const char*)map_ + 8, (size_t)n));
      if (hdr.kind != json::Value::Obj) throw Error(Q3_ERR_INVALID, path + ": header is not a JSON object");
      for (const auto& kv : hdr.o) {
        if (kv.first == "__metadata__") continue;
        TensorEntry e;
        e.dtype = kv.second["dtype"].as_str().value_or("");
        for (const auto& s : kv.second["shape"].a) e.shape.push_back((int64_t)s.as_u64().value_or(0));
        const auto& off = kv.second["data_offsets"];
        if (off.kind != json::Value::Array || off.a.size() != 2) throw Error(Q3_ERR_INVALID, path + ": tensor " + kv.first + " has no data_offsets");
        e.begin = (size_t)off.a[0].as_u64().value_or(0);
        e.end = (size_t)off.a[1].as_u64().value_or(0);
        const size_t es = elem_size(e.dtype);
        if (es == 0) throw Error(Q3_ERR_UNSUPPORTED, path + ": tensor " + kv.first + " has unsupported dtype " + e.dtype);
        // element count with overflow guard: a crafted shape must not wrap around to a plausible byte count
        size_t count = 1;
        bool overflow = false;
        for (int64_t dim : e.shape) {
          if (dim < 0 || (dim != 0 && count > size_ / (size_t)dim)) { overflow = true; break; }
          count *= (size_t)dim;
        }
        if (overflow || e.begin > e.end || e.end > size_ - base_ || e.end - e.begin != count * es)
          throw Error(Q3_ERR_INVALID, path + ": tensor " + kv.first + " has bad data_offsets");
        entries_.emplace(kv.first, std::move(e));
      }
    } catch (...) {
      ::munmap((void*)map_, size_);
      ::close(fd_);
      throw;
    }
  }
  ~SafeTensorsFile() {
    if (map_ && map_ != MAP_FAILED) ::munmap((void*)map_, size_);
    if (fd_ >= 0) ::close(fd_);
  }
  SafeTensorsFile(const SafeTensorsFile&) = delete;
  SafeTensorsFile& operator=(const SafeTensorsFile&) = delete;

  static size_t elem_size(const std::string& dt) {
    if (dt == "F64" || dt == "I64") return 8;
    if (dt == "F32" || dt == "I32") return 4;
    if (dt == "F16" || dt == "BF16" || dt == "I16") return 2;
    if (dt == "I8" || dt == "U8" || dt == "BOOL") return 1;
    return 0;
  }
  const std::map<std::string, TensorEntry>& entries() const { return entries_; }
  const TensorEntry* find(const std::string& name) const {
    auto it = entries_.find(name);
    return it == entries_.end() ? nullptr : &it->second;
  }
  const uint8_t* data(const TensorEntry& e) const { return map_ + base_ + e.begin; }
  const std::string& path() const { return path_; }

 private:
  std::string path_;
  int fd_ = -1;
  size_t size_ = 0, base_ = 0;
  const uint8_t* map_ = nullptr;
  std::map<std::string, TensorEntry> entries_;
};

// ---- model / session handles -----------------------------------------------------------------------------------------------
class Model {
 public:
  explicit Model(const q3_model_desc& d) : desc_(d) { check(q3_model_create(&d, &h_)); }
  ~Model() { if (h_) q3_model_destroy(h_); }
  Model(const Model&) = delete;
  Model& operator=(const Model&) = delete;

  void set_tensor(const std::string& name, const void* data, q3_dtype dt, const std::vector<int64_t>& shape, bool on_device = false) {
    check(q3_model_set_tensor(h_, name.c_str(), data, dt, shape.data(), (int32_t)shape.size(), on_device ? 1 : 0));
  }
  /* Upload every tensor of `f` whose name starts with `prefix`, straight from the file mapping: bf16 and f32 as they
   * are, f16 widened to f32 on the way (the library stores talker.* as bf16 and decoder.* as f32 whatever arrives). */
  size_t load_from(const SafeTensorsFile& f, const std::string& prefix) {
    size_t n = 0;
    for (const auto& kv : f.entries()) {
      if (kv.first.compare(0, prefix.size(), prefix) != 0) continue;
      const TensorEntry& e = kv.second;
      std::vector<int64_t> shape = e.shape.empty() ? std::vector<int64_t>{1} : e.shape;
      if (e.dtype == "BF16") set_tensor(kv.first, f.data(e), Q3_BF16, shape);
      else if (e.dtype == "F32") set_tensor(kv.first, f.data(e), Q3_F32, shape);
      else if (e.dtype == "F16") {
        std::vector<float> w(e.numel());
        const uint8_t* p = f.data(e);
        for (size_t i = 0; i < w.size(); ++i) w[i] = half_to_float(get_le<uint16_t>(p + 2 * i));
        set_tensor(kv.first, w.data(), Q3_F32, shape);
      } else {
        continue;  // integer buffers are not weights of the decode path
      }
      ++n;
    }
    return n;
  }
  void finalize() { check(q3_model_finalize(h_)); }
  q3_model* handle() const { return h_; }
  const q3_model_desc& desc() const { return desc_; }
  int total_upsample() const {
    int n = 1;
    for (int i = 0; i < desc_.v_n_upsampling; ++i) n *= desc_.v_upsampling[i];
    for (int i = 0; i < desc_.v_n_rates; ++i) n *= desc_.v_rates[i];
    return n;
  }
  static float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000) << 16;
    uint32_t exp = (h >> 10) & 0x1F, man = h & 0x3FF, u;
    if (exp == 0) {
      if (man == 0) u = sign;
      else {
        int e = -1;
        do { ++e; man <<= 1; } while (!(man & 0x400));
        u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FF) << 13);
      }
    } else if (exp == 31) u = sign | 0x7F800000u | (man << 13);
    else u = sign | ((exp + 112) << 23) | (man << 13);
    float f;
    std::memcpy(&f, &u, 4);
    return f;
  }

 private:
  q3_model_desc desc_;
  q3_model* h_ = nullptr;
};

class Session {
 public:
  Session(const Model& m, int32_t batch, int32_t max_seq, const SynthesisOptions& o, const std::vector<uint64_t>& seeds)
      : model_(m), batch_(batch), opts_(o) {
    if ((int32_t)seeds.size() != batch) throw Error(Q3_ERR_INVALID, "one seed per row is required");
    q3_gen_config g = o.to_gen_config();
    check(q3_session_create(m.handle(), batch, max_seq, &g, seeds.data(), &h_));
    if (o.stream_left_context != 0) {
      const q3_status st = q3_session_set_stream_context(h_, o.stream_left_context);
      if (st != Q3_OK) { q3_session_destroy(h_); h_ = nullptr; check(st); }
    }
    if (o.stream_first_chunk != 0) {
      const q3_status st = q3_session_set_first_chunk(h_, o.stream_first_chunk);
      if (st != Q3_OK) { q3_session_destroy(h_); h_ = nullptr; check(st); }
    }
  }
  ~Session() { if (h_) q3_session_destroy(h_); }
  Session(const Session&) = delete;
  Session& operator=(const Session&) = delete;

  /* rows of (text id, codec id) pairs, -1 = absent (ref: talker.rs:451-491, 585-627) */
  void prefill_ids(const std::vector<std::vector<int32_t>>& text, const std::vector<std::vector<int32_t>>& codec) {
    std::vector<int32_t> lens(batch_);
    int32_t lmax = 0;
    for (int b = 0; b < batch_; ++b) { lens[b] = (int32_t)text[b].size(); lmax = std::max(lmax, lens[b]); }
    std::vector<int32_t> ti((size_t)batch_ * lmax, -1), ci((size_t)batch_ * lmax, -1);
    for (int b = 0; b < batch_; ++b)
      for (int p = 0; p < lens[b]; ++p) { ti[(size_t)b * lmax + p] = text[b][p]; ci[(size_t)b * lmax + p] = codec[b][p]; }
    check(q3_prefill_ids(h_, ti.data(), ci.data(), lens.data(), lmax));
  }
  /* batch-1 voice-clone prefill (ref: prefill_voice_clone talker.rs:511-564 ++ build_icl_prompt talker.rs:646-705): codec parts
     may be Q3_POS_SPEAKER / Q3_POS_REF_FRAME(t); speaker: bf16 bits [hidden]; ref_codes: [t_ref][16] or empty */
  void prefill_voice_clone(const std::vector<int32_t>& text, const std::vector<int32_t>& codec, const std::vector<uint16_t>& speaker,
                           const std::vector<uint32_t>& ref_codes) {
    if (batch_ != 1) throw Error(Q3_ERR_INVALID, "prefill_voice_clone: one utterance per session in this mirror");
    const int32_t len = (int32_t)text.size(), t_ref = (int32_t)(ref_codes.size() / 16);
    check(q3_prefill_voice_clone(h_, text.data(), codec.data(), &len, len, speaker.data(), ref_codes.empty() ? nullptr : ref_codes.data(),
                                 &t_ref, t_ref));
  }
  /* no trailing rows at all: every frame adds tts_pad (an ICL prompt that consumed the whole text, talker.rs:691-703) */
  void set_trailing_none(int32_t tts_eos_id, int32_t tts_pad_id) {
    std::vector<int32_t> n(batch_, -1), buf(batch_, 0);
    check(q3_set_trailing_ids(h_, buf.data(), n.data(), 1, tts_eos_id, tts_pad_id));
  }
  /* ref: build_trailing_text, lib.rs:508-519 */
  void set_trailing_ids(const std::vector<std::vector<int32_t>>& ids, int32_t tts_eos_id, int32_t tts_pad_id) {
    std::vector<int32_t> n(batch_);
    int32_t nmax = 1;
    for (int b = 0; b < batch_; ++b) { n[b] = (int32_t)ids[b].size(); nmax = std::max(nmax, n[b]); }
    std::vector<int32_t> buf((size_t)batch_ * nmax, 0);
    for (int b = 0; b < batch_; ++b) std::copy(ids[b].begin(), ids[b].end(), buf.begin() + (size_t)b * nmax);
    check(q3_set_trailing_ids(h_, buf.data(), n.data(), nmax, tts_eos_id, tts_pad_id));
  }
  std::vector<FrameCodes> generate(int32_t max_frames) {
    std::vector<uint32_t> codes((size_t)batch_ * max_frames * 16);
    n_frames_.assign(batch_, 0);
    check(q3_generate(h_, max_frames, codes.data(), n_frames_.data()));
    std::vector<FrameCodes> out(batch_);
    for (int b = 0; b < batch_; ++b)
      for (int f = 0; f < n_frames_[b]; ++f) {
        const uint32_t* p = codes.data() + ((size_t)b * max_frames + f) * 16;
        out[b].emplace_back(p, p + 16);
      }
    return out;
  }
  std::vector<AudioBuffer> vocode(int32_t max_frames) {
    const size_t up = (size_t)model_.total_upsample();
    std::vector<float> pcm((size_t)batch_ * max_frames * up);
    check(q3_vocode_session(h_, max_frames, pcm.data()));
    std::vector<AudioBuffer> out;
    for (int b = 0; b < batch_; ++b) {
      const float* p = pcm.data() + (size_t)b * max_frames * up;
      out.emplace_back(std::vector<float>(p, p + (size_t)n_frames_[b] * up), 24000u);
    }
    return out;
  }
  /* one chunk for row 0 of a batch-1 streaming session; false when finished and nothing was produced */
  bool stream_next(FrameCodes& codes, std::vector<float>& pcm, bool& done) {
    const int32_t chunk = std::max(1, opts_.chunk_frames);
    const size_t up = (size_t)model_.total_upsample();
    std::vector<uint32_t> c((size_t)batch_ * chunk * 16);
    std::vector<float> p((size_t)batch_ * chunk * up);
    std::vector<int32_t> n(batch_, 0);
    int32_t d = 0;
    check(q3_stream_next(h_, c.data(), p.data(), n.data(), &d));
    done = d != 0;
    codes.clear();
    for (int f = 0; f < n[0]; ++f) codes.emplace_back(c.begin() + (size_t)f * 16, c.begin() + (size_t)(f + 1) * 16);
    pcm.assign(p.begin(), p.begin() + (size_t)n[0] * up);
    return n[0] > 0;
  }
  SynthesisTiming timing() {
    q3_timing t{};
    check(q3_session_timing(h_, &t));
    return SynthesisTiming{t.prefill_ms, t.generation_ms, t.decode_ms, t.generation_frames};
  }
  const std::vector<int32_t>& n_frames() const { return n_frames_; }

 private:
  const Model& model_;
  int32_t batch_;
  SynthesisOptions opts_;
  q3_session* h_ = nullptr;
  std::vector<int32_t> n_frames_;
};

// ---- prompts (host logic only; the embedding math runs on the device) ----------------------------------------------------------
struct Prompt {
  std::vector<int32_t> text, codec;  // position-wise pairs, -1 = absent
};
/* Special text ids sit at a fixed distance from the end of the text vocab; scaled-down test models keep that distance. */
inline int32_t special_text_id(int32_t text_vocab, int32_t t) {
  if (text_vocab == 151936) return t;
  return t >= 151643 ? t - 151936 + text_vocab : t % (text_vocab - 300);
}
/* ref: prefill_custom_voice, talker.rs:451-488: 3 role + 5 tts_pad/1 tts_bos over think/lang/speaker/pad + first text token. */
inline Prompt custom_voice_prompt(int32_t text_vocab, const std::vector<int32_t>& text_ids, Speaker sp, Language lang) {
  auto sid = [&](int32_t t) { return special_text_id(text_vocab, t); };
  Prompt p;
  p.text = {sid(tok::IM_START), sid(tok::ASSISTANT), sid(tok::NEWLINE), sid(tok::TTS_PAD), sid(tok::TTS_PAD), sid(tok::TTS_PAD),
            sid(tok::TTS_PAD), sid(tok::TTS_PAD), sid(tok::TTS_BOS)};
  p.codec = {-1, -1, -1, tok::CODEC_THINK, tok::CODEC_THINK_BOS, (int32_t)lang, tok::CODEC_THINK_EOS, (int32_t)sp, tok::CODEC_PAD};
  if (!text_ids.empty()) { p.text.push_back(text_ids[0]); p.codec.push_back(tok::CODEC_BOS); }
  return p;
}
/* ref: prefill_voice_design, talker.rs:585-624. */
inline Prompt voice_design_prompt(int32_t text_vocab, const std::vector<int32_t>& text_ids, const std::vector<int32_t>& instruct_ids,
                                  Language lang) {
  auto sid = [&](int32_t t) { return special_text_id(text_vocab, t); };
  Prompt p;
  p.text = instruct_ids;
  p.codec.assign(instruct_ids.size() + 3, -1);
  for (int32_t t : {tok::IM_START, tok::ASSISTANT, tok::NEWLINE, tok::TTS_PAD, tok::TTS_PAD, tok::TTS_PAD, tok::TTS_PAD, tok::TTS_BOS})
    p.text.push_back(sid(t));
  for (int32_t c : {tok::CODEC_THINK, tok::CODEC_THINK_BOS, (int32_t)lang, tok::CODEC_THINK_EOS, tok::CODEC_PAD}) p.codec.push_back(c);
  if (!text_ids.empty()) { p.text.push_back(text_ids[0]); p.codec.push_back(tok::CODEC_BOS); }
  return p;
}

/* ref: VoiceClonePrompt, lib.rs:127-134.  speaker_embedding: f32 [hidden]; ref_codes: [T_ref][16] (ICL mode) */
struct VoiceClonePrompt {
  std::vector<float> speaker_embedding;
  std::optional<FrameCodes> ref_codes;
  std::optional<std::vector<int32_t>> ref_text_ids;
  bool is_icl() const { return ref_codes.has_value() && ref_text_ids.has_value(); }
};
constexpr int32_t ICL_MIN_FRAMES = 75, ICL_FRAMES_PER_TOKEN = 6;   // lib.rs:1472-1475
constexpr double ICL_MIN_REPETITION_PENALTY = 1.5;                 // lib.rs:1478
struct ClonePrompt {
  Prompt p;
  std::vector<int32_t> trailing;   // ids for set_trailing_ids (tts_eos is appended there)
  bool no_trailing = false;        // ICL prompt that consumed the whole text: every frame adds tts_pad
};
/* ref: prefill_voice_clone (talker.rs:511-564) followed, in ICL mode, by the streaming overlay of build_icl_prompt
   (talker.rs:646-705; lib.rs:953-987 runs it as a second causal chunk -- one causal prefill over the concatenation is the same
   computation). */
inline ClonePrompt voice_clone_prompt(int32_t text_vocab, const std::vector<int32_t>& text_ids, const VoiceClonePrompt& vc, Language lang) {
  auto sid = [&](int32_t t) { return special_text_id(text_vocab, t); };
  ClonePrompt c;
  c.p.text = {sid(tok::IM_START), sid(tok::ASSISTANT), sid(tok::NEWLINE), sid(tok::TTS_PAD), sid(tok::TTS_PAD), sid(tok::TTS_PAD),
              sid(tok::TTS_PAD), sid(tok::TTS_PAD), sid(tok::TTS_BOS)};
  c.p.codec = {-1, -1, -1, tok::CODEC_THINK, tok::CODEC_THINK_BOS, (int32_t)lang, tok::CODEC_THINK_EOS, Q3_POS_SPEAKER, tok::CODEC_PAD};
  if (!vc.is_icl()) {
    if (!text_ids.empty()) { c.p.text.push_back(text_ids[0]); c.p.codec.push_back(tok::CODEC_BOS); }
    c.trailing.assign(text_ids.size() > 1 ? text_ids.begin() + 1 : text_ids.end(), text_ids.end());
    return c;
  }
  std::vector<int32_t> all_text = *vc.ref_text_ids;
  all_text.insert(all_text.end(), text_ids.begin(), text_ids.end());
  all_text.push_back(sid(tok::TTS_EOS));
  const int32_t n_text = (int32_t)all_text.size(), n_codec = (int32_t)vc.ref_codes->size() + 1;
  for (int32_t i = 0; i < n_codec; ++i) {
    c.p.text.push_back(i < n_text ? all_text[i] : sid(tok::TTS_PAD));
    c.p.codec.push_back(i == 0 ? (int32_t)tok::CODEC_BOS : Q3_POS_REF_FRAME(i - 1));
  }
  if (n_text > n_codec) c.trailing.assign(all_text.begin() + n_codec, all_text.end() - 1);
  else c.no_trailing = true;
  return c;
}

// ---- facade ---------------------------------------------------------------------------------------------------------------
class Qwen3TTS;

class StreamingSession {  // ref: src/lib.rs:1484-1782
 public:
  std::optional<AudioBuffer> next_chunk() {
    if (done_) return std::nullopt;
    FrameCodes c;
    std::vector<float> pcm;
    bool produced = sess_->stream_next(c, pcm, done_);
    frames_ += c.size();
    if (!produced) return std::nullopt;
    return AudioBuffer(std::move(pcm), 24000u);
  }
  size_t frames_generated() const { return frames_; }
  bool is_done() const { return done_; }

 private:
  friend class Qwen3TTS;
  explicit StreamingSession(std::unique_ptr<Session> s) : sess_(std::move(s)) {}
  std::unique_ptr<Session> sess_;
  size_t frames_ = 0;
  bool done_ = false;
};

class Qwen3TTS {
 public:
  /* ref: from_pretrained, lib.rs:183-262: config.json when present (a config that fails to parse falls back to weight
   * inspection, lib.rs:203-216 / 370-381), model.safetensors, speech_tokenizer/model.safetensors inside or beside the
   * directory; the reference's error texts. */
  static Qwen3TTS from_pretrained(const std::string& model_dir, int device = 0) {
    std::optional<ParsedModelConfig> cfg;
    if (file_exists(model_dir + "/config.json")) {
      try { cfg = ParsedModelConfig::from_file(model_dir + "/config.json"); } catch (const Error&) { cfg.reset(); }
    }
    const std::string model_path = model_dir + "/model.safetensors";
    if (!file_exists(model_path)) throw Error(Q3_ERR_INVALID, "Model weights not found at " + model_path + ". Please download the model first.");
    std::string st_dir = model_dir + "/speech_tokenizer";
    if (!file_exists(st_dir + "/model.safetensors")) {
      std::string d = model_dir;
      while (d.size() > 1 && d.back() == '/') d.pop_back();
      const size_t slash = d.find_last_of('/');
      st_dir = (slash == std::string::npos ? std::string(".") : d.substr(0, slash)) + "/speech_tokenizer";
      if (!file_exists(st_dir + "/model.safetensors")) throw Error(Q3_ERR_INVALID, "Speech tokenizer weights not found");
    }
    SafeTensorsFile weights(model_path), st_weights(st_dir + "/model.safetensors");
    q3_model_desc d;
    if (cfg) d = desc_from_config(*cfg, device);
    else {
      const TensorEntry* norm = weights.find("talker.model.norm.weight");
      if (!norm) throw Error(Q3_ERR_MISSING_WEIGHT, "Missing talker.model.norm.weight");
      d = default_desc(!norm->shape.empty() && norm->shape[0] == 2048, device);
    }
    if (file_exists(st_dir + "/config.json")) {
      std::ifstream f(st_dir + "/config.json");
      apply_vocoder_config(d, std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>()));
    }
    Qwen3TTS t;
    t.model_ = std::make_unique<Model>(d);
    t.model_->load_from(weights, "talker.");        // speaker_encoder.* stays in the file (voice-clone front end)
    t.model_->load_from(st_weights, "decoder.");    // encoder.* of the speech tokenizer likewise
    t.model_->finalize();
    if (cfg) t.model_type_ = cfg->model_type;
    return t;
  }

  std::optional<ModelType> model_type() const { return model_type_; }
  bool supports_preset_speakers() const { return !model_type_ || *model_type_ == ModelType::CustomVoice; }  // lib.rs:398-404
  bool supports_voice_design() const { return model_type_ && *model_type_ == ModelType::VoiceDesign; }     // lib.rs:409-411
  const Model& model() const { return *model_; }

  /* ref: generate_codes, lib.rs:530-656 (one utterance; `input_ids` are the text token ids). */
  FrameCodes generate_codes(const std::vector<int32_t>& input_ids, Speaker sp, Language lang, const SynthesisOptions& o) const {
    auto sess = new_session({input_ids}, {custom_voice_prompt(tv(), input_ids, sp, lang)}, o, {seed_of(o)}, o.max_length);
    return sess->generate(o.max_length)[0];
  }
  /* ref: synthesize_with_voice, lib.rs:718-784: prefill -> generate_codes -> decode_codes. */
  AudioBuffer synthesize_with_voice(const std::vector<int32_t>& input_ids, Speaker sp, Language lang, const SynthesisOptions& o,
                                    SynthesisTiming* timing = nullptr, FrameCodes* codes_out = nullptr) const {
    return synthesize({input_ids}, {custom_voice_prompt(tv(), input_ids, sp, lang)}, o, {seed_of(o)}, timing, codes_out)[0];
  }
  /* ref: synthesize_voice_design, lib.rs:802-870. */
  AudioBuffer synthesize_voice_design(const std::vector<int32_t>& input_ids, const std::vector<int32_t>& instruct_ids, Language lang,
                                      const SynthesisOptions& o, SynthesisTiming* timing = nullptr, FrameCodes* codes_out = nullptr) const {
    return synthesize({input_ids}, {voice_design_prompt(tv(), input_ids, instruct_ids, lang)}, o, {seed_of(o)}, timing, codes_out)[0];
  }
  /* Batched form (not in the reference, which has no batching): row i == an independent run with seeds[i]. */
  std::vector<AudioBuffer> synthesize_batch(const std::vector<std::vector<int32_t>>& batch_ids, Speaker sp, Language lang,
                                            const SynthesisOptions& o, const std::vector<uint64_t>& seeds,
                                            std::vector<FrameCodes>* codes_out = nullptr) const {
    std::vector<Prompt> prompts;
    for (const auto& ids : batch_ids) prompts.push_back(custom_voice_prompt(tv(), ids, sp, lang));
    std::vector<FrameCodes> codes;
    auto sess = new_session(batch_ids, prompts, o, seeds, o.max_length);
    codes = sess->generate(o.max_length);
    auto audio = sess->vocode(o.max_length);
    if (codes_out) *codes_out = std::move(codes);
    return audio;
  }
  /* ref: synthesize_streaming, lib.rs:1070-1093. */
  StreamingSession synthesize_streaming(const std::vector<int32_t>& input_ids, Speaker sp, Language lang, const SynthesisOptions& o) const {
    return StreamingSession(new_session({input_ids}, {custom_voice_prompt(tv(), input_ids, sp, lang)}, o, {seed_of(o)}, o.max_length));
  }
  /* ref: synthesize_voice_clone / synthesize_voice_clone_debug, lib.rs:895-1060: ICL adjustments of the generation config
     (913-927), voice-clone prefill + ICL block, frame loop, and in ICL mode the reference frames decoded in front of the
     generated ones with ref_len / total_len of the waveform cut from its start (1021-1040). */
  AudioBuffer synthesize_voice_clone(const std::vector<int32_t>& input_ids, const VoiceClonePrompt& vc, Language lang,
                                     const SynthesisOptions& options, FrameCodes* codes_out = nullptr) const {
    SynthesisOptions o = options;
    if ((int32_t)vc.speaker_embedding.size() != model_->desc().hidden) throw Error(Q3_ERR_INVALID, "speaker embedding must have `hidden` elements");
    if (vc.is_icl()) {
      o.repetition_penalty = std::max(o.repetition_penalty, ICL_MIN_REPETITION_PENALTY);
      o.max_length = std::min(o.max_length, std::max(ICL_MIN_FRAMES, (int32_t)input_ids.size() * ICL_FRAMES_PER_TOKEN));
    }
    const ClonePrompt cp = voice_clone_prompt(tv(), input_ids, vc, lang);
    const int32_t max_seq = std::max(o.max_length + 256, (int32_t)cp.p.text.size() + o.max_length);
    Session sess(*model_, 1, max_seq, o, {seed_of(o)});
    std::vector<uint16_t> spk(vc.speaker_embedding.size());
    for (size_t i = 0; i < spk.size(); ++i) spk[i] = f32_to_bf16_bits(vc.speaker_embedding[i]);
    std::vector<uint32_t> ref;
    if (vc.is_icl())
      for (const auto& fr : *vc.ref_codes) {
        if (fr.size() != 16) throw Error(Q3_ERR_INVALID, "reference frames are 16 codes wide");
        ref.insert(ref.end(), fr.begin(), fr.end());
      }
    sess.prefill_voice_clone(cp.p.text, cp.p.codec, spk, ref);
    const int32_t eos = special_text_id(tv(), tok::TTS_EOS), pad = special_text_id(tv(), tok::TTS_PAD);
    if (cp.no_trailing) sess.set_trailing_none(eos, pad);
    else sess.set_trailing_ids({cp.trailing}, eos, pad);
    FrameCodes codes = sess.generate(o.max_length)[0];
    if (codes_out) *codes_out = codes;
    if (!vc.is_icl()) return decode_codes(codes);
    FrameCodes combined = *vc.ref_codes;
    combined.insert(combined.end(), codes.begin(), codes.end());
    AudioBuffer audio = decode_codes(combined);
    const size_t cut = vc.ref_codes->size() * audio.samples.size() / std::max<size_t>(1, combined.size());
    audio.samples.erase(audio.samples.begin(), audio.samples.begin() + std::min(cut, audio.samples.size()));
    return audio;
  }
  /* ref: SpeakerEncoder::forward, src/models/speaker.rs:448-476 (has_speaker_encoder: lib.rs:389-391).  mel: f32 [mel_dim][t] of one
     utterance (the mel front end, src/audio/mel.rs, is the caller's) -> the VoiceClonePrompt's speaker embedding. */
  bool supports_voice_cloning() const { return q3_speaker_embed_dim(model_->handle()) > 0; }
  std::vector<float> speaker_encode(const std::vector<float>& mel, int32_t t) const {
    const int32_t dim = q3_speaker_embed_dim(model_->handle());
    if (dim <= 0) throw Error(Q3_ERR_STATE, "model has no speaker-encoder weights (speaker_encoder.*)");
    std::vector<float> out((size_t)dim);
    check(q3_speaker_encode(model_->handle(), mel.data(), 1, t, out.data()));
    return out;
  }
  /* ref: decode_codes, lib.rs:881-890. */
  AudioBuffer decode_codes(const FrameCodes& codes) const {
    const std::vector<int64_t> t = codes_to_tensor(codes);
    std::vector<float> pcm(codes.size() * (size_t)model_->total_upsample());
    if (!codes.empty()) check(q3_vocoder_decode(model_->handle(), t.data(), 1, (int32_t)codes.size(), pcm.data()));
    return AudioBuffer(std::move(pcm), 24000u);
  }

 private:
  Qwen3TTS() = default;
  int32_t tv() const { return model_->desc().text_vocab; }
  static uint64_t seed_of(const SynthesisOptions& o) {
    if (!o.seed) throw Error(Q3_ERR_INVALID, "a seed is required (the reference's unseeded mode is time-based and not reproducible)");
    return *o.seed;
  }
  std::unique_ptr<Session> new_session(const std::vector<std::vector<int32_t>>& batch_ids, const std::vector<Prompt>& prompts,
                                       const SynthesisOptions& o, const std::vector<uint64_t>& seeds, int32_t max_frames) const {
    int32_t lmax = 0;
    for (const auto& p : prompts) lmax = std::max<int32_t>(lmax, (int32_t)p.text.size());
    const int32_t max_seq = std::max(o.max_length + 256, lmax + max_frames);   // lib.rs:756
    auto s = std::make_unique<Session>(*model_, (int32_t)prompts.size(), max_seq, o, seeds);
    std::vector<std::vector<int32_t>> text, codec, trailing;
    for (const auto& p : prompts) { text.push_back(p.text); codec.push_back(p.codec); }
    for (const auto& ids : batch_ids) trailing.emplace_back(ids.size() > 1 ? ids.begin() + 1 : ids.end(), ids.end());
    s->prefill_ids(text, codec);
    s->set_trailing_ids(trailing, special_text_id(tv(), tok::TTS_EOS), special_text_id(tv(), tok::TTS_PAD));
    return s;
  }
  std::vector<AudioBuffer> synthesize(const std::vector<std::vector<int32_t>>& batch_ids, const std::vector<Prompt>& prompts,
                                      const SynthesisOptions& o, const std::vector<uint64_t>& seeds, SynthesisTiming* timing,
                                      FrameCodes* codes_out) const {
    auto sess = new_session(batch_ids, prompts, o, seeds, o.max_length);
    auto codes = sess->generate(o.max_length);
    auto audio = sess->vocode(o.max_length);
    if (timing) *timing = sess->timing();
    if (codes_out) *codes_out = codes[0];
    return audio;
  }

  std::unique_ptr<Model> model_;
  std::optional<ModelType> model_type_;
};

}  // namespace q3tts
#endif  // Q3TTS_HPP
