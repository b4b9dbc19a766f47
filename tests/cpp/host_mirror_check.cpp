// Driver for tests/test_cpp_host.py: exercises include/q3tts.hpp (the C++ host mirror) and prints machine-readable
// lines that the test compares with the Python mirror (qwen3_tts_rs_b200/api.py, formats.py).
#include <cinttypes>
#include <functional>
#include <iostream>
#include <sstream>

#include "q3tts.hpp"

using namespace q3tts;

static std::vector<int32_t> parse_ids(const std::string& s) {
  std::vector<int32_t> out;
  std::stringstream ss(s);
  std::string t;
  while (std::getline(ss, t, ','))
    if (!t.empty()) out.push_back((int32_t)std::stol(t));
  return out;
}
static void print_ids(const char* tag, const std::vector<int32_t>& v) {
  std::printf("%s", tag);
  for (int32_t x : v) std::printf(" %d", x);
  std::printf("\n");
}
static uint64_t fnv1a(const uint8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char** argv) {
  try {
    const std::string mode = argc > 1 ? argv[1] : "";
    if (mode == "prompts") {  // prompts <text_vocab> <text ids> <instruct ids>
      const int32_t tv = (int32_t)std::stol(argv[2]);
      const auto ids = parse_ids(argv[3]);
      const auto inst = parse_ids(argc > 4 ? argv[4] : "");
      Prompt a = custom_voice_prompt(tv, ids, *speaker_from_name("Ryan"), *language_from_name("english"));
      Prompt b = voice_design_prompt(tv, ids, inst, Language::Japanese);
      print_ids("cv_text", a.text); print_ids("cv_codec", a.codec);
      print_ids("vd_text", b.text); print_ids("vd_codec", b.codec);
      std::printf("speaker_unknown %d\n", (int)speaker_from_name("nobody").has_value());
      return 0;
    }
    if (mode == "clone_prompt") {  // clone_prompt <text_vocab> <text ids> <ref text ids | -> <ref frames>: voice-clone id layout
      const int32_t tv = (int32_t)std::stol(argv[2]);
      const auto ids = parse_ids(argv[3]);
      VoiceClonePrompt vc;
      vc.speaker_embedding.assign(8, 0.f);
      const int32_t t_ref = (int32_t)std::stol(argv[5]);
      if (std::string(argv[4]) != "-") {
        vc.ref_text_ids = parse_ids(argv[4]);
        vc.ref_codes = FrameCodes((size_t)t_ref, std::vector<uint32_t>(16, 1u));
      }
      const ClonePrompt c = voice_clone_prompt(tv, ids, vc, *language_from_name("english"));
      print_ids("text", c.p.text); print_ids("codec", c.p.codec); print_ids("trailing", c.trailing);
      std::printf("no_trailing %d\n", (int)c.no_trailing);
      std::printf("bf16 %04x %04x %04x\n", f32_to_bf16_bits(1.0f), f32_to_bf16_bits(0.3f), f32_to_bf16_bits(-2.0078125f));
      return 0;
    }
    if (mode == "formats") {  // formats <dir>: write the dump / WAV formats from deterministic data, read Python's files back
      const std::string d = argv[2];
      FrameCodes codes;
      for (uint32_t f = 0; f < 5; ++f) {
        codes.emplace_back();
        for (uint32_t q = 0; q < 16; ++q) codes.back().push_back((f * 131u + q * 17u) % 3072u);
      }
      std::vector<float> audio(3000);
      for (size_t i = 0; i < audio.size(); ++i) audio[i] = 1.3f * std::sin(0.01f * (float)i) * ((i % 7) ? 1.f : -1.f);
      save_codes_binary(codes, d + "/cpp_codes.bin");
      save_audio_binary(audio, d + "/cpp_audio.bin");
      AudioBuffer(audio, 24000).save(d + "/cpp.wav");
      auto t = codes_to_tensor(codes);
      std::printf("tensor");
      for (int64_t x : t) std::printf(" %" PRId64, x);
      std::printf("\n");
      if (file_exists(d + "/py.wav")) {
        AudioBuffer w = AudioBuffer::load(d + "/py.wav");
        std::printf("py_wav %u %zu %016" PRIx64 "\n", w.sample_rate, w.len(), fnv1a((const uint8_t*)w.samples.data(), w.len() * 4));
      }
      if (file_exists(d + "/py_codes.bin")) {
        FrameCodes c = load_codes_binary(d + "/py_codes.bin");
        std::printf("py_codes_equal %d\n", (int)(c == codes));
      }
      if (file_exists(d + "/codes_seed7_frames5.bin")) {
        audio[10] += 0.5f;
        codes[2][3] += 1;
        CompareReport r = compare_with_reference(d, 7, 5, codes, audio);
        std::printf("compare %d %d %zu %zu %zu %.9g %.9g %.9g\n", (int)r.codes_found, (int)r.codes_match, r.n_code_diffs, r.n_ref_codes,
                    r.n_audio_compared, (double)r.max_diff, r.mean_diff, r.rmse);
      }
      AudioBuffer n(std::vector<float>{0.5f, -0.25f, 0.1f}, 24000);
      n.normalize();
      std::printf("normalize %.9g %.9g %.9g\n", n.samples[0], n.samples[1], n.samples[2]);
      return 0;
    }
    if (mode == "config") {  // config <config.json>
      ParsedModelConfig c = ParsedModelConfig::from_file(argv[2]);
      std::printf("label %s\n", c.label().c_str());
      std::printf("talker %zu %zu %zu %zu %zu %zu %zu %zu %zu %.9g %.9g %zu\n", c.talker_hidden_size, c.talker_intermediate_size,
                  c.talker_num_hidden_layers, c.talker_num_attention_heads, c.talker_num_key_value_heads, c.talker_head_dim,
                  c.talker_vocab_size, c.talker_text_vocab_size, c.talker_text_hidden_size, c.talker_rms_norm_eps, c.talker_rope_theta,
                  c.talker_max_position_embeddings);
      std::printf("cp %zu %zu %zu %zu %zu %zu %zu %zu %.9g %.9g\n", c.cp_hidden_size, c.cp_intermediate_size, c.cp_num_hidden_layers,
                  c.cp_num_attention_heads, c.cp_num_key_value_heads, c.cp_head_dim, c.cp_vocab_size, c.cp_num_code_groups,
                  c.cp_rms_norm_eps, c.cp_rope_theta);
      if (c.mrope_section) std::printf("mrope %zu %zu %zu\n", (*c.mrope_section)[0], (*c.mrope_section)[1], (*c.mrope_section)[2]);
      else std::printf("mrope none\n");
      std::printf("speaker_enc_dim %ld\n", c.speaker_enc_dim ? (long)*c.speaker_enc_dim : -1L);
      return 0;
    }
    if (mode == "wav") {  // wav <file>: rate, sample count, hash of the f32 samples
      AudioBuffer w = AudioBuffer::load(argv[2]);
      std::printf("%u %zu %016" PRIx64 "\n", w.sample_rate, w.len(), fnv1a((const uint8_t*)w.samples.data(), w.len() * 4));
      return 0;
    }
    if (mode == "json") {  // json <file>: parse and print a canonical one-line form (strings as hex, numbers as int / %.17g)
      auto raw = read_file(argv[2]);
      const json::Value v = json::parse(std::string(raw.begin(), raw.end()));
      std::function<void(const json::Value&)> dump = [&](const json::Value& x) {
        switch (x.kind) {
          case json::Value::Null: std::printf("n"); break;
          case json::Value::Bool: std::printf(x.b ? "t" : "f"); break;
          case json::Value::Int: std::printf("i%" PRId64, x.i); break;
          case json::Value::Float: std::printf("d%.17g", x.d); break;
          case json::Value::String:
            std::printf("s");
            for (unsigned char c : x.s) std::printf("%02x", c);
            break;
          case json::Value::Array:
            std::printf("[");
            for (const auto& e : x.a) { dump(e); std::printf(","); }
            std::printf("]");
            break;
          case json::Value::Obj:
            std::printf("{");
            for (const auto& kv : x.o) {
              std::printf("s");
              for (unsigned char c : kv.first) std::printf("%02x", c);
              std::printf(":");
              dump(kv.second);
              std::printf(",");
            }
            std::printf("}");
            break;
        }
      };
      dump(v);
      std::printf("\n");
      return 0;
    }
    if (mode == "safetensors") {  // safetensors <file>
      SafeTensorsFile f(argv[2]);
      for (const auto& kv : f.entries()) {
        std::printf("%s %s [", kv.first.c_str(), kv.second.dtype.c_str());
        for (size_t i = 0; i < kv.second.shape.size(); ++i) std::printf(i ? ",%" PRId64 : "%" PRId64, kv.second.shape[i]);
        std::printf("] %016" PRIx64 "\n", fnv1a(f.data(kv.second), kv.second.end - kv.second.begin));
      }
      return 0;
    }
    if (mode == "generate") {  // generate <model_dir> <ids> <seed> <frames> <out_dir> [instruct ids]
      Qwen3TTS tts = Qwen3TTS::from_pretrained(argv[2]);
      const auto ids = parse_ids(argv[3]);
      SynthesisOptions o;
      o.seed = (uint64_t)std::stoull(argv[4]);
      o.max_length = (int32_t)std::stol(argv[5]);
      const std::string out = argv[6];
      FrameCodes codes;
      SynthesisTiming tm;
      AudioBuffer a = argc > 7 ? tts.synthesize_voice_design(ids, parse_ids(argv[7]), Language::English, o, &tm, &codes)
                               : tts.synthesize_with_voice(ids, Speaker::Ryan, Language::English, o, &tm, &codes);
      save_codes_binary(codes, out + "/codes_seed" + argv[4] + "_frames" + std::to_string(codes.size()) + ".bin");
      save_audio_binary(a.samples, out + "/audio_seed" + argv[4] + "_frames" + std::to_string(codes.size()) + ".bin");
      a.save(out + "/audio.wav");
      // the same utterance through the other entry points
      FrameCodes again = tts.generate_codes(ids, Speaker::Ryan, Language::English, o);
      AudioBuffer dec = tts.decode_codes(codes);
      o.chunk_frames = 3;
      StreamingSession st = tts.synthesize_streaming(ids, Speaker::Ryan, Language::English, o);
      size_t streamed = 0, chunks = 0;
      while (auto c = st.next_chunk()) { streamed += c->len(); ++chunks; }
      float dec_diff = dec.len() == a.len() ? 0.f : 1e9f;
      for (size_t i = 0; i < std::min(dec.len(), a.len()); ++i) dec_diff = std::max(dec_diff, std::fabs(dec.samples[i] - a.samples[i]));
      std::printf("frames %zu samples %zu generate_codes_equal %d decode_codes_maxdiff %.9g streamed_samples %zu chunks %zu stream_frames %zu "
                  "model_type %d launches %" PRIu64 "\n",
                  codes.size(), a.len(), (int)(argc > 7 ? 1 : again == codes), (double)dec_diff, streamed, chunks,
                  st.frames_generated(), tts.model_type() ? (int)*tts.model_type() : -1, q3_kernel_launch_count());
      return 0;
    }
    std::fprintf(stderr, "usage: host_mirror_check prompts|formats|config|safetensors|generate ...\n");
    return 64;
  } catch (const Error& e) {
    std::fprintf(stderr, "q3tts::Error %d: %s\n", (int)e.code, e.what());
    return 10 + (int)e.code;
  }
}
