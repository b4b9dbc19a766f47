/* TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 * Plain-C restatement of the integer / byte-exact parts of the hot path:
 *   PCG-XSH-RR 64/32 + seed mixing      src/generation/sampling.rs:32-51, 84-94
 *   suppression rule                    src/generation/tts.rs:21-43
 *   codes_to_tensor layout              src/lib.rs:1417-1431
 *   codes dump (i64 LE, frame-major)    src/bin/generate_audio.rs:788-801
 *   PCM16 conversion for WAV            src/audio/io.rs:143-165  ((clamp(x) * 32767) as i16)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

uint64_t q3o_seed_state(uint64_t seed) { return seed * 2685821657736338717ULL + 1442695040888963407ULL; }

uint32_t q3o_pcg_next(uint64_t* state) {
  uint64_t old = *state;
  *state = old * 6364136223846793005ULL + 1442695040888963407ULL;
  uint32_t xorshifted = (uint32_t)(((old >> 18) ^ old) >> 27);
  uint32_t rot = (uint32_t)(old >> 59);
  return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
}

float q3o_rand_f32(uint64_t* state) { return (float)q3o_pcg_next(state) / (float)UINT32_MAX; }

void q3o_suppression_mask(int vocab, int eos, uint8_t* mask) {
  memset(mask, 0, (size_t)vocab);
  for (int v = vocab - 1024; v < vocab; ++v)
    if (v != eos) mask[v] = 1;
}

/* codes: [n_frames][16] u32 -> out: [16][n_frames] i64 */
void q3o_codes_to_tensor(const uint32_t* codes, int n_frames, int64_t* out) {
  for (int f = 0; f < n_frames; ++f)
    for (int q = 0; q < 16; ++q) out[(size_t)q * n_frames + f] = (int64_t)codes[(size_t)f * 16 + q];
}

/* frame-major i64 little-endian dump, as generate_audio writes it */
void q3o_codes_dump(const uint32_t* codes, int n_frames, uint8_t* out) {
  for (size_t i = 0; i < (size_t)n_frames * 16; ++i) {
    uint64_t v = (uint64_t)(int64_t)codes[i];
    for (int b = 0; b < 8; ++b) out[i * 8 + b] = (uint8_t)(v >> (8 * b));
  }
}

void q3o_pcm16(const float* x, int n, int16_t* out) {
  for (int i = 0; i < n; ++i) {
    float c = x[i] < -1.0f ? -1.0f : (x[i] > 1.0f ? 1.0f : x[i]);
    out[i] = (int16_t)(c * 32767.0f);
  }
}
