// TEST INFRASTRUCTURE ONLY -- a do-nothing stand-in for the inference library behind
// lib/beatricelib/beatrice.h, so that the REFERENCE's own call site (src/common/*.cc, compiled in place by
// oracle/Makefile) can be driven at millions of frames per second when only the call site's arithmetic is under
// test: the fp64 pitch transform of ProcessorCore2::Process1 (processor_core_2.cc:190-252).
//
//   EstimatePitch1 returns the bin set with Stub_SetNextPitch; GenerateWaveform1 records the bin the call site
//   hands it (Stub_LastPitch / Stub_WaveformCalls) and writes silence; everything else succeeds and does nothing.
//   ReadSpeakerEmbeddings fills the caller's tables with a fixed pseudo-random pattern (the call site normalises
//   them for its spherical averages; zeros would divide by zero).
//
//   Echo mode (environment STUB_ECHO=1, used by _ref/callsite_runner_stub): GenerateWaveform1 returns the 160 input
//   samples ExtractPhone1 was last given, followed by 80 zeros -- a "model" whose output is a known function of its
//   input, so that the call site's host-rate adapter (gain, resampler, block FIFO: gain.h, resample.h) can be compared
//   BIT FOR BIT with the device adapter of the product driven by the same stand-in (BeatriceB200_SetEchoModel).
//
// All 77 symbols of beatrice.h exist so that ProcessorProxy (cores 0, 1, 2) links.  C linkage: only names matter.
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace {
int g_next_q = 1, g_last_q = -1;
long g_wave_calls = 0;
float g_last_in[160] = {};
bool Echo() {
  static const bool on = [] {
    const char* ev = std::getenv("STUB_ECHO");
    return ev && ev[0] == '1';
  }();
  return on;
}
void Wave(float* out) {
  std::memset(out, 0, sizeof(float) * 240);
  if (Echo()) std::memcpy(out, g_last_in, sizeof(g_last_in));
}
constexpr int kSpeakers = 2;
void Fill(float* p, size_t n, uint32_t seed) {
  uint32_t s = seed * 2654435761u + 12345u;
  for (size_t i = 0; i < n; ++i) {
    s = s * 1664525u + 1013904223u;
    p[i] = (static_cast<float>(s >> 8) / 8388608.0f - 1.0f) * 0.5f;
  }
}
}  // namespace

extern "C" {
void Stub_SetNextPitch(int q) { g_next_q = q; }
int Stub_LastPitch(void) { return g_last_q; }
long Stub_WaveformCalls(void) { return g_wave_calls; }

#define STUB_COMMON(P, PHONE)                                                                       \
  void* P##_CreatePhoneExtractor(void) { return std::calloc(1, 8); }                                \
  void P##_DestroyPhoneExtractor(void* p) { std::free(p); }                                         \
  void* P##_CreatePhoneContext1(void) { return std::calloc(1, 8); }                                 \
  void P##_DestroyPhoneContext1(void* p) { std::free(p); }                                          \
  int P##_ReadPhoneExtractorParameters(void*, const char*) { return 0; }                            \
  void P##_ExtractPhone1(const void*, const float* in, float* out, void*) {                         \
    std::memcpy(g_last_in, in, sizeof(g_last_in));                                                  \
    std::memset(out, 0, sizeof(float) * PHONE);                                                     \
  }                                                                                                 \
  void* P##_CreatePitchEstimator(void) { return std::calloc(1, 8); }                                \
  void P##_DestroyPitchEstimator(void* p) { std::free(p); }                                         \
  void* P##_CreatePitchContext1(void) { return std::calloc(1, 8); }                                 \
  void P##_DestroyPitchContext1(void* p) { std::free(p); }                                          \
  int P##_ReadPitchEstimatorParameters(void*, const char*) { return 0; }                            \
  void P##_SetMinQuantizedPitch(void*, int) {}                                                      \
  void P##_SetMaxQuantizedPitch(void*, int) {}                                                      \
  void P##_EstimatePitch1(const void*, const float*, int* q, float* feat, void*) {                  \
    *q = g_next_q;                                                                                  \
    std::memset(feat, 0, sizeof(float) * 4);                                                        \
  }                                                                                                 \
  int P##_ReadNSpeakers(const char*, int* n) {                                                      \
    *n = kSpeakers;                                                                                 \
    return 0;                                                                                       \
  }                                                                                                 \
  void* P##_CreateWaveformGenerator(void) { return std::calloc(1, 8); }                             \
  void P##_DestroyWaveformGenerator(void* p) { std::free(p); }                                      \
  void* P##_CreateWaveformContext1(void) { return std::calloc(1, 8); }                              \
  void P##_DestroyWaveformContext1(void* p) { std::free(p); }                                       \
  int P##_ReadWaveformGeneratorParameters(void*, const char*) { return 0; }

STUB_COMMON(Beatrice20a2, 256)
STUB_COMMON(Beatrice20b1, 256)
STUB_COMMON(Beatrice20rc0, 128)

#define STUB_LEGACY(P)                                                                              \
  int P##_ReadSpeakerEmbeddings(const char*, float* table) {                                        \
    Fill(table, static_cast<size_t>(kSpeakers) * 256, 7);                                           \
    return 0;                                                                                       \
  }                                                                                                 \
  void P##_GenerateWaveform1(const void*, const float*, const int* q, const float*, const float*,   \
                             float* out, void*) {                                                   \
    g_last_q = *q;                                                                                  \
    ++g_wave_calls;                                                                                 \
    Wave(out);                                                                                      \
  }
STUB_LEGACY(Beatrice20a2)
STUB_LEGACY(Beatrice20b1)

void Beatrice20rc0_SetVQNumNeighbors(void*, int) {}
int Beatrice20rc0_ReadSpeakerEmbeddings(const char*, float* codebooks, float* additive, float* formant, float* kv) {
  Fill(codebooks, static_cast<size_t>(kSpeakers) * 512 * 128, 1);
  Fill(additive, static_cast<size_t>(kSpeakers) * 256, 2);
  Fill(formant, 9 * 256, 3);
  Fill(kv, static_cast<size_t>(kSpeakers) * 384 * 128, 4);
  return 0;
}
void Beatrice20rc0_GenerateWaveform1(const void*, const float*, const int* q, const float*, float* out, void*) {
  g_last_q = *q;
  ++g_wave_calls;
  Wave(out);
}
void* Beatrice20rc0_CreateEmbeddingSetter(void) { return std::calloc(1, 8); }
void Beatrice20rc0_DestroyEmbeddingSetter(void* p) { std::free(p); }
void* Beatrice20rc0_CreateEmbeddingContext(void) { return std::calloc(1, 8); }
void Beatrice20rc0_DestroyEmbeddingContext(void* p) { std::free(p); }
int Beatrice20rc0_ReadEmbeddingSetterParameters(void*, const char*) { return 0; }
void Beatrice20rc0_SetCodebook(void*, const float*) {}
void Beatrice20rc0_SetAdditiveSpeakerEmbedding(const void*, const float*, void*, void*) {}
void Beatrice20rc0_SetFormantShiftEmbedding(const void*, const float*, void*, void*) {}
void Beatrice20rc0_RegisterKeyValueSpeakerEmbedding(const void*, const float*, void*) {}
void Beatrice20rc0_SetKeyValueSpeakerEmbedding(const void*, int, void*, void*) {}
}  // extern "C"
