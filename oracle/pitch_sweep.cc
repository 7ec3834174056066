// TEST INFRASTRUCTURE ONLY -- exhaustive known answers for the call-site pitch transform.
//
//   pitch_sweep <model.toml> <params.f64> <out.i32>
//
// params.f64 holds rows of 5 doubles (average_source_pitch, intonation_intensity, pitch_shift, pitch_correction,
// pitch_correction_type).  For every row the parameters are set through the reference's own ProcessorProxy
// (callsite_harness.cc) and ProcessorCore2::Process is called once per raw bin q = 1 .. 447 (the stub library,
// stub_beatricelib.cc, makes EstimatePitch1 return q); the bin the call site then hands to GenerateWaveform1 --
// the result of processor_core_2.cc:190-252 as compiled from /root/reference -- is written to out.i32 as
// [rows][447] int32.  The model behind the call site is a stub: only the call site's arithmetic runs.
#include <cstdio>
#include <cstdlib>
#include <vector>

extern "C" {
void* Callsite_Create(double sample_rate);
void Callsite_Destroy(void* h);
int Callsite_LoadModel(void* h, const char* toml_utf8);
int Callsite_SetInt(void* h, int id, int value);
int Callsite_SetDouble(void* h, int id, double value);
int Callsite_Process(void* h, const float* in, float* out, int n);
void Stub_SetNextPitch(int q);
int Stub_LastPitch(void);
long Stub_WaveformCalls(void);
}

int main(int argc, char** argv) {
  if (argc != 4) {
    std::fprintf(stderr, "usage: pitch_sweep <model.toml> <params.f64> <out.i32>\n");
    return 2;
  }
  std::vector<double> params;
  {
    FILE* f = std::fopen(argv[2], "rb");
    if (!f) return 3;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f) / 8;
    std::fseek(f, 0, SEEK_SET);
    params.resize(n);
    if (n > 0 && std::fread(params.data(), 8, n, f) != static_cast<size_t>(n)) return 3;
    std::fclose(f);
  }
  const size_t rows = params.size() / 5;
  void* h = Callsite_Create(48000.0);
  if (Callsite_LoadModel(h, argv[1]) != 0) {
    std::fprintf(stderr, "pitch_sweep: LoadModel failed\n");
    return 4;
  }
  // ids: reference src/common/parameter_schema.h:44-70 (the table in callsite_runner.cc)
  constexpr int kPitchShift = 4, kAverageSourcePitch = 5, kIntonation = 9, kCorrection = 10, kCorrectionType = 11;
  std::vector<float> x(480, 0.25f), y(480);
  std::vector<int> out(rows * 447);
  for (size_t r = 0; r < rows; ++r) {
    const double* p = &params[r * 5];
    Callsite_SetDouble(h, kAverageSourcePitch, p[0]);
    Callsite_SetDouble(h, kIntonation, p[1]);
    Callsite_SetDouble(h, kPitchShift, p[2]);
    Callsite_SetDouble(h, kCorrection, p[3]);
    Callsite_SetInt(h, kCorrectionType, static_cast<int>(p[4]));
    for (int q = 1; q <= 447; ++q) {
      Stub_SetNextPitch(q);
      const long before = Stub_WaveformCalls();
      if (Callsite_Process(h, x.data(), y.data(), 480) != 0) return 5;
      if (Stub_WaveformCalls() != before + 1) return 6;   // one Process1 per 480-sample block at 48 kHz
      out[r * 447 + (q - 1)] = Stub_LastPitch();
    }
  }
  Callsite_Destroy(h);
  FILE* f = std::fopen(argv[3], "wb");
  if (!f) return 7;
  std::fwrite(out.data(), 4, out.size(), f);
  std::fclose(f);
  std::printf("rows=%zu\n", rows);
  return 0;
}
