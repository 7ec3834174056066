// TEST INFRASTRUCTURE ONLY -- a thin C wrapper around the REFERENCE's own, unmodified
// call-site code so Python tests and the CPU baseline can drive it.
//
// This file is compiled TOGETHER WITH the reference sources where they lie
// (/root/reference/src/common/*.cc, header-only toml11) by oracle/Makefile; outputs go
// to oracle/_ref/ only.  Nothing of the reference is copied into this repository.
//
// What it exercises (all reference code):
//   ProcessorProxy::LoadModel / SetParameter      src/common/processor_proxy.h:41-100
//   ProcessorCore2::Process -> gain -> AnyFreqInOut -> Process1 -> gain
//                                                 src/common/processor_core_2.cc:24-256
//   resample.h:401-438, gain.h:41-71
// The Beatrice20{a2,b1,rc0}_* symbols it needs are left undefined here and are resolved
// at load time from whichever implementation the .so was linked against (the CPU oracle
// for libcallsite_oracle.so, the CUDA library for libcallsite_b200.so).

#include <array>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common/parameter_schema.h"
#include "common/processor_proxy.h"

using beatrice::common::ErrorCode;
using beatrice::common::kSchema;
using beatrice::common::ParameterID;
using beatrice::common::ProcessorProxy;

namespace {
struct Harness {
  ProcessorProxy proxy;
  explicit Harness(double sample_rate) : proxy(kSchema) { (void)proxy.SetSampleRate(sample_rate); }
};
std::u8string ToU8(const char* s) { return std::u8string(reinterpret_cast<const char8_t*>(s)); }
}  // namespace

extern "C" {

void* Callsite_Create(double sample_rate) { return new Harness(sample_rate); }
void Callsite_Destroy(void* h) { delete static_cast<Harness*>(h); }

// Same entry the VST uses for "param_change" of kModel (src/vst/processor.cc:270-298).
int Callsite_LoadModel(void* h, const char* toml_utf8) {
  return static_cast<int>(static_cast<Harness*>(h)->proxy.SetParameter(ParameterID::kModel, ToU8(toml_utf8)));
}
int Callsite_SetInt(void* h, int id, int value) {
  return static_cast<int>(static_cast<Harness*>(h)->proxy.SetParameter(static_cast<ParameterID>(id), value));
}
int Callsite_SetDouble(void* h, int id, double value) {
  return static_cast<int>(static_cast<Harness*>(h)->proxy.SetParameter(static_cast<ParameterID>(id), value));
}
int Callsite_SetSampleRate(void* h, double sr) {
  return static_cast<int>(static_cast<Harness*>(h)->proxy.SetSampleRate(sr));
}
int Callsite_GetVersion(void* h) { return static_cast<Harness*>(h)->proxy.GetCore()->GetVersion(); }
int Callsite_ResetContext(void* h) {
  return static_cast<int>(static_cast<Harness*>(h)->proxy.GetCore()->ResetContext());
}
// ProcessorCore2::SetSpeakerMorphingWeights (processor_core_2.cc:498-505), the call the voice-morph parameters end in
// (parameter_schema.cc:36-40).  weights: 256 floats (std::array<float, kMaxNSpeakers>).
int Callsite_SetMorphWeights(void* h, const float* weights256) {
  std::array<float, beatrice::common::kMaxNSpeakers> w{};
  std::memcpy(w.data(), weights256, sizeof(float) * w.size());
  return static_cast<int>(static_cast<Harness*>(h)->proxy.GetCore()->SetSpeakerMorphingWeights(w));
}
// In-place allowed, like Processor::process does (src/vst/processor.cc:216-217).
int Callsite_Process(void* h, const float* in, float* out, int n) {
  return static_cast<int>(static_cast<Harness*>(h)->proxy.GetCore()->Process(in, out, n));
}

// CPU baseline: `n_threads` independent streams, one thread each, every thread pushes
// `n_frames` 480-sample blocks of its own slice of `signal` (n_threads x n_frames x 480)
// through ProcessorCore2::Process after `warmup` untimed blocks.  Returns frames/s summed
// over threads (0 on load failure).  Per-thread seconds go to `seconds_out` if non-null.
double Callsite_Bench(const char* toml_utf8, double sample_rate, int n_threads, int n_frames, int warmup,
                      const float* signal, double* seconds_out) {
  std::vector<std::unique_ptr<Harness>> hs;
  for (int i = 0; i < n_threads; ++i) {
    hs.emplace_back(new Harness(sample_rate));
    if (hs.back()->proxy.SetParameter(ParameterID::kModel, ToU8(toml_utf8)) != ErrorCode::kSuccess) return 0.0;
  }
  std::vector<double> secs(n_threads, 0.0);
  std::vector<std::thread> th;
  for (int i = 0; i < n_threads; ++i) {
    th.emplace_back([&, i] {
      std::vector<float> out(480);
      const float* x = signal + static_cast<size_t>(i) * n_frames * 480;
      auto& core = hs[i]->proxy.GetCore();
      for (int f = 0; f < warmup; ++f) (void)core->Process(x + static_cast<size_t>(f % n_frames) * 480, out.data(), 480);
      const auto t0 = std::chrono::steady_clock::now();
      for (int f = 0; f < n_frames; ++f) (void)core->Process(x + static_cast<size_t>(f) * 480, out.data(), 480);
      secs[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
  }
  for (auto& t : th) t.join();
  double worst = 0.0;
  for (int i = 0; i < n_threads; ++i) {
    worst = std::max(worst, secs[i]);
    if (seconds_out) seconds_out[i] = secs[i];
  }
  return worst > 0.0 ? static_cast<double>(n_threads) * n_frames / worst : 0.0;
}

}  // extern "C"
