// TEST INFRASTRUCTURE ONLY -- command-line driver around callsite_harness.cc (the
// reference's unmodified src/common call site).  Runs in its own process so that the
// reference's C++ code never shares a symbol namespace with Python extension modules.
//
//   callsite_runner run   <model.toml|-> <in.f32> <out.f32> <rate> <block> [idx:name=value ...]
//       Feeds in.f32 through ProcessorCore::Process in `block`-sample calls (in place, like
//       src/vst/processor.cc:216-217).  "idx:name=value" applies SetParameter before block
//       `idx` (idx = -1: before LoadModel; names as in parameter_schema.h:44-70, see table).
//       "idx:reset=1" calls ResetContext().  "idx:morphw<k>=w" + "idx:morph_apply=1": voice-morphing weights.  Prints one line: "error_codes: load=<e> last=<e>".
//   callsite_runner bench <model.toml> <signal.f32> <threads> <frames> <warmup>
//       signal.f32 holds threads*frames*480 floats.  Prints JSON with frames/s.

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern "C" {
void* Callsite_Create(double sample_rate);
void Callsite_Destroy(void* h);
int Callsite_LoadModel(void* h, const char* toml_utf8);
int Callsite_SetInt(void* h, int id, int value);
int Callsite_SetDouble(void* h, int id, double value);
int Callsite_GetVersion(void* h);
int Callsite_ResetContext(void* h);
int Callsite_Process(void* h, const float* in, float* out, int n);
int Callsite_SetMorphWeights(void* h, const float* weights256);
double Callsite_Bench(const char* toml_utf8, double sample_rate, int n_threads, int n_frames, int warmup,
                      const float* signal, double* seconds_out);
}

namespace {
struct Param {
  const char* name;
  int id;
  bool is_int;
};
// ids: reference src/common/parameter_schema.h:44-70
const Param kParams[] = {{"voice", 2, true},
                         {"formant_shift", 3, false},
                         {"pitch_shift", 4, false},
                         {"average_source_pitch", 5, false},
                         {"lock", 6, true},
                         {"input_gain", 7, false},
                         {"output_gain", 8, false},
                         {"intonation_intensity", 9, false},
                         {"pitch_correction", 10, false},
                         {"pitch_correction_type", 11, true},
                         {"min_source_pitch", 12, false},
                         {"max_source_pitch", 13, false},
                         {"vq_num_neighbors", 14, false}};

struct Event {
  long block;
  std::string name;
  double value;
};

std::vector<float> ReadF32(const char* path) {
  std::vector<float> v;
  FILE* f = std::fopen(path, "rb");
  if (!f) return v;
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f) / 4;
  std::fseek(f, 0, SEEK_SET);
  v.resize(n);
  if (n > 0 && std::fread(v.data(), 4, n, f) != static_cast<size_t>(n)) v.clear();
  std::fclose(f);
  return v;
}

// "morphw<k>=v" stages the morphing weight of speaker k; "morph_apply=1" hands the staged array to
// ProcessorCore2::SetSpeakerMorphingWeights.  ("voice=<n_speakers>" selects the morphing slot.)
float g_morph_weights[256] = {};

int Apply(void* h, const Event& e) {
  if (e.name == "reset") return Callsite_ResetContext(h);
  if (e.name.rfind("morphw", 0) == 0) {
    const int k = std::atoi(e.name.c_str() + 6);
    if (k < 0 || k >= 256) return -1;
    g_morph_weights[k] = static_cast<float>(e.value);
    return 0;
  }
  if (e.name == "morph_apply") return Callsite_SetMorphWeights(h, g_morph_weights);
  for (const Param& p : kParams)
    if (e.name == p.name)
      return p.is_int ? Callsite_SetInt(h, p.id, static_cast<int>(e.value)) : Callsite_SetDouble(h, p.id, e.value);
  std::fprintf(stderr, "unknown parameter %s\n", e.name.c_str());
  return -1;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc >= 7 && std::strcmp(argv[1], "run") == 0) {
    const char* toml = argv[2];
    std::vector<float> x = ReadF32(argv[3]);
    const double rate = std::atof(argv[5]);
    const int block = std::atoi(argv[6]);
    std::vector<Event> events;
    for (int i = 7; i < argc; ++i) {
      std::string s = argv[i];
      const size_t c = s.find(':'), q = s.find('=');
      if (c == std::string::npos || q == std::string::npos) return 2;
      events.push_back({std::atol(s.substr(0, c).c_str()), s.substr(c + 1, q - c - 1), std::atof(s.substr(q + 1).c_str())});
    }
    void* h = Callsite_Create(rate);
    for (const Event& e : events)
      if (e.block < 0) Apply(h, e);
    int load_err = -1;
    if (std::strcmp(toml, "-") != 0) load_err = Callsite_LoadModel(h, toml);
    int last = 0;
    long bi = 0;
    for (size_t i = 0; i < x.size(); i += block, ++bi) {
      for (const Event& e : events)
        if (e.block == bi) Apply(h, e);
      const int n = static_cast<int>(std::min<size_t>(block, x.size() - i));
      last = Callsite_Process(h, x.data() + i, x.data() + i, n);
    }
    FILE* f = std::fopen(argv[4], "wb");
    if (!f) return 3;
    std::fwrite(x.data(), 4, x.size(), f);
    std::fclose(f);
    std::printf("error_codes: load=%d last=%d version=%d\n", load_err, last, Callsite_GetVersion(h));
    Callsite_Destroy(h);
    return 0;
  }
  if (argc >= 7 && std::strcmp(argv[1], "bench") == 0) {
    std::vector<float> sig = ReadF32(argv[3]);
    const int threads = std::atoi(argv[4]), frames = std::atoi(argv[5]), warmup = std::atoi(argv[6]);
    if (sig.size() < static_cast<size_t>(threads) * frames * 480) return 2;
    std::vector<double> secs(threads);
    const double fps = Callsite_Bench(argv[2], 48000.0, threads, frames, warmup, sig.data(), secs.data());
    double worst = 0.0;
    for (double s : secs) worst = s > worst ? s : worst;
    std::printf("{\"frames_per_s\": %.3f, \"threads\": %d, \"frames_per_thread\": %d, \"seconds\": %.6f}\n", fps,
                threads, frames, worst);
    return fps > 0.0 ? 0 : 4;
  }
  std::fprintf(stderr, "usage: see header of callsite_runner.cc\n");
  return 2;
}
