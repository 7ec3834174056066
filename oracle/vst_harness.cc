// TEST INFRASTRUCTURE ONLY -- headless VST3 host around the reference's UNMODIFIED plug-in processor
// (reference src/vst/processor.cc, class beatrice::vst::Processor), SURVEY.md section 8 row (f-2).
//
// oracle/Makefile compiles src/vst/processor.cc, src/common/*.cc and the vendored vst3sdk sources
// (lib/vst3sdk: base, pluginterfaces, public.sdk without VSTGUI) where they lie under /root/reference
// and links them with this file against either the CPU oracle or the CUDA product library.  This file
// plays the host: it drives the processor exactly through the interfaces a DAW uses --
//   IComponent::initialize / setBusArrangements / setupProcessing / activateBus / setActive,
//   IConnectionPoint::notify("param_change")      -> model load   (processor.cc:274-300)
//   IAudioProcessor::process(ProcessData)          -> parameter queues + audio (processor.cc:103-231)
// so that a replacement library is proven behind the whole untouched VST path, not only behind
// src/common.
//
//   vst_harness run <model.toml|-> <in.f32> <out.f32> <rate> <block> [idx:name=value ...]
//
// Same command line as callsite_runner.cc.  "idx:name=value" queues the parameter (normalised the
// way the plug-in's controller does, src/vst/parameter.h:20-38) into the IParameterChanges of block
// idx (idx = -1: a zero-sample process call before the model-load message); "idx:reset=1" is
// setActive(false) + setActive(true) (processor.cc:88-98 -> ResetContext).  Prints
//   applied <name>=<plain value the processor de-normalised>     one line per event
//   status: load=<tresult> process=<last tresult>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <variant>
#include <vector>

#include "common/parameter_schema.h"
#include "vst/parameter.h"
#include "vst/processor.h"
#include "vst3sdk/pluginterfaces/vst/ivstaudioprocessor.h"
#include "vst3sdk/pluginterfaces/vst/ivstcomponent.h"
#include "vst3sdk/pluginterfaces/vst/ivstmessage.h"
#include "vst3sdk/pluginterfaces/vst/ivstprocesscontext.h"
#include "vst3sdk/pluginterfaces/vst/vstspeaker.h"
#include "vst3sdk/public.sdk/source/vst/hosting/hostclasses.h"
#include "vst3sdk/public.sdk/source/vst/hosting/parameterchanges.h"
#include "vst3sdk/public.sdk/source/vst/hosting/processdata.h"

namespace {
namespace sv = Steinberg::Vst;
using Steinberg::FUnknownPtr;
using Steinberg::IPtr;
using Steinberg::kResultOk;
using Steinberg::kResultTrue;
using Steinberg::tresult;

struct Named {
  const char* name;
  beatrice::common::ParameterID id;
};
// names as in callsite_runner.cc; ids: reference src/common/parameter_schema.h:44-70
const Named kNames[] = {{"voice", beatrice::common::ParameterID::kVoice},
                        {"formant_shift", beatrice::common::ParameterID::kFormantShift},
                        {"pitch_shift", beatrice::common::ParameterID::kPitchShift},
                        {"average_source_pitch", beatrice::common::ParameterID::kAverageSourcePitch},
                        {"lock", beatrice::common::ParameterID::kLock},
                        {"input_gain", beatrice::common::ParameterID::kInputGain},
                        {"output_gain", beatrice::common::ParameterID::kOutputGain},
                        {"intonation_intensity", beatrice::common::ParameterID::kIntonationIntensity},
                        {"pitch_correction", beatrice::common::ParameterID::kPitchCorrection},
                        {"pitch_correction_type", beatrice::common::ParameterID::kPitchCorrectionType},
                        {"min_source_pitch", beatrice::common::ParameterID::kMinSourcePitch},
                        {"max_source_pitch", beatrice::common::ParameterID::kMaxSourcePitch},
                        {"vq_num_neighbors", beatrice::common::ParameterID::kVQNumNeighbors}};

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

// Queues one parameter point the way a host forwards a controller edit; reports the plain value the
// processor will see after its own Denormalize (processor.cc:141-158).
bool Queue(sv::ParameterChanges* changes, const Event& e) {
  for (const Named& n : kNames) {
    if (e.name != n.name) continue;
    const auto& param = beatrice::common::kSchema.GetParameter(n.id);
    double normalized = 0.0, plain = 0.0;
    if (const auto* num = std::get_if<beatrice::common::NumberParameter>(&param)) {
      normalized = beatrice::vst::Normalize(*num, e.value);
      plain = beatrice::vst::Denormalize(*num, normalized);
    } else if (const auto* list = std::get_if<beatrice::common::ListParameter>(&param)) {
      normalized = beatrice::vst::Normalize(*list, static_cast<int>(e.value));
      plain = beatrice::vst::Denormalize(*list, normalized);
    } else {
      return false;
    }
    Steinberg::int32 index = 0;
    sv::IParamValueQueue* q = changes->addParameterData(static_cast<sv::ParamID>(n.id), index);
    if (!q) return false;
    Steinberg::int32 point = 0;
    q->addPoint(0, normalized, point);
    std::printf("applied %s=%.17g\n", n.name, plain);
    return true;
  }
  std::fprintf(stderr, "unknown parameter %s\n", e.name.c_str());
  return false;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 7 || std::strcmp(argv[1], "run") != 0) {
    std::fprintf(stderr, "usage: see header of vst_harness.cc\n");
    return 2;
  }
  const char* toml = argv[2];
  std::vector<float> x = ReadF32(argv[3]);
  const double rate = std::atof(argv[5]);
  const int block = std::atoi(argv[6]);
  std::vector<Event> events;
  for (int i = 7; i < argc; ++i) {
    const std::string s = argv[i];
    const size_t c = s.find(':'), q = s.find('=');
    if (c == std::string::npos || q == std::string::npos) return 2;
    events.push_back({std::atol(s.substr(0, c).c_str()), s.substr(c + 1, q - c - 1), std::atof(s.substr(q + 1).c_str())});
  }

  // ---- instantiate and set up, in the order of the VST3 component life cycle ----
  sv::HostApplication host;
  IPtr<sv::IAudioProcessor> processor =
      Steinberg::owned(static_cast<sv::IAudioProcessor*>(beatrice::vst::Processor::createInstance(nullptr)));
  FUnknownPtr<sv::IComponent> component(processor);
  FUnknownPtr<sv::IConnectionPoint> connection(processor);
  if (!component || !connection) return 3;
  if (component->initialize(&host) != kResultTrue) return 3;
  sv::SpeakerArrangement in_arr = sv::SpeakerArr::kMono, out_arr = sv::SpeakerArr::kMono;
  if (processor->setBusArrangements(&in_arr, 1, &out_arr, 1) != kResultTrue) return 3;
  sv::ProcessSetup setup{sv::kRealtime, sv::kSample32, block, rate};
  if (processor->setupProcessing(setup) != kResultOk) return 3;
  component->activateBus(sv::kAudio, sv::kInput, 0, true);
  component->activateBus(sv::kAudio, sv::kOutput, 0, true);
  if (component->setActive(true) != kResultOk) return 3;
  processor->setProcessing(true);

  sv::HostProcessData data;
  if (!data.prepare(*component, block, sv::kSample32)) return 3;
  data.processMode = sv::kRealtime;
  sv::ProcessContext ctx{};
  ctx.sampleRate = rate;
  data.processContext = &ctx;
  sv::ParameterChanges in_changes, out_changes;
  data.inputParameterChanges = &in_changes;
  data.outputParameterChanges = &out_changes;

  // parameters set before the model is loaded: a zero-sample process call carries them
  bool any = false;
  for (const Event& e : events)
    if (e.block < 0 && e.name != "reset") any |= Queue(&in_changes, e);
  if (any) {
    data.numSamples = 0;
    processor->process(data);
    in_changes.clearQueue();
  }

  // ---- model load: the controller's "param_change" message (controller.cc -> processor.cc:274-300) ----
  tresult load = -1;
  if (std::strcmp(toml, "-") != 0) {
    IPtr<sv::IMessage> msg = Steinberg::owned(new sv::HostMessage());
    msg->setMessageID("param_change");
    const sv::ParamID pid = static_cast<sv::ParamID>(beatrice::common::ParameterID::kModel);
    msg->getAttributes()->setBinary("param_id", &pid, sizeof(pid));
    msg->getAttributes()->setBinary("data", toml, static_cast<Steinberg::uint32>(std::strlen(toml)));
    load = connection->notify(msg);
  }

  // ---- audio ----
  tresult last = kResultOk;
  long bi = 0;
  for (size_t i = 0; i < x.size(); i += block, ++bi) {
    for (const Event& e : events) {
      if (e.block != bi) continue;
      if (e.name == "reset") {
        processor->setProcessing(false);
        component->setActive(false);
        component->setActive(true);
        processor->setProcessing(true);
      } else {
        Queue(&in_changes, e);
      }
    }
    const int n = static_cast<int>(std::min<size_t>(block, x.size() - i));
    data.numSamples = n;
    std::memcpy(data.inputs[0].channelBuffers32[0], x.data() + i, sizeof(float) * n);
    data.inputs[0].silenceFlags = 0;
    data.outputs[0].silenceFlags = 0;
    last = processor->process(data);
    std::memcpy(x.data() + i, data.outputs[0].channelBuffers32[0], sizeof(float) * n);
    in_changes.clearQueue();
  }

  FILE* f = std::fopen(argv[4], "wb");
  if (!f) return 4;
  std::fwrite(x.data(), 4, x.size(), f);
  std::fclose(f);
  std::printf("status: load=%d process=%d\n", static_cast<int>(load), static_cast<int>(last));

  processor->setProcessing(false);
  component->setActive(false);
  data.unprepare();
  component->terminate();
  return 0;
}
