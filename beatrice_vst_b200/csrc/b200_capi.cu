// The drop-in boundary: every extern "C" symbol of the reference header
// lib/beatricelib/beatrice.h (20a2 :39-120, 20b1 :122-203, 20rc0 :205-343), backed by the
// CUDA engine.  One context = one voice stream = a batch of 1 with its own CUDA stream,
// pinned staging and (lazily captured) CUDA graph, so that independent plug-in instances on
// different host threads do not serialise against each other.
//
// Call-site contract honoured here (reference src/common/processor_core_2.cc):
//  * Read*Parameters return Beatrice_ErrorCode values 0..4 (:302-351, error.h:13-16);
//  * Process functions return void and always write the whole output (:183-255);
//  * embedding pointers are caller-owned; SetCodebook receives ONE speaker's 512x128 slice
//    (:118-121, :447-450);
//  * contexts are destroyed / re-created by ResetContext (:258-266) -> fresh zero state.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "b200_common.h"
#include "b200_engine.h"
#include "b200_kernels.h"

using namespace b200;

namespace {

struct Pinned {
  void* p = nullptr;
  size_t bytes = 0;
  void Alloc(size_t n) {
    Free();
    B200_CHECK(cudaMallocHost(&p, n));
    bytes = n;
    std::memset(p, 0, n);
  }
  void Free() {
    if (p) cudaFreeHost(p);
    p = nullptr;
  }
  ~Pinned() { Free(); }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

struct StreamOwner {
  cudaStream_t s = nullptr;
  int device = -1;
  void Ensure(int dev) {
    if (s) return;
    device = dev;
    B200_CHECK(cudaSetDevice(dev));
    InstallSpinDebug(dev);
    B200_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  }
  ~StreamOwner() {
    if (s) {
      cudaStreamSynchronize(s);
      cudaStreamDestroy(s);
    }
  }
};

struct PhoneExtractorObj {
  EncoderModel m;
};
struct PitchEstimatorObj {
  EncoderModel m;
};
struct WaveformGeneratorObj {
  WaveModel m;
};
struct EmbeddingSetterObj {
  SetterModel m;
};

struct PhoneContextObj {
  FamilyDims dims;
  StreamOwner stream;
  EncoderState st;
  GraphRunner graph;
  Pinned in, out;
  DeviceBuffer phone_out;    // after VQ
  DeviceBuffer codebook;     // device copy of the current speaker's codebook
  DeviceBuffer vq_args;      // { const float* codebook; int n; }
  int vq_n = 0;
  const float* host_codebook = nullptr;
  bool codebook_dirty = false;
  uint64_t codebook_sum = 0;
  bool args_dirty = true;
};
struct VqArgs {
  const float* codebook;
  int n;
  int pad;
};

struct PitchContextObj {
  FamilyDims dims;
  StreamOwner stream;
  EncoderState st;
  GraphRunner graph;
  Pinned in, out;            // out: int q + 4 floats
  DeviceBuffer range;        // int[2] = {min, max}
  DeviceBuffer q, feat;
  int min_q = 1, max_q = 1;
  bool range_dirty = true;
};

struct WaveformContextObj {
  FamilyDims dims;
  StreamOwner stream;
  WaveState st;
  GraphRunner graph;
  Pinned in, out;            // in: phone | q | feat | speaker(legacy)
  DeviceBuffer emb_tmp;      // [256] staging for the rc0 setters
  uint64_t hops = 0;
  void EnsureCond() {
    stream.Ensure(DefaultDevice());
    st.AllocCond(dims, 1, stream.device);
    if (!emb_tmp.p) emb_tmp.Alloc(stream.device, sizeof(float) * kHidden, true);
  }
};

struct EmbeddingContextObj {
  DeviceBuffer kv;  // registered 384 x 128 embedding
  bool registered = false;
};

uint64_t Checksum(const float* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  const uint32_t* u = reinterpret_cast<const uint32_t*>(p);
  for (size_t i = 0; i < n; ++i) h = (h ^ u[i]) * 1099511628211ull;
  return h;
}

// ---------------------------------------------------------------------------------------
// per-frame entry points
// ---------------------------------------------------------------------------------------
void ExtractPhoneImpl(const PhoneExtractorObj* pe, const float* in, float* out, PhoneContextObj* c) {
  const int P = c->dims.phone_channels;
  if (!pe->m.loaded) {
    std::memset(out, 0, sizeof(float) * P);
    return;
  }
  const int dev = pe->m.device;
  c->stream.Ensure(dev);
  B200_CHECK(cudaSetDevice(dev));
  cudaStream_t s = c->stream.s;
  if (!c->st.Matches(&pe->m)) {
    B200_CHECK(cudaStreamSynchronize(s));
    c->graph.Reset();
    c->st.Build(&pe->m, 1, dev, nullptr, DefaultTcMode());
    c->in.Alloc(sizeof(float) * kInHop);
    c->out.Alloc(sizeof(float) * P);
    c->phone_out.Alloc(dev, sizeof(float) * P, true);
    c->vq_args.Alloc(dev, sizeof(VqArgs), true);
    c->args_dirty = true;
  }
  const bool vq = c->dims.has_setter;
  if (vq) {
    // upload the codebook only when VQ is on and the caller's current slice changed
    if (c->vq_n > 0 && c->host_codebook) {
      const size_t n = static_cast<size_t>(kCodebookSize) * P;
      const uint64_t sum = Checksum(c->host_codebook, n);
      if (c->codebook_dirty || sum != c->codebook_sum || !c->codebook.p) {
        if (!c->codebook.p) c->codebook.Alloc(dev, n * sizeof(float), false);
        B200_CHECK(cudaMemcpyAsync(c->codebook.p, c->host_codebook, n * sizeof(float), cudaMemcpyHostToDevice, s));
        c->codebook_sum = sum;
        c->codebook_dirty = false;
        c->args_dirty = true;
      }
    }
    if (c->args_dirty) {
      VqArgs a;
      a.codebook = (c->vq_n > 0 && c->codebook.p) ? c->codebook.as<float>() : nullptr;
      a.n = a.codebook ? c->vq_n : 0;
      a.pad = 0;
      B200_CHECK(cudaMemcpyAsync(c->vq_args.p, &a, sizeof(a), cudaMemcpyHostToDevice, s));
      c->args_dirty = false;
    }
  }
  std::memcpy(c->in.p, in, sizeof(float) * kInHop);
  const float* result = vq ? c->phone_out.as<float>() : c->st.head_out.as<float>();
  c->graph.Run(
      s,
      [&](cudaStream_t st) {
        B200_CHECK(cudaMemcpyAsync(c->st.in_stage.p, c->in.p, sizeof(float) * kInHop, cudaMemcpyHostToDevice, st));
        RunProgram(c->st.program, st);
        if (vq) {
          const VqArgs* a = c->vq_args.as<VqArgs>();
          LaunchVq(c->st.head_out.as<float>(), c->phone_out.as<float>(), &a->codebook, &a->n, P, 1, st);
        }
        B200_CHECK(cudaMemcpyAsync(c->out.p, result, sizeof(float) * P, cudaMemcpyDeviceToHost, st));
      },
      GraphsEnabled());
  g_kernel_launches.fetch_add(c->st.program.size() + (vq ? 1 : 0), std::memory_order_relaxed);
  B200_CHECK(cudaStreamSynchronize(s));
  std::memcpy(out, c->out.p, sizeof(float) * P);
}

void EstimatePitchImpl(const PitchEstimatorObj* pi, const float* in, int* q, float* feat, PitchContextObj* c) {
  if (!pi->m.loaded) {
    *q = 1;
    std::memset(feat, 0, sizeof(float) * kPitchFeatures);
    return;
  }
  const int dev = pi->m.device;
  c->stream.Ensure(dev);
  B200_CHECK(cudaSetDevice(dev));
  cudaStream_t s = c->stream.s;
  if (!c->st.Matches(&pi->m)) {
    B200_CHECK(cudaStreamSynchronize(s));
    c->graph.Reset();
    c->st.Build(&pi->m, 1, dev, nullptr, DefaultTcMode());
    c->in.Alloc(sizeof(float) * kInHop);
    c->out.Alloc(sizeof(float) * 8);
    c->range.Alloc(dev, sizeof(int) * 2, true);
    c->q.Alloc(dev, sizeof(int) * 2, true);
    c->feat.Alloc(dev, sizeof(float) * kPitchFeatures, true);
    c->range_dirty = true;
  }
  if (c->range_dirty) {
    const int r[2] = {c->min_q, c->max_q};
    B200_CHECK(cudaMemcpyAsync(c->range.p, r, sizeof(r), cudaMemcpyHostToDevice, s));
    c->range_dirty = false;
  }
  std::memcpy(c->in.p, in, sizeof(float) * kInHop);
  c->graph.Run(
      s,
      [&](cudaStream_t st) {
        B200_CHECK(cudaMemcpyAsync(c->st.in_stage.p, c->in.p, sizeof(float) * kInHop, cudaMemcpyHostToDevice, st));
        RunProgram(c->st.program, st);
        LaunchPitchArgmax(c->st.head_out.as<float>(), c->dims.pitch_bins, c->range.as<int>(), c->range.as<int>() + 1,
                          c->q.as<int>(), c->feat.as<float>(), 1, st);
        B200_CHECK(cudaMemcpyAsync(c->out.p, c->q.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        B200_CHECK(cudaMemcpyAsync(c->out.as<float>() + 1, c->feat.p, sizeof(float) * kPitchFeatures,
                                   cudaMemcpyDeviceToHost, st));
      },
      GraphsEnabled());
  g_kernel_launches.fetch_add(c->st.program.size() + 1, std::memory_order_relaxed);
  B200_CHECK(cudaStreamSynchronize(s));
  std::memcpy(q, c->out.p, sizeof(int));
  std::memcpy(feat, c->out.as<float>() + 1, sizeof(float) * kPitchFeatures);
}

void GenerateWaveformImpl(const WaveformGeneratorObj* wg, const float* phone, const int* q, const float* feat,
                          const float* speaker_or_null, float* out, WaveformContextObj* c) {
  if (!wg->m.loaded) {
    std::memset(out, 0, sizeof(float) * kOutHop);
    return;
  }
  const int P = c->dims.phone_channels;
  const int dev = wg->m.device;
  c->stream.Ensure(dev);
  B200_CHECK(cudaSetDevice(dev));
  cudaStream_t s = c->stream.s;
  if (!c->st.Matches(&wg->m)) {
    B200_CHECK(cudaStreamSynchronize(s));
    c->graph.Reset();
    c->st.Build(&wg->m, 1, dev, DefaultTcMode());
    c->in.Alloc(sizeof(float) * (P + 8 + kHidden));
    c->out.Alloc(sizeof(float) * kOutHop);
  }
  float* hin = c->in.as<float>();
  std::memcpy(hin, phone, sizeof(float) * P);
  std::memcpy(hin + P, q, sizeof(int));
  std::memcpy(hin + P + 4, feat, sizeof(float) * kPitchFeatures);
  const bool legacy = !c->dims.has_setter;
  if (legacy) {
    if (speaker_or_null) std::memcpy(hin + P + 8, speaker_or_null, sizeof(float) * kHidden);
    else std::memset(hin + P + 8, 0, sizeof(float) * kHidden);
  }
  c->graph.Run(
      s,
      [&](cudaStream_t st) {
        B200_CHECK(cudaMemcpyAsync(c->st.phone_in.p, hin, sizeof(float) * P, cudaMemcpyHostToDevice, st));
        B200_CHECK(cudaMemcpyAsync(c->st.q_in.p, hin + P, sizeof(int), cudaMemcpyHostToDevice, st));
        B200_CHECK(cudaMemcpyAsync(c->st.feat_in.p, hin + P + 4, sizeof(float) * kPitchFeatures,
                                   cudaMemcpyHostToDevice, st));
        if (legacy)
          B200_CHECK(cudaMemcpyAsync(c->st.spk.p, hin + P + 8, sizeof(float) * kHidden, cudaMemcpyHostToDevice, st));
        RunProgram(c->st.program, st);
        B200_CHECK(cudaMemcpyAsync(c->out.p, c->st.out.p, sizeof(float) * kOutHop, cudaMemcpyDeviceToHost, st));
      },
      GraphsEnabled());
  g_kernel_launches.fetch_add(static_cast<size_t>(c->st.LaunchesPerHop()), std::memory_order_relaxed);
  B200_CHECK(cudaStreamSynchronize(s));
  ++c->hops;
  std::memcpy(out, c->out.p, sizeof(float) * kOutHop);
}

// The guarded forms the ABI calls: on a (latched) failure the whole output is written as silence, like the call
// site's own guards (processor_core_2.cc:26-43) -- never an abort, never a CPU fallback.
void ExtractPhone(const PhoneExtractorObj* pe, const float* in, float* out, PhoneContextObj* c) {
  B200_GUARDED(c->st.model = nullptr; std::memset(out, 0, sizeof(float) * c->dims.phone_channels), ExtractPhoneImpl(pe, in, out, c););
}
void EstimatePitch(const PitchEstimatorObj* pi, const float* in, int* q, float* feat, PitchContextObj* c) {
  B200_GUARDED(c->st.model = nullptr; *q = 1; std::memset(feat, 0, sizeof(float) * kPitchFeatures), EstimatePitchImpl(pi, in, q, feat, c););
}
void GenerateWaveform(const WaveformGeneratorObj* wg, const float* phone, const int* q, const float* feat,
                      const float* speaker_or_null, float* out, WaveformContextObj* c) {
  B200_GUARDED(c->st.model = nullptr; std::memset(out, 0, sizeof(float) * kOutHop), GenerateWaveformImpl(wg, phone, q, feat, speaker_or_null, out, c););
}
// Read*Parameters: host-side validation errors keep their Beatrice_ErrorCode (beatrice.h:30-37); a device failure
// while uploading reads as kFileOpenError (1) -- the model is not usable, and the call site surfaces the code
// (processor_core_2.cc:302-351)
template <class M>
int GuardedLoad(M* m, const char* path) {
  int rc = 1;
  B200_GUARDED(m->loaded = false; rc = 1, rc = m->LoadFromFile(path););
  return rc;
}

int ReadSpeakerTable(int family, const char* path, std::vector<uint8_t>* bytes, FileImage* img) {
  if (const int e = LoadFileBytes(path, bytes)) return e;
  const FamilyDims d = kFamilies[family];
  return ParseFileImage(bytes->data(), bytes->size(), family, kKindSpeakers, kKindFormant,
                        [&](uint32_t n) { return static_cast<long long>(SpeakerPayloadFloats(d, n)); }, img);
}

}  // namespace

// =========================================================================================
// exported symbols
// =========================================================================================
#define B200_COMMON_API(PFX, FAM)                                                                   \
  extern "C" {                                                                                      \
  void* PFX##_CreatePhoneExtractor(void) {                                                          \
    auto* o = new PhoneExtractorObj();                                                              \
    o->m.dims = kFamilies[FAM];                                                                     \
    o->m.is_pitch = false;                                                                          \
    return o;                                                                                       \
  }                                                                                                 \
  void PFX##_DestroyPhoneExtractor(void* p) { delete static_cast<PhoneExtractorObj*>(p); }          \
  void* PFX##_CreatePhoneContext1(void) {                                                           \
    auto* o = new PhoneContextObj();                                                                \
    o->dims = kFamilies[FAM];                                                                       \
    return o;                                                                                       \
  }                                                                                                 \
  void PFX##_DestroyPhoneContext1(void* p) { delete static_cast<PhoneContextObj*>(p); }             \
  int PFX##_ReadPhoneExtractorParameters(void* m, const char* path) {                               \
    return GuardedLoad(&static_cast<PhoneExtractorObj*>(m)->m, path);                                \
  }                                                                                                 \
  void PFX##_ExtractPhone1(const void* m, const float* in, float* out, void* c) {                   \
    ExtractPhone(static_cast<const PhoneExtractorObj*>(m), in, out, static_cast<PhoneContextObj*>(c)); \
  }                                                                                                 \
  void* PFX##_CreatePitchEstimator(void) {                                                          \
    auto* o = new PitchEstimatorObj();                                                              \
    o->m.dims = kFamilies[FAM];                                                                     \
    o->m.is_pitch = true;                                                                           \
    return o;                                                                                       \
  }                                                                                                 \
  void PFX##_DestroyPitchEstimator(void* p) { delete static_cast<PitchEstimatorObj*>(p); }          \
  void* PFX##_CreatePitchContext1(void) {                                                           \
    auto* o = new PitchContextObj();                                                                \
    o->dims = kFamilies[FAM];                                                                       \
    o->max_q = kFamilies[FAM].pitch_bins - 1;                                                       \
    return o;                                                                                       \
  }                                                                                                 \
  void PFX##_DestroyPitchContext1(void* p) { delete static_cast<PitchContextObj*>(p); }             \
  int PFX##_ReadPitchEstimatorParameters(void* m, const char* path) {                               \
    return GuardedLoad(&static_cast<PitchEstimatorObj*>(m)->m, path);                                \
  }                                                                                                 \
  void PFX##_SetMinQuantizedPitch(void* c, int v) {                                                 \
    auto* o = static_cast<PitchContextObj*>(c);                                                     \
    o->min_q = v;                                                                                   \
    o->range_dirty = true;                                                                          \
  }                                                                                                 \
  void PFX##_SetMaxQuantizedPitch(void* c, int v) {                                                 \
    auto* o = static_cast<PitchContextObj*>(c);                                                     \
    o->max_q = v;                                                                                   \
    o->range_dirty = true;                                                                          \
  }                                                                                                 \
  void PFX##_EstimatePitch1(const void* m, const float* in, int* q, float* feat, void* c) {         \
    EstimatePitch(static_cast<const PitchEstimatorObj*>(m), in, q, feat, static_cast<PitchContextObj*>(c)); \
  }                                                                                                 \
  int PFX##_ReadNSpeakers(const char* path, int* n) {                                               \
    std::vector<uint8_t> bytes;                                                                     \
    FileImage img;                                                                                  \
    if (const int e = ReadSpeakerTable(FAM, path, &bytes, &img)) return e;                          \
    *n = static_cast<int>(img.count);                                                               \
    return 0;                                                                                       \
  }                                                                                                 \
  void* PFX##_CreateWaveformGenerator(void) {                                                       \
    auto* o = new WaveformGeneratorObj();                                                           \
    o->m.dims = kFamilies[FAM];                                                                     \
    return o;                                                                                       \
  }                                                                                                 \
  void PFX##_DestroyWaveformGenerator(void* p) { delete static_cast<WaveformGeneratorObj*>(p); }    \
  void* PFX##_CreateWaveformContext1(void) {                                                        \
    auto* o = new WaveformContextObj();                                                             \
    o->dims = kFamilies[FAM];                                                                       \
    return o;                                                                                       \
  }                                                                                                 \
  void PFX##_DestroyWaveformContext1(void* p) { delete static_cast<WaveformContextObj*>(p); }       \
  int PFX##_ReadWaveformGeneratorParameters(void* m, const char* path) {                            \
    return GuardedLoad(&static_cast<WaveformGeneratorObj*>(m)->m, path);                             \
  }                                                                                                 \
  }

B200_COMMON_API(Beatrice20a2, 0)
B200_COMMON_API(Beatrice20b1, 1)
B200_COMMON_API(Beatrice20rc0, 2)

#define B200_LEGACY_API(PFX, FAM)                                                                   \
  extern "C" {                                                                                      \
  int PFX##_ReadSpeakerEmbeddings(const char* path, float* table) {                                 \
    std::vector<uint8_t> bytes;                                                                     \
    FileImage img;                                                                                  \
    if (const int e = ReadSpeakerTable(FAM, path, &bytes, &img)) return e;                          \
    std::memcpy(table, img.payload, img.n_floats * sizeof(float));                                  \
    return 0;                                                                                       \
  }                                                                                                 \
  void PFX##_GenerateWaveform1(const void* m, const float* phone, const int* q, const float* feat,  \
                               const float* speaker, float* out, void* c) {                         \
    GenerateWaveform(static_cast<const WaveformGeneratorObj*>(m), phone, q, feat, speaker, out,     \
                     static_cast<WaveformContextObj*>(c));                                          \
  }                                                                                                 \
  }

B200_LEGACY_API(Beatrice20a2, 0)
B200_LEGACY_API(Beatrice20b1, 1)

extern "C" {

// beatrice.h:239-242
void Beatrice20rc0_SetVQNumNeighbors(void* c, int n) {
  auto* o = static_cast<PhoneContextObj*>(c);
  o->vq_n = std::min(std::max(n, 0), kCodebookSize);
  o->args_dirty = true;
}

// beatrice.h:276-290
int Beatrice20rc0_ReadSpeakerEmbeddings(const char* path, float* codebooks, float* additive, float* formant,
                                        float* kv) {
  std::vector<uint8_t> bytes;
  FileImage img;
  if (const int e = ReadSpeakerTable(2, path, &bytes, &img)) return e;
  if (img.kind != kKindSpeakers) return 4;
  const float* p = img.payload;
  std::memcpy(formant, p, sizeof(float) * kNFormant * kHidden);
  p += kNFormant * kHidden;
  const size_t cb = static_cast<size_t>(kCodebookSize) * kFamilies[2].phone_channels;
  const size_t kvn = static_cast<size_t>(kKvLength) * kKvChannels;
  for (uint32_t i = 0; i < img.count; ++i) {
    std::memcpy(codebooks + i * cb, p, sizeof(float) * cb);
    p += cb;
    std::memcpy(additive + static_cast<size_t>(i) * kHidden, p, sizeof(float) * kHidden);
    p += kHidden;
    std::memcpy(kv + i * kvn, p, sizeof(float) * kvn);
    p += kvn;
  }
  return 0;
}

// beatrice.h:301-307
void Beatrice20rc0_GenerateWaveform1(const void* m, const float* phone, const int* q, const float* feat, float* out,
                                     void* c) {
  GenerateWaveform(static_cast<const WaveformGeneratorObj*>(m), phone, q, feat, nullptr, out,
                   static_cast<WaveformContextObj*>(c));
}

// beatrice.h:309-317
void* Beatrice20rc0_CreateEmbeddingSetter(void) {
  auto* o = new EmbeddingSetterObj();
  o->m.dims = kFamilies[2];
  return o;
}
void Beatrice20rc0_DestroyEmbeddingSetter(void* p) { delete static_cast<EmbeddingSetterObj*>(p); }
void* Beatrice20rc0_CreateEmbeddingContext(void) { return new EmbeddingContextObj(); }
void Beatrice20rc0_DestroyEmbeddingContext(void* p) { delete static_cast<EmbeddingContextObj*>(p); }
int Beatrice20rc0_ReadEmbeddingSetterParameters(void* m, const char* path) {
  return GuardedLoad(&static_cast<EmbeddingSetterObj*>(m)->m, path);
}

// beatrice.h:318-322.  Only the pointer is recorded; the 256 KiB slice is uploaded lazily and
// only while VQ is enabled (in morph mode the call site re-points it every frame,
// processor_core_2.cc:118-121).
void Beatrice20rc0_SetCodebook(void* c, const float* codebook) {
  auto* o = static_cast<PhoneContextObj*>(c);
  if (o->host_codebook != codebook) o->codebook_dirty = true;
  o->host_codebook = codebook;
}

static void ProjectInto(const float* W, const float* b, const float* emb, WaveformContextObj* wc, bool formant) {
  wc->EnsureCond();
  float* dst = formant ? wc->st.formant.as<float>() : wc->st.spk.as<float>();
  B200_CHECK(cudaSetDevice(wc->stream.device));
  cudaStream_t s = wc->stream.s;
  B200_CHECK(cudaMemcpyAsync(wc->emb_tmp.p, emb, sizeof(float) * kHidden, cudaMemcpyHostToDevice, s));
  LaunchProject256(W, b, wc->emb_tmp.as<float>(), kHidden, nullptr, dst, nullptr, 1, s);
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  B200_CHECK(cudaStreamSynchronize(s));
}

// beatrice.h:323-327
void Beatrice20rc0_SetAdditiveSpeakerEmbedding(const void* m, const float* emb, void* /*ec*/, void* wc_) {
  const auto* es = static_cast<const EmbeddingSetterObj*>(m);
  if (!es->m.loaded) return;
  auto* wc = static_cast<WaveformContextObj*>(wc_);
  B200_GUARDED((void)0, ProjectInto(es->m.add_w, es->m.add_b, emb, wc, false););
}
// beatrice.h:328-332
void Beatrice20rc0_SetFormantShiftEmbedding(const void* m, const float* emb, void* /*ec*/, void* wc_) {
  const auto* es = static_cast<const EmbeddingSetterObj*>(m);
  if (!es->m.loaded) return;
  auto* wc = static_cast<WaveformContextObj*>(wc_);
  B200_GUARDED((void)0, ProjectInto(es->m.for_w, es->m.for_b, emb, wc, true););
}
// beatrice.h:333-338
void Beatrice20rc0_RegisterKeyValueSpeakerEmbedding(const void* /*m*/, const float* kv, void* ec_) {
  auto* ec = static_cast<EmbeddingContextObj*>(ec_);
  B200_GUARDED(ec->registered = false, {
    const int dev = DefaultDevice();
    const size_t bytes = sizeof(float) * kKvLength * kKvChannels;
    if (!ec->kv.p) ec->kv.Alloc(dev, bytes, false);
    B200_CHECK(cudaSetDevice(ec->kv.device));
    // the consumer (kv_film_kernel) runs on the waveform context's non-blocking stream: complete the copy first
    UploadSync(ec->kv.p, kv, bytes);
    ec->registered = true;
  });
}
// beatrice.h:339-343
void Beatrice20rc0_SetKeyValueSpeakerEmbedding(const void* m, int block, void* ec_, void* wc_) {
  const auto* es = static_cast<const EmbeddingSetterObj*>(m);
  auto* ec = static_cast<EmbeddingContextObj*>(ec_);
  auto* wc = static_cast<WaveformContextObj*>(wc_);
  if (!es->m.loaded || !ec->registered || block < 0 || block >= kNBlocks) return;
  B200_GUARDED((void)0, {
    wc->EnsureCond();
    B200_CHECK(cudaSetDevice(wc->stream.device));
    static const int kC[4] = {128, 64, 32, 16};
    LaunchKvFilm(ec->kv.as<float>(), nullptr, 0, es->m.query[block], es->m.film_w[block], es->m.film_b[block],
                 kC[block], wc->st.film[block].as<float>(), nullptr, 1, wc->stream.s);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    B200_CHECK(cudaStreamSynchronize(wc->stream.s));
  });
}

// test-only tap (include/beatrice_b200.h)
int BeatriceB200_WaveformTap(const void* wc_, int which, float* out, int capacity) {
  const auto* wc = static_cast<const WaveformContextObj*>(wc_);
  if (!wc->st.model || wc->hops == 0 || Failed()) return -1;
  try {
  B200_CHECK(cudaSetDevice(wc->stream.device));
  B200_CHECK(cudaStreamSynchronize(wc->stream.s));
  const uint64_t last = wc->hops - 1;
  auto fetch = [&](const Ring& r, std::vector<float>* v) {
    v->resize(static_cast<size_t>(r.T) * r.C);
    const size_t off = static_cast<size_t>(last % r.slots) * r.T * r.C;
    B200_CHECK(cudaMemcpy(v->data(), r.base + off, v->size() * sizeof(float), cudaMemcpyDeviceToHost));
  };
  std::vector<float> v;
  if (which == 0) {
    fetch(wc->st.arena.ring(wc->st.ring_hidden), &v);
  } else if (which == 1) {
    fetch(wc->st.arena.ring(wc->st.ring_pre), &v);
  } else if (which >= 2 && which < 6) {
    std::vector<float> a, b, c;
    fetch(wc->st.arena.ring(wc->st.ring_stage_out[which - 2][0]), &a);
    fetch(wc->st.arena.ring(wc->st.ring_stage_out[which - 2][1]), &b);
    fetch(wc->st.arena.ring(wc->st.ring_stage_out[which - 2][2]), &c);
    v.resize(a.size());
    for (size_t i = 0; i < a.size(); ++i) v[i] = ((a[i] + b[i]) + c[i]) * (1.0f / 3.0f);
  } else {
    return -1;
  }
  const int n = static_cast<int>(v.size());
  if (out && capacity >= n) std::memcpy(out, v.data(), sizeof(float) * n);
  return n;
  } catch (const Failure&) {
    return -1;
  }
}

// Arithmetic of the contexts behind beatrice.h that are BUILT after this call (a context builds on its first
// per-frame call, or when it meets a new model): BEATRICE_B200_PRECISION_* or -1 for the default.
void BeatriceB200_SetDefaultPrecision(int precision) { SetDefaultTcMode(precision); }

// sticky failure state (b200_common.h)
int BeatriceB200_LastError(void) { return LastErrorCode(); }
const char* BeatriceB200_LastErrorString(void) { return LastErrorText(); }
void BeatriceB200_ClearError(void) { ClearError(); }

}  // extern "C"
