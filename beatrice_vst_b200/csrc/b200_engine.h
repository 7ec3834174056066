// Host-side engine of libbeatrice_b200: parameter files -> device weights, per-stream ring
// state, and the per-hop launch programs of the three model parts.
#ifndef BEATRICE_B200_ENGINE_H_
#define BEATRICE_B200_ENGINE_H_

#include <atomic>
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "b200_common.h"
#include "b200_kernels.h"
#include "b200_enc.h"
#include "b200_mrf.h"

namespace b200 {

// ---------------------------------------------------------------------------------------
// parameter files (format: beatrice_vst_b200/model_spec.py docstring)
// ---------------------------------------------------------------------------------------
enum FileKind : uint32_t { kKindPhone = 1, kKindPitch = 2, kKindWavegen = 3, kKindSetter = 4, kKindSpeakers = 5, kKindFormant = 6 };
constexpr uint32_t kFileMagic = 0x42323042u;

size_t PhoneParamCount(const FamilyDims& d);
size_t PitchParamCount(const FamilyDims& d);
size_t WaveParamCount(const FamilyDims& d);
size_t SetterParamCount();
size_t SpeakerPayloadFloats(const FamilyDims& d, uint32_t n_speakers);

struct FileImage {
  uint32_t family = 0, kind = 0, count = 0;
  const float* payload = nullptr;  // points into the caller's bytes
  size_t n_floats = 0;
};
// Returns a Beatrice_ErrorCode value (0 ok, 1 open, 2 too small, 3 too large, 4 invalid).
int LoadFileBytes(const char* utf8_path, std::vector<uint8_t>* bytes);
// expected_floats(count) < 0 -> invalid
int ParseFileImage(const void* data, size_t size, int family, uint32_t kind_a, uint32_t kind_b,
                   const std::function<long long(uint32_t)>& expected_floats, FileImage* out);

// ---------------------------------------------------------------------------------------
// device plumbing
// ---------------------------------------------------------------------------------------
void InstallSpinDebug(int device);  // dead-lock records of the polling kernels become visible to Fail()
int UsableDeviceCount();           // never aborts
int DefaultDevice();               // env BEATRICE_B200_DEVICE or 0; aborts when no GPU is usable
bool GraphsEnabled();              // env BEATRICE_B200_NO_GRAPH=1 disables CUDA graphs
bool FusedMrfEnabled();            // env BEATRICE_B200_NO_FUSED_MRF=1 keeps the per-conv kernels

struct DeviceBuffer {
  void* p = nullptr;
  size_t bytes = 0;
  int device = -1;
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  ~DeviceBuffer();
  void Alloc(int dev, size_t n_bytes, bool zero);
  void Free();
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

struct ConvW {
  const float* w = nullptr;
  const float* b = nullptr;
  int k = 0, cin = 0, cout = 0;
  // tensor-core operand images (b200_tc.cu), filled by TcWeights::Pack
  const void* tc_hi = nullptr;
  const void* tc_lo = nullptr;
  int tc_bn = 0, tc_kc = 0;
  int tc_bn_cap = 128;  // requested N-tile cap (smaller = more CTAs for layers with few rows)
};

// Precision of the conv GEMMs: fp32 CUDA cores (parity path), bf16 tcgen05, or split-bf16
// tcgen05 (hi/lo decomposition, three MMAs, near-fp32 accuracy).
enum TcMode : int { kTcOff = 0, kTcBf16 = 1, kTcSplit = 2 };
TcMode DefaultTcMode();  // BeatriceB200_SetDefaultPrecision, else env BEATRICE_B200_PRECISION = f32 | bf16 | bf16x3 (default bf16x3)
void SetDefaultTcMode(int mode);  // 0..2, anything else: back to the environment / built-in default

// bf16 re-packing of a set of conv weights, resident in HBM next to the fp32 blob.
struct TcWeights {
  DeviceBuffer buf;
  // packs every conv in `convs` whose shape the tensor-core kernel supports; host_blob/dev_blob
  // give the fp32 payload the ConvW pointers refer to
  void Pack(int device, const float* host_blob, const float* dev_blob, const std::vector<ConvW*>& convs);
};

// ---------------------------------------------------------------------------------------
// immutable model objects (weights in HBM)
// ---------------------------------------------------------------------------------------
struct EncoderModel {  // PhoneExtractor / PitchEstimator
  FamilyDims dims;
  bool is_pitch = false;
  bool loaded = false;
  uint64_t generation = 0;
  int device = -1;
  DeviceBuffer blob;
  ConvW front[6];
  int stride[6];
  int n_res = 0, width = 0, head_out = 0;
  const float* gamma[6];
  const float* beta[6];
  ConvW res[6];
  int dil[6];
  ConvW head;
  TcWeights tc;
  // fused residual-stack kernel (b200_enc.cu): weight image + the blocks' bias / gamma / beta made contiguous
  DeviceBuffer rs_w, rs_par;
  bool rs_ok = false;
  int rs_n_blk = 0, rs_kind[8] = {}, rs_dil[8] = {};   // the chain: [front layer 5] + residual blocks + [head]
  bool rs_head = false;                                 // head (k = 1) is the chain's last block
  const float* rs_bias = nullptr;
  const float* rs_gamma = nullptr;
  const float* rs_beta = nullptr;
  // returns Beatrice_ErrorCode; host validation happens before any CUDA call
  int LoadFromImage(const void* data, size_t size, int on_device = -1);
  int LoadFromFile(const char* utf8_path, int on_device = -1);
};

struct WaveModel {
  FamilyDims dims;
  bool loaded = false;
  uint64_t generation = 0;
  int device = -1;
  DeviceBuffer blob, ups_bias;
  ConvW embed;
  const float* pitch_emb = nullptr;
  const float* feat_proj = nullptr;
  ConvW pre;
  ConvW ups[4];  // as 2-tap conv, cout = r*C_out, bias replicated per phase
  ConvW c1[4][3][3], c2[4][3][3];
  ConvW post;
  TcWeights tc;
  // fused MRF kernel (b200_mrf.cu): weight images per precision [0] bf16, [1] split bf16, and
  // the six biases of a branch made contiguous
  DeviceBuffer mrf_w[2], mrf_bias;
  // pre conv (k = 7) as a one-block chain of the cluster kernel in b200_enc.cu
  DeviceBuffer pre_w, pre_par;
  bool pre_chain_ok = false;
  // the same with the conditioning (1x1 phone embedding + pitch / feature / speaker terms) as the chain's first block
  DeviceBuffer cond_pre_w, cond_pre_par;
  bool cond_chain_ok = false;
  const uint16_t* mrf_w_ptr[2][4][3] = {};
  const float* mrf_bias_ptr[4][3] = {};
  // upsampler images for the fused kernel's prologue (MrfUpsDesc, split precision) and the un-replicated biases
  DeviceBuffer ups_img;
  const uint16_t* ups_img_ptr[4] = {};
  const float* ups_bias_raw[4] = {};
  int LoadFromImage(const void* data, size_t size, int on_device = -1);
  int LoadFromFile(const char* utf8_path, int on_device = -1);
};

struct SetterModel {
  FamilyDims dims;
  bool loaded = false;
  int device = -1;
  DeviceBuffer blob;
  const float *add_w = nullptr, *add_b = nullptr, *for_w = nullptr, *for_b = nullptr;
  const float* query[4];
  const float* film_w[4];
  const float* film_b[4];
  int LoadFromImage(const void* data, size_t size, int on_device = -1);
  int LoadFromFile(const char* utf8_path, int on_device = -1);
};

// ---------------------------------------------------------------------------------------
// per-hop launch program
// ---------------------------------------------------------------------------------------
struct Op {
  std::string name;
  std::function<void(cudaStream_t)> launch;
  double flops = 0.0;  // algorithmic, whole batch
  double bytes = 0.0;  // algorithmic: weights + activations read + written
  bool is_mrf = false; // belongs to the vocoder MRF Conv1d stage (the dominant roofline)
};

struct Ring {
  float* base = nullptr;    // fp32 ring, or nullptr for a bf16-only ring
  uint16_t* hi = nullptr;   // bf16 ring: rounded value
  uint16_t* lo = nullptr;   // bf16 of the rounding residual (split-bf16 mode only)
  int slots = 1, T = 1, C = 1;
  bool is_bf16 = false, has_lo = false;
  size_t StreamStride() const { return static_cast<size_t>(slots) * T * C; }
};

// Arena of rings for B streams + the device hop counter.
class StateArena {
 public:
  StateArena() = default;
  ~StateArena() = default;
  // two-phase: Plan() rings, then Commit() allocates and patches the base pointers
  int Plan(int history_rows, int T, int C);  // fp32 ring; returns ring id
  int PlanH(int history_rows, int T, int C, bool with_lo);  // bf16 ring (hi [+ lo] planes)
  void Commit(int device, int B);
  const Ring& ring(int id) const { return rings_[id]; }
  int* frame() const { return frame_.as<int>(); }
  void ZeroStream(int b, cudaStream_t s);
  void ZeroAll(cudaStream_t s);
  void Clear();
  size_t bytes() const { return buf_.bytes; }
  int B() const { return B_; }

 private:
  std::vector<Ring> rings_;
  std::vector<size_t> offsets_;  // byte offsets
  DeviceBuffer buf_, frame_;
  int B_ = 0;
};

// Encoder (front end + normalised residual backbone + head) for B streams.
struct EncoderState {
  int B = 0, device = -1;
  const EncoderModel* model = nullptr;
  uint64_t model_generation = 0;
  StateArena arena;
  DeviceBuffer descs;     // ConvDesc table
  DeviceBuffer in_stage;  // [B][160]
  DeviceBuffer head_out;  // [B][head_out]
  // batched engine: where the head output is ALSO written when the head is the last block of the fused chain
  // (the vocoder's phone input: with kNN-VQ off for every stream the VQ launch, a copy then, is skipped).
  // Set before Build; head_dual: Build took the offer.
  float* head_copy = nullptr;
  bool head_dual = false;
  std::vector<Op> program;
  DeviceBuffer rs_hist, rs_blocks;   // fused residual stack: conv-input histories + their per-stream reset table
  int n_rs_blocks = 0;
  void ZeroStream(int b, cudaStream_t s);   // rings + fused-stack histories of stream b
  void ZeroAll(cudaStream_t s);
  const float* stage_ptr = nullptr;  // == in_stage unless an external staging buffer is shared
  // Builds rings + program for `m`.  `external_stage` (device, [B][160]) replaces in_stage.
  void Build(const EncoderModel* m, int B, int device, const float* external_stage = nullptr,
             TcMode tc = kTcOff);
  bool Matches(const EncoderModel* m) const { return model == m && m && model_generation == m->generation; }
};

struct WaveState {
  int B = 0, device = -1;
  const WaveModel* model = nullptr;
  uint64_t model_generation = 0;
  StateArena arena;
  DeviceBuffer descs;
  DeviceBuffer phone_in;  // [B][P]
  DeviceBuffer q_in;      // [B] int
  DeviceBuffer feat_in;   // [B][4]
  DeviceBuffer spk;       // [B][256]   rc0: projected additive embedding; a2/b1: per-call vector
  DeviceBuffer formant;   // [B][256]   rc0 only
  DeviceBuffer film[4];   // [B][2*C_s] rc0 only (zero = identity)
  DeviceBuffer out;       // [B][240]
  std::vector<Op> program;
  // Upsamplers of the stages whose fused MRF kernel can compute them in its prologue (MrfUpsDesc).  The choice is made
  // when a hop is ENQUEUED (read by the ops' launchers, i.e. at graph capture): in the prologue for the latency path
  // (one launch and one kernel boundary less per stage on the hop's critical path), as launches of their own at pipeline
  // depth 2, where the hop is bound by SM time and the prologue form computes every upsampler once per branch CTA.
  std::shared_ptr<int> ups_in_prologue;   // bit s: stage s's upsampler runs in the prologue of its MRF kernel
  int fusable_ups_mask = 0;               // stages that have the prologue form ("wave.ups<s>" launches nothing while selected)
  int LaunchesPerHop() const {
    const int m = ups_in_prologue ? (*ups_in_prologue & fusable_ups_mask) : 0;
    return static_cast<int>(program.size()) - ((m >> 1) & 1) - ((m >> 2) & 1) - ((m >> 3) & 1);
  }
  int ring_hidden = -1, ring_pre = -1, ring_stage_out[4][3];
  bool cond_ready = false;
  // batched engine: the post conv, last kernel of a hop, advances this state's hop counter and the two encoders'
  // (set before Build; see AdvanceFold in b200_kernels.h).  advance_folded: Build took the offer.
  int* fold_frames[2] = {nullptr, nullptr};
  bool advance_folded = false;
  DeviceBuffer fold_done;
  std::function<void(cudaStream_t)> post_own_advance;   // "wave.post" advancing only this state's hop counter
  // fused MRF stages: conv-input histories (bf16, stream-group layout) + reset table
  DeviceBuffer mrf_hist, mrf_blocks;
  int n_mrf_blocks = 0;
  DeviceBuffer pre_hist, pre_blocks;   // fused pre conv: its six-row input history + per-stream reset table
  int n_pre_blocks = 0;
  void ZeroStream(int b, cudaStream_t s);   // arena + fused-kernel histories of stream b
  void ZeroAll(cudaStream_t s);             // ... of every stream
  // conditioning buffers depend only on the family, not on the weights: the rc0 setters
  // (beatrice.h:323-343) may run before the first GenerateWaveform1 names the model
  void AllocCond(const FamilyDims& dims, int B, int device);
  void Build(const WaveModel* m, int B, int device, TcMode tc = kTcOff);
  bool Matches(const WaveModel* m) const { return model == m && m && model_generation == m->generation; }
};

// Enqueues every op; the caller accounts launches in g_kernel_launches per executed hop.
void RunProgram(const std::vector<Op>& program, cudaStream_t s);

// A program captured once into a CUDA graph and replayed per hop.
class GraphRunner {
 public:
  ~GraphRunner();
  void Reset();
  // Runs body either directly or as a (lazily captured) graph on `s`.
  void Run(cudaStream_t s, const std::function<void(cudaStream_t)>& body, bool use_graph);

 private:
  cudaGraphExec_t exec_ = nullptr;
};

extern std::atomic<uint64_t> g_kernel_launches;  // launches issued through Op programs

}  // namespace b200

#endif  // BEATRICE_B200_ENGINE_H_
