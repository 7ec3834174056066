// Kernel-facing descriptors and launchers of libbeatrice_b200 (sm_100a).
//
// Data layout in HBM: every activation that a causal convolution
// needs history of lives in a RING of rows, channel-last:
//     ring[b][slot * T + t][c]      b < B streams, slot < slots, t < T rows per 10 ms hop
// All streams of a batch advance together, so one device-resident hop counter `frame`
// selects the slot (slot = frame % slots) for every ring of that batch; nothing about a
// launch changes from hop to hop, which is what lets a whole hop be one CUDA graph.
// The producer's output rows ARE the consumer's history: no shifting, no second copy.
#ifndef BEATRICE_B200_KERNELS_H_
#define BEATRICE_B200_KERNELS_H_

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace b200 {

enum Act : int { kActNone = 0, kActLrelu = 1, kActGelu = 2, kActTanh = 3 };

// One causal Conv1d (also strided, also ConvTranspose1d-as-2-tap-conv) as an implicit GEMM
//   Y[(b,t)][n] = epilogue( sum_{j<k} sum_{ci} act_in(X[b][u(t,j)][ci]) * W[j][ci][n] )
//   u(t,j) = t*stride + (stride-1) - (k-1-j)*dil     (negative u reaches into older slots)
struct ConvDesc {
  const float* x[3];  // up to three input rings of identical geometry, summed, times in_scale
  int n_x;
  float in_scale;
  int x_slots, x_T, x_C;  // x_C == C_in
  int in_act;
  const float* w;     // [k][C_in][N]
  const float* bias;  // [N] or nullptr
  int k, dil, stride, C_in, N;
  int T;              // GEMM rows (output steps) per stream per hop
  float* y;           // output ring; row (b,t) stores N contiguous floats (y_T*y_C == T*N)
  int y_slots, y_T, y_C;
  const float* res;   // residual ring with N channels (current-hop rows), or nullptr
  int res_slots, res_T;
  const float* film;  // [B][2*film_C] = gamma | beta, applied as v*(1+gamma)+beta, or nullptr
  int film_C;
  int out_act;
  // tcgen05 path (b200_tc.cu): weights re-packed as bf16 UMMA operand tiles, see PackWeightsTc
  const void* w_tc;   // nullptr -> this conv has no tensor-core form
  const void* w_tc_lo;  // low-order bf16 residual of the weights (split-bf16 mode) or nullptr
  int tc_bn;          // N tile width (multiple of 16, <= 256)
  int tc_kc;          // K elements per chunk (taps_per_chunk * C_in or 64)
  // bf16 activation rings of the tensor-core path.  xh/xl: the input already activated and
  // rounded to bf16 (hi) plus the bf16 of the rounding residual (lo, split mode), same ring
  // geometry as x[]; when set, activation rows are moved global -> shared with cp.async and
  // never touch registers.  yh/yl: where the epilogue stores bf16(yh_act(v)) for the consumer.
  const uint16_t* xh;
  const uint16_t* xl;
  uint16_t* yh;
  uint16_t* yl;
  int yh_slots;
  int yh_act;
  int trace;          // developer aid (BEATRICE_B200_TC_TRACE=<op name>): CTA (0,0,0) prints its timeline
};

struct NormDesc {  // y = GELU(ChanNorm(x) * gamma + beta), one row per warp
  const float* x;
  int x_slots, T, C;
  const float* gamma;
  const float* beta;
  float* y;
  int y_slots;
  uint16_t* yh;  // optional bf16 copy (hi / lo planes) for the tensor-core consumer
  uint16_t* yl;
  int yh_slots;
};

// Per-stream pitch-transform parameters == the ProcessorCore2 members of the same name
// (reference src/common/processor_core_2.h:103-113).
struct PitchParams {
  double average_source_pitch;
  double intonation_intensity;
  double pitch_shift;
  double pitch_correction;
  int pitch_correction_type;
  int pad_;
};

// ---- launchers (all asynchronous on `s`) ----
// descs: device array of nz descriptors sharing N, T, C_in geometry class; h0 = host copy of descs[0]
void LaunchConvGemm(const ConvDesc* d_descs, const ConvDesc& h0, int nz, int B, const int* d_frame,
                    cudaStream_t s);
// Same contraction on the 5th-gen tensor cores: bf16 operands staged in shared memory (weights
// by TMA bulk copy), fp32 accumulators in TMEM, same fused epilogue.  split = true adds the two
// cross terms of a hi/lo bf16 decomposition of both operands (near-fp32 accuracy, 3 MMAs).
void LaunchConvGemmTc(const ConvDesc* d_descs, const ConvDesc& h0, int nz, int B, const int* d_frame, bool split,
                      cudaStream_t s);
// Host-side packing of one conv's weights w[k][C_in][N] (fp32) into the tile order the kernel
// consumes; returns bytes written per array.  hi/lo may be nullptr to query the size.
size_t PackWeightsTc(const float* w, int k, int C_in, int N, int bn_cap, int* bn_out, int* kc_out, uint16_t* hi,
                     uint16_t* lo);
// post conv of the vocoder (16 -> 1 channels, k = 7, tanh): one thread per output sample
// Hop counters the LAST kernel of a hop advances itself (its last block to finish, when every block has long
// read them), instead of one single-thread advance launch per counter; done == nullptr: nothing to advance.
struct AdvanceFold {
  int* done = nullptr;                               // zero-initialised block counter, left at zero again
  int* frames[3] = {nullptr, nullptr, nullptr};
};
void LaunchPostConv(const ConvDesc* d_desc, const ConvDesc& h0, int B, const int* d_frame, const AdvanceFold& fold,
                    cudaStream_t s);
bool PostConvFused(const ConvDesc& h0);   // the shape post_conv_kernel handles (else the generic direct conv runs)
void LaunchDirectConv(const ConvDesc* d_desc, const ConvDesc& h0, int B, const int* d_frame, cudaStream_t s);
// encoder front-end layer 0 (C_in = 1) fused with the hop's ingest: staging [B][x_T] -> ring slot + conv
bool Frontend0Supported(const ConvDesc& h0);
void LaunchFrontend0(const ConvDesc& h0, const float* staging, float* ring, int B, const int* d_frame, cudaStream_t s);
void LaunchNorm(const NormDesc& d, int B, const int* d_frame, cudaStream_t s);
// staging [B][T*C] -> ring slot of the current hop
void LaunchIngest(const float* staging, float* ring, int slots, int T, int C, int B, const int* d_frame,
                  cudaStream_t s);
void LaunchAdvance(int* d_frame, cudaStream_t s);
// head [B][bins+4] -> q [B] (arg-max over [min_q[b], max_q[b]]), feat [B][4]
// params / q_used non-null: also applies the call site's pitch transform (PitchParams per stream) in the same launch
void LaunchPitchArgmax(const float* head, int bins, const int* min_q, const int* max_q, int* q, float* feat,
                       int B, cudaStream_t s, const PitchParams* params = nullptr, int* q_used = nullptr);
// reference call-site transform, fp64 (processor_core_2.cc:190-252)
// Voice morphing (b200_morph.cu): one stream's pruned, arg-sorted morphing weights (processor_core_2.cc:507-532) and
// where its averages go.  Item i of a job averages row item0 + i of the speakers idx[0..n).
struct MorphJob {
  int n;          // speakers with non-zero weight, <= 8 (kSphAvgMaxNSpeakers)
  int item0;      // first row (key-value embedding: 96 rows per frame; additive embedding: 0)
  int dst_row;    // destination slot, in units of dst_stride
  int pad;
  int idx[8];     // speaker ids, by descending weight
  float w[8];     // their weights (not normalised)
};
void LaunchSphAvg(int M, const float* table, long long speaker_stride, const MorphJob* d_jobs, int n_jobs,
                  int items_per_job, float* dst, long long dst_stride, cudaStream_t s);
void LaunchPitchTransform(const int* q_in, const PitchParams* params, int bins, int* q_out, int B,
                          cudaStream_t s);
// conditioning: phone 1x1 + pitch embedding gather + feature projection + speaker (+ formant)
void LaunchCond(const float* phone, int P, const int* q, int bins, const float* feat, const float* We,
                const float* be, const float* pitch_emb, const float* Wf, const float* spk /*[B][256]|null*/,
                const float* formant /*[B][256]|null*/, float* ring, uint16_t* ring_hi /*null|bf16 copy*/,
                uint16_t* ring_lo, int slots, int B, const int* d_frame,
                cudaStream_t s);
// kNN-VQ: phone_in [B][C] -> phone_out [B][C]; codebooks[b] -> 512 x C device table (or null), n[b] neighbours
void LaunchVq(const float* phone_in, float* phone_out, const float* const* codebooks, const int* n_neighbors,
              int C, int B, cudaStream_t s);
// y[row][256] = b + e[src] . W   (W stored [in][out]); src = e_index ? e_index[item] : item,
// row = out_index ? out_index[item] : item
void LaunchProject256(const float* W, const float* b, const float* e, size_t e_stride, const int* e_index,
                      float* out, const int* out_index, int n_items, cudaStream_t s);
// attention-pool kv (384 x 128) with `query`, then film = b + pooled . W  (W [128][2C])
// item i reads kv_base + kv_index[i]*kv_stride (kv_index null -> 0) and writes film_base +
// (out_index ? out_index[i] : i) * 2C
// Setter launches whose (source row, destination stream[, block]) lists travel BY VALUE in the kernel parameters: a
// parameter change touches a handful of streams, and an index upload per launch (a pageable cudaMemcpyAsync: host
// staging + a copy operation in front of the kernel) costs more than the kernel.  Longer lists go out in chunks.
constexpr int kSetterItems = 48;
struct SetterItems {
  int n;
  int src[kSetterItems];   // row of the table read (speaker / formant index / key-value slot)
  int dst[kSetterItems];   // stream written
  int blk[kSetterItems];   // key-value block = vocoder stage (LaunchKvFilmItems); kNN-VQ neighbour count (LaunchVqPatchItems)
};
struct KvBlockTable {      // EmbeddingSetter block parameters, by value
  const float* query[4];
  const float* W[4];
  const float* bias[4];
  float* film[4];          // [B][2 C_blk]
  int C[4];
};
void LaunchProject256Items(const float* W, const float* b, const float* e, size_t e_stride, float* out, const SetterItems& items,
                           cudaStream_t s);
void LaunchKvFilmItems(const float* kv_base, size_t kv_stride, const KvBlockTable& tab, const SetterItems& items, cudaStream_t s);
// vq_n[dst] = blk, codebook_ptrs[dst] = blk > 0 || src >= 0 ? codebooks + cb_stride * src : nullptr
void LaunchVqPatchItems(int* vq_n, const float** codebook_ptrs, const float* codebooks, size_t cb_stride, const SetterItems& items,
                        cudaStream_t s);
void LaunchKvFilm(const float* kv_base, const int* kv_index, size_t kv_stride, const float* query,
                  const float* W, const float* b, int C, float* film_base, const int* out_index, int n_items,
                  cudaStream_t s);
void LaunchFill(float* p, float v, size_t n, cudaStream_t s);

}  // namespace b200

#endif  // BEATRICE_B200_KERNELS_H_
