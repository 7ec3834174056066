// fp32 CUDA kernels of libbeatrice_b200 (sm_100a): the parity path.
//
// conv_gemm_kernel is the workhorse: every Conv1d / strided Conv1d / ConvTranspose1d of the
// content encoder, pitch estimator and HiFi-GAN-style vocoder (spec M0)
// runs through it as an implicit GEMM over channel-last ring buffers, with the input
// activation, bias, FiLM, residual add and output activation fused.  The bf16 tcgen05
// variant of the same contraction lives in b200_tc.cu.
#include <cuda_bf16.h>

#include <cfloat>
#include <climits>

#include "b200_common.h"
#include "b200_kernels.h"

namespace b200 {
namespace {

// The transcendental activations are kept out of line: inlining erff/tanhf at every use made
// the conv kernels > 100 KB of SASS and instruction-fetch bound (ncu: stall "no_instructions").
__device__ __noinline__ float SlowAct(float v, int act) {
  if (act == kActGelu) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
  return tanhf(v);
}
// inline GELU for the tensor-core modes (same form as b200_tc_common.cuh): erf by Abramowitz & Stegun
// 7.1.26, |error| <= 1.5e-7; the fp32 parity mode keeps the exact erff above
__device__ __forceinline__ float GeluFastK(float v) {
  const float x = fabsf(v) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-x * x);
  return 0.5f * v * (1.0f + copysignf(e, v));
}
__device__ __forceinline__ float ActApply(float v, int act) {
  if (act == kActNone) return v;
  if (act == kActLrelu) return v > 0.0f ? v : 0.1f * v;
  return SlowAct(v, act);
}
// input side: only "none" and LeakyReLU occur (spec M0 applies GELU/tanh on the output side)
__device__ __forceinline__ float4 InAct4(float4 v, int act) {
  if (act == kActLrelu) {
    v.x = v.x > 0.0f ? v.x : 0.1f * v.x;
    v.y = v.y > 0.0f ? v.y : 0.1f * v.y;
    v.z = v.z > 0.0f ? v.z : 0.1f * v.z;
    v.w = v.w > 0.0f ? v.w : 0.1f * v.w;
  }
  return v;
}
__device__ __forceinline__ float4 ActApply4(float4 v, int act) {
  if (act == kActLrelu) return InAct4(v, act);
  if (act != kActNone) {
    v.x = SlowAct(v.x, act);
    v.y = SlowAct(v.y, act);
    v.z = SlowAct(v.z, act);
    v.w = SlowAct(v.w, act);
  }
  return v;
}
__device__ __forceinline__ float4 Ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------------------------------
// implicit-GEMM causal conv.  CTA tile BM x BN, 128 threads, 8x4 accumulators per thread,
// K walked tap by tap in chunks of 16 input channels, register-prefetched double buffering.
// ------------------------------------------------------------------------------------------
template <int BM, int BN>
__global__ void __launch_bounds__(128) conv_gemm_kernel(const ConvDesc* __restrict__ descs, int B,
                                                        const int* __restrict__ frame_ptr) {
  constexpr int TM = 8, TN = 4, BK = 16;
  constexpr int TX = BN / TN;  // threads along N
  static_assert((BM / TM) * TX == 128, "tile must map onto 128 threads");
  constexpr int XLD = BM + 4;
  constexpr int XV = BM / 32;                         // float4 activation loads / thread / chunk
  constexpr int WV = (BK * BN / 4 + 127) / 128;       // float4 weight loads / thread / chunk
  __shared__ __align__(16) float Xs[2][BK][XLD];
  __shared__ __align__(16) float Ws[2][BK][BN];

  const ConvDesc d = descs[blockIdx.z];
  const int frame = *frame_ptr;
  const int M = B * d.T;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
  const int C_in = d.C_in, N = d.N;

  const int x_L = d.x_slots * d.x_T;
  const int x_cur = (frame % d.x_slots) * d.x_T;
  const int q4 = (tid & 3) * 4;  // which 4 of the 16 chunk channels this thread stages

  // rows this thread stages: row = (tid>>2) + 32*i
  long long xbase[XV];  // element offset of stream b's ring, or -1 for rows past M
  int xu0[XV];          // t*stride + stride-1
#pragma unroll
  for (int i = 0; i < XV; ++i) {
    const int m = m0 + (tid >> 2) + 32 * i;
    if (m < M) {
      const int b = m / d.T, t = m - b * d.T;
      xbase[i] = static_cast<long long>(b) * x_L * C_in;
      xu0[i] = t * d.stride + d.stride - 1;
    } else {
      xbase[i] = -1;
      xu0[i] = 0;
    }
  }

  const int nck = C_in / BK;
  const int nchunks = d.k * nck;

  float4 xr[XV];
  float4 wr[WV];

  auto load_chunk = [&](int c) {
    const int j = c / nck;
    const int ci0 = (c - j * nck) * BK;
    const int off = (d.k - 1 - j) * d.dil;
#pragma unroll
    for (int i = 0; i < XV; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (xbase[i] >= 0) {
        int r = x_cur + xu0[i] - off;
        if (r < 0) r += x_L;
        const long long a = xbase[i] + static_cast<long long>(r) * C_in + ci0 + q4;
        v = Ldg4(d.x[0] + a);
        if (d.n_x > 1) {
          const float4 v1 = Ldg4(d.x[1] + a);
          const float4 v2 = Ldg4(d.x[2] + a);
          v.x = ((v.x + v1.x) + v2.x) * d.in_scale;
          v.y = ((v.y + v1.y) + v2.y) * d.in_scale;
          v.z = ((v.z + v1.z) + v2.z) * d.in_scale;
          v.w = ((v.w + v1.w) + v2.w) * d.in_scale;
        }
        v = InAct4(v, d.in_act);
      }
      xr[i] = v;
    }
#pragma unroll
    for (int i = 0; i < WV; ++i) {
      const int idx = tid + i * 128;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < BK * BN / 4) {
        const int kk = idx / (BN / 4), c4 = idx % (BN / 4);
        const int col = n0 + c4 * 4;
        if (col < N) v = Ldg4(d.w + (static_cast<long long>(j) * C_in + ci0 + kk) * N + col);
      }
      wr[i] = v;
    }
  };
  auto store_chunk = [&](int buf) {
#pragma unroll
    for (int i = 0; i < XV; ++i) {
      const int row = (tid >> 2) + 32 * i;
      Xs[buf][q4 + 0][row] = xr[i].x;
      Xs[buf][q4 + 1][row] = xr[i].y;
      Xs[buf][q4 + 2][row] = xr[i].z;
      Xs[buf][q4 + 3][row] = xr[i].w;
    }
#pragma unroll
    for (int i = 0; i < WV; ++i) {
      const int idx = tid + i * 128;
      if (idx < BK * BN / 4) {
        const int kk = idx / (BN / 4), c4 = idx % (BN / 4);
        *reinterpret_cast<float4*>(&Ws[buf][kk][c4 * 4]) = wr[i];
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) load_chunk(c + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&Xs[buf][kk][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&Xs[buf][kk][ty * TM + 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * TN]);
      const float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[TN] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (c + 1 < nchunks) store_chunk(buf ^ 1);
    __syncthreads();
  }

  // ---- fused epilogue: bias -> FiLM -> residual -> activation -> ring store ----
  const int col = n0 + tx * TN;
  if (col >= N) return;
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (d.bias) bias4 = Ldg4(d.bias + col);
  const int y_L = d.y_slots * d.y_T;
  const int y_cur = (frame % d.y_slots) * d.y_T;
  const int res_L = d.res_slots * d.res_T;
  const int res_cur = d.res ? (frame % d.res_slots) * d.res_T : 0;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) break;
    const int b = m / d.T, t = m - b * d.T;
    float4 v = make_float4(acc[i][0] + bias4.x, acc[i][1] + bias4.y, acc[i][2] + bias4.z, acc[i][3] + bias4.w);
    if (d.film) {
      const int fc = col % d.film_C;
      const float* f = d.film + static_cast<long long>(b) * 2 * d.film_C;
      const float4 g = Ldg4(f + fc), be = Ldg4(f + d.film_C + fc);
      v.x = v.x * (1.0f + g.x) + be.x;
      v.y = v.y * (1.0f + g.y) + be.y;
      v.z = v.z * (1.0f + g.z) + be.z;
      v.w = v.w * (1.0f + g.w) + be.w;
    }
    if (d.res) {
      const float4 r = Ldg4(d.res + (static_cast<long long>(b) * res_L + res_cur + t) * N + col);
      v.x += r.x;
      v.y += r.y;
      v.z += r.z;
      v.w += r.w;
    }
    v = ActApply4(v, d.out_act);
    float* out = d.y + (static_cast<long long>(b) * y_L + y_cur) * d.y_C + static_cast<long long>(t) * N + col;
    *reinterpret_cast<float4*>(out) = v;
  }
}

// One thread per output element; for the two layers that are not GEMM shaped
// (front-end conv with C_in = 1, post conv with C_out = 1).
__global__ void direct_conv_kernel(const ConvDesc* __restrict__ descs, int B, const int* __restrict__ frame_ptr) {
  PdlWait();
  PdlLaunchDependents();
  const ConvDesc d = descs[0];
  const int frame = *frame_ptr;
  const long long total = static_cast<long long>(B) * d.T * d.N;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = static_cast<int>(idx % d.N);
  const int m = static_cast<int>(idx / d.N);
  const int b = m / d.T, t = m - b * d.T;
  const int x_L = d.x_slots * d.x_T;
  const int x_cur = (frame % d.x_slots) * d.x_T;
  const long long xb = static_cast<long long>(b) * x_L * d.C_in;
  float acc = d.bias ? __ldg(d.bias + n) : 0.f;
  for (int j = 0; j < d.k; ++j) {
    int r = x_cur + t * d.stride + d.stride - 1 - (d.k - 1 - j) * d.dil;
    if (r < 0) r += x_L;
    const long long a = xb + static_cast<long long>(r) * d.C_in;
    const float* wj = d.w + static_cast<long long>(j) * d.C_in * d.N + n;
    for (int ci = 0; ci < d.C_in; ++ci) {
      float v = __ldg(d.x[0] + a + ci);
      if (d.n_x > 1) v = ((v + __ldg(d.x[1] + a + ci)) + __ldg(d.x[2] + a + ci)) * d.in_scale;
      v = ActApply(v, d.in_act);
      acc = fmaf(v, __ldg(wj + static_cast<long long>(ci) * d.N), acc);
    }
  }
  if (d.res) {
    const int res_L = d.res_slots * d.res_T;
    const int res_cur = (frame % d.res_slots) * d.res_T;
    acc += __ldg(d.res + (static_cast<long long>(b) * res_L + res_cur + t) * d.N + n);
  }
  acc = ActApply(acc, d.out_act);
  const int y_L = d.y_slots * d.y_T;
  const int y_cur = (frame % d.y_slots) * d.y_T;
  if (d.y) d.y[(static_cast<long long>(b) * y_L + y_cur) * d.y_C + static_cast<long long>(t) * d.N + n] = acc;
  if (d.yh) {   // bf16 copy (hi [+ lo] planes) for a tensor-core consumer
    const int yh_L = d.yh_slots * d.y_T;
    const int yh_cur = (frame % d.yh_slots) * d.y_T;
    const long long o = (static_cast<long long>(b) * yh_L + yh_cur) * d.y_C + static_cast<long long>(t) * d.N + n;
    const float hv = ActApply(acc, d.yh_act);
    const __nv_bfloat16 h = __float2bfloat16_rn(hv);
    d.yh[o] = __bfloat16_as_ushort(h);
    if (d.yl) d.yl[o] = __bfloat16_as_ushort(__float2bfloat16_rn(hv - __bfloat162float(h)));
  }
}

// Vocoder post conv: out[b][t] = tanh(b0 + sum_{j<7} sum_{ci<16} lrelu(x[b][t-(6-j)][ci]) * w[j][ci]) with
// x = (x0 + x1 + x2) * in_scale (the three MRF branches).  One block per stream: the hop's rows plus six
// rows of history are combined, activated and staged in shared memory once (coalesced float4 loads),
// then one thread per output sample walks the 7 x 16 window.  Summation order is the oracle's (tap,
// then channel), so the result does not depend on the staging.
constexpr int kPostRowLd = 20;   // floats per staged row (16 + 4: conflict-free 16-byte reads, rows 80 B apart)
__device__ __forceinline__ int NextFrame(int f) { return (f + 1 >= 738017280) ? 0 : f + 1; }   // see advance_kernel

__global__ void __launch_bounds__(256) post_conv_kernel(const __grid_constant__ ConvDesc d, int B,
                                                         const int* __restrict__ frame_ptr, const AdvanceFold fold) {
  extern __shared__ float post_smem[];
  float* ws = post_smem;                 // [7][16]
  float* xs = post_smem + 7 * 16;        // [T + 6][kPostRowLd]
  const int tid = threadIdx.x, b = blockIdx.x;
  if (tid < 7 * 16) ws[tid] = __ldg(d.w + tid);            // constants: before the dependency wait
  const float bias = d.bias ? __ldg(d.bias) : 0.f;
  const int frame = *frame_ptr;                             // written only by the chain's advance kernel
  PdlWait();
  PdlLaunchDependents();
  const int x_L = d.x_slots * d.x_T;
  const int x_cur = (frame % d.x_slots) * d.x_T;
  const long long xb = static_cast<long long>(b) * x_L * 16;
  // T = 240: four rounds per thread.  The branch rings are read through unconditional pointers (the single-input
  // form aliases all three to the first) so that the twelve loads of a thread are in flight together.
  const bool mean3 = d.n_x > 1;
  const float* x0 = d.x[0];
  const float* x1 = mean3 ? d.x[1] : d.x[0];
  const float* x2 = mean3 ? d.x[2] : d.x[0];
  constexpr int kRounds = 4;
  float4 v0[kRounds], v1[kRounds], v2[kRounds];
  const int n_items = (d.T + 6) * 4;
#pragma unroll
  for (int j = 0; j < kRounds; ++j) {
    const int idx = tid + 256 * j;
    const int row = idx >> 2, q = idx & 3;
    int r = x_cur + row - 6;
    if (r < 0) r += x_L;
    const long long a = xb + static_cast<long long>(r) * 16 + 4 * q;
    const bool ok = idx < n_items;
    v0[j] = ok ? Ldg4(x0 + a) : make_float4(0.f, 0.f, 0.f, 0.f);
    v1[j] = ok ? Ldg4(x1 + a) : make_float4(0.f, 0.f, 0.f, 0.f);
    v2[j] = ok ? Ldg4(x2 + a) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int j = 0; j < kRounds; ++j) {
    const int idx = tid + 256 * j;
    if (idx >= n_items) break;
    const int row = idx >> 2, q = idx & 3;
    float4 v = v0[j];
    if (mean3) {
      v.x = ((v.x + v1[j].x) + v2[j].x) * d.in_scale;
      v.y = ((v.y + v1[j].y) + v2[j].y) * d.in_scale;
      v.z = ((v.z + v1[j].z) + v2[j].z) * d.in_scale;
      v.w = ((v.w + v1[j].w) + v2[j].w) * d.in_scale;
    }
    v = InAct4(v, d.in_act);
    *reinterpret_cast<float4*>(xs + row * kPostRowLd + 4 * q) = v;
  }
  for (int idx = tid + 256 * kRounds; idx < n_items; idx += 256) {   // T > 250 (not spec M0): plain rounds
    const int row = idx >> 2, q = idx & 3;
    int r = x_cur + row - 6;
    if (r < 0) r += x_L;
    const long long a = xb + static_cast<long long>(r) * 16 + 4 * q;
    float4 v = Ldg4(x0 + a);
    if (mean3) {
      const float4 w1 = Ldg4(x1 + a), w2 = Ldg4(x2 + a);
      v.x = ((v.x + w1.x) + w2.x) * d.in_scale;
      v.y = ((v.y + w1.y) + w2.y) * d.in_scale;
      v.z = ((v.z + w1.z) + w2.z) * d.in_scale;
      v.w = ((v.w + w1.w) + w2.w) * d.in_scale;
    }
    v = InAct4(v, d.in_act);
    *reinterpret_cast<float4*>(xs + row * kPostRowLd + 4 * q) = v;
  }
  __syncthreads();
  const int y_L = d.y_slots * d.y_T;
  const int y_cur = (frame % d.y_slots) * d.y_T;
  for (int t = tid; t < d.T; t += 256) {
    float acc = bias;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(xs + (t + j) * kPostRowLd + 4 * q);
        const float* wq = ws + j * 16 + 4 * q;
        acc = fmaf(v.x, wq[0], acc);
        acc = fmaf(v.y, wq[1], acc);
        acc = fmaf(v.z, wq[2], acc);
        acc = fmaf(v.w, wq[3], acc);
      }
    }
    acc = ActApply(acc, d.out_act);
    d.y[(static_cast<long long>(b) * y_L + y_cur) * d.y_C + t] = acc;
  }
  if (fold.done != nullptr && tid == 0) {
    // the hop ends here: the last block to arrive advances the hop counters.  Every block read them on entry
    // (its output addresses depend on the value), so no fence is needed: the last arrival publishes nothing
    // another block reads.
    if (atomicAdd(fold.done, 1) == static_cast<int>(gridDim.x) - 1) {
      *fold.done = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (fold.frames[i] != nullptr) *fold.frames[i] = NextFrame(*fold.frames[i]);
    }
  }
}

// Front-end layer 0 of the encoders (C_in = 1, k = 10, stride 5 -> N <= 32 channels) fused with the
// ingest of the hop: one block per stream copies the hop's samples staging -> ring (they are next
// hop's history), stages them with the k - stride samples of history, and computes the T x N outputs
// from shared memory.  Writes the fp32 ring and / or the bf16 (hi [+ lo]) ring of the next layer.
__global__ void __launch_bounds__(256) frontend0_kernel(const __grid_constant__ ConvDesc d, const float* __restrict__ staging,
                                                         float* __restrict__ ring, int B, const int* __restrict__ frame_ptr) {
  __shared__ float xs[kInHop + 32];
  __shared__ float ws[16 * 32];
  __shared__ float bs[32];
  const int tid = threadIdx.x, b = blockIdx.x;
  for (int i = tid; i < d.k * d.N; i += 256) ws[i] = __ldg(d.w + i);   // constants: before the dependency wait
  if (tid < d.N) bs[tid] = d.bias ? __ldg(d.bias + tid) : 0.f;
  PdlWait();
  PdlLaunchDependents();
  // first kernel of a hop: its stream predecessor may be the previous hop's last kernel, which advances the
  // counter when the advance is folded (AdvanceFold) -- so read it only after the dependency wait
  const int frame = *frame_ptr;
  const int hist = d.k - d.stride;
  const int x_L = d.x_slots * d.x_T;
  const int x_cur = (frame % d.x_slots) * d.x_T;
  float* rb = ring + static_cast<long long>(b) * x_L;
  for (int i = tid; i < d.x_T; i += 256) {
    const float v = staging[static_cast<long long>(b) * d.x_T + i];
    xs[hist + i] = v;
    rb[x_cur + i] = v;
  }
  if (tid < hist) {
    int r = x_cur - hist + tid;
    if (r < 0) r += x_L;
    xs[tid] = rb[r];
  }
  __syncthreads();
  const int y_L = d.y_slots * d.y_T;
  const int y_cur = (frame % d.y_slots) * d.y_T;
  const int yh_L = d.yh_slots * d.y_T;
  const int yh_cur = d.yh ? (frame % d.yh_slots) * d.y_T : 0;
  for (int item = tid; item < d.T * d.N; item += 256) {
    const int t = item / d.N, n = item - t * d.N;
    float acc = bs[n];
    for (int j = 0; j < d.k; ++j) acc = fmaf(xs[t * d.stride + j], ws[j * d.N + n], acc);
    acc = (d.yh && d.out_act == kActGelu) ? GeluFastK(acc) : ActApply(acc, d.out_act);   // exact erff on the fp32 parity path
    if (d.y) d.y[(static_cast<long long>(b) * y_L + y_cur) * d.y_C + static_cast<long long>(t) * d.N + n] = acc;
    if (d.yh) {
      const long long o = (static_cast<long long>(b) * yh_L + yh_cur) * d.y_C + static_cast<long long>(t) * d.N + n;
      const float hv = ActApply(acc, d.yh_act);
      const __nv_bfloat16 h = __float2bfloat16_rn(hv);
      d.yh[o] = __bfloat16_as_ushort(h);
      if (d.yl) d.yl[o] = __bfloat16_as_ushort(__float2bfloat16_rn(hv - __bfloat162float(h)));
    }
  }
}

__device__ __forceinline__ float WarpSum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// y = GELU(ChanNorm(x)*gamma+beta); one warp per (stream, row); channel statistics by
// warp-shuffle reduction, two-pass like the oracle (mean, then centred variance).
__global__ void channorm_gelu_kernel(NormDesc d, int B, const int* __restrict__ frame_ptr) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  // gamma / beta are constants: fetched before the dependency wait
  float gm[8], bt[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    gm[i] = c < d.C ? __ldg(d.gamma + c) : 0.0f;
    bt[i] = c < d.C ? __ldg(d.beta + c) : 0.0f;
  }
  const int frame = *frame_ptr;   // written only by the chain's advance kernel, never the immediate predecessor
  PdlWait();
  PdlLaunchDependents();
  if (warp >= B * d.T) return;
  const int b = warp / d.T, t = warp - b * d.T;
  const float* x = d.x + (static_cast<long long>(b) * d.x_slots * d.T + (frame % d.x_slots) * d.T + t) * d.C;
  float* y = d.y + (static_cast<long long>(b) * d.y_slots * d.T + (frame % d.y_slots) * d.T + t) * d.C;
  constexpr int kMaxPerLane = 8;  // C <= 256
  float v[kMaxPerLane];
  const int per = d.C / 32;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i)
    if (i < per) {
      v[i] = x[lane + 32 * i];
      s += v[i];
    }
  const float mean = WarpSum(s) / static_cast<float>(d.C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i)
    if (i < per) {
      const float dl = v[i] - mean;
      q = fmaf(dl, dl, q);
    }
  const float var = WarpSum(q) / static_cast<float>(d.C);
  const float rstd = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i)
    if (i < per) {
      const int c = lane + 32 * i;
      const float z = (v[i] - mean) * rstd * gm[i] + bt[i];
      const float o = d.yh ? GeluFastK(z) : ActApply(z, kActGelu);   // exact erff only on the fp32 parity path
      if (d.y) y[c] = o;
      if (d.yh) {
        const long long off = (static_cast<long long>(b) * d.yh_slots * d.T + (frame % d.yh_slots) * d.T + t) * d.C + c;
        const __nv_bfloat16 h = __float2bfloat16_rn(o);
        d.yh[off] = __bfloat16_as_ushort(h);
        if (d.yl) d.yl[off] = __bfloat16_as_ushort(__float2bfloat16_rn(o - __bfloat162float(h)));
      }
    }
}

__global__ void ingest_kernel(const float* __restrict__ staging, float* __restrict__ ring, int slots, int TC,
                              long long total, const int* __restrict__ frame_ptr) {
  PdlWait();
  PdlLaunchDependents();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int frame = *frame_ptr;
  const long long b = idx / TC;
  const int e = static_cast<int>(idx - b * TC);
  ring[(b * slots + (frame % slots)) * TC + e] = staging[idx];
}

__global__ void advance_kernel(int* frame) {
  PdlWait();
  PdlLaunchDependents();
  // wrap at a multiple of every slot count in use (slots <= 64 by construction: lcm-free
  // choice 2^20 * 3*5*7*9*11*13 would overflow; slots are recomputed modulo so any wrap point
  // that is a common multiple works -- use 720720 * 1024 (lcm(1..16) * 1024) < 2^31)
  *frame = NextFrame(*frame);
}

// reference src/common/processor_core_2.cc:190-252, evaluated in fp64 with explicit
// round-to-nearest mul/add so that no FMA contraction changes a rounding the CPU makes.
__device__ int PitchTransformOne(int q_raw, const PitchParams& p, int bins) {
  constexpr double kBps = 96.0 / 12.0;  // BEATRICE_PITCH_BINS_PER_OCTAVE / 12
  const double q = static_cast<double>(q_raw);
  double tmp = __dadd_rn(__dadd_rn(p.average_source_pitch,
                                   __dmul_rn(__dadd_rn(q, -p.average_source_pitch), p.intonation_intensity)),
                         __dmul_rn(kBps, p.pitch_shift));
  if (p.pitch_correction != 0.0) {
    if (p.pitch_correction_type == 0) {
      const double nearest = __dmul_rn(__dadd_rn(floor(tmp / kBps), 0.5), kBps);
      const double nd = __dmul_rn(__dadd_rn(tmp, -nearest), 2.0 / kBps);
      if (fabs(nd) < 1e-4) {
        tmp = nearest;
      } else {
        tmp = __dadd_rn(nearest, __dmul_rn(__dmul_rn(nd, pow(fabs(nd), -p.pitch_correction)), kBps / 2.0));
      }
    } else if (p.pitch_correction_type == 1) {
      const double nearest = __dmul_rn(round(tmp / kBps), kBps);
      const double nd = __dmul_rn(__dadd_rn(tmp, -nearest), 2.0 / kBps);
      if (p.pitch_correction > 1 - 1e-4) {
        tmp = nearest;
      } else if (nd >= 0.0) {
        tmp = __dadd_rn(nearest, __dmul_rn(pow(nd, 1.0 / (1.0 - p.pitch_correction)), kBps / 2.0));
      } else {
        tmp = __dadd_rn(nearest, -__dmul_rn(pow(-nd, 1.0 / (1.0 - p.pitch_correction)), kBps / 2.0));
      }
    }
  }
  const double r = round(tmp);
  int qi = (r < 1.0) ? 1 : ((r > static_cast<double>(bins - 1)) ? bins - 1 : static_cast<int>(r));
  return qi;
}

__global__ void pitch_argmax_kernel(const float* __restrict__ head, int bins, const int* __restrict__ min_q,
                                    const int* __restrict__ max_q, int* __restrict__ q, float* __restrict__ feat,
                                    int B, const PitchParams* __restrict__ params, int* __restrict__ q_used) {
  PdlWait();
  PdlLaunchDependents();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* h = head + static_cast<long long>(warp) * (bins + kPitchFeatures);
  int lo = min(max(min_q[warp], 1), bins - 1);
  int hi = min(max(max_q[warp], 1), bins - 1);
  if (hi < lo) hi = lo;
  float best = -FLT_MAX;
  int besti = INT_MAX;
  for (int i = lo + lane; i <= hi; i += 32) {
    const float v = h[i];
    if (v > best) {  // strictly greater: the first maximum wins, like the oracle's scan
      best = v;
      besti = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) {
      best = ov;
      besti = oi;
    }
  }
  if (lane == 0) {
    const int qr = (besti == INT_MAX) ? lo : besti;
    q[warp] = qr;
    if (params) q_used[warp] = PitchTransformOne(qr, params[warp], bins);   // batched engine: the call site's transform, same launch
  }
  if (lane < kPitchFeatures) feat[warp * kPitchFeatures + lane] = h[bins + lane];
}

__global__ void pitch_transform_kernel(const int* __restrict__ q_in, const PitchParams* __restrict__ params,
                                       int bins, int* __restrict__ q_out, int B) {
  PdlWait();
  PdlLaunchDependents();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  q_out[b] = PitchTransformOne(q_in[b], params[b], bins);
}

// hidden[b][c] = be[c] + phone[b].We[:,c] + pitch_emb[q[b]][c] + feat[b].Wf[:,c] (+ spk + formant)
// One block = kCondStreams streams x 256 channels: every weight fetched from L2 is used for all of the
// block's streams, and the loads of a batch of 8 phone channels are in flight together.
constexpr int kCondStreams = 2;
__global__ void __launch_bounds__(256) cond_kernel(const float* __restrict__ phone, int P, const int* __restrict__ q,
                                                   int bins, const float* __restrict__ feat,
                                                   const float* __restrict__ We, const float* __restrict__ be,
                                                   const float* __restrict__ pitch_emb, const float* __restrict__ Wf,
                                                   const float* __restrict__ spk, const float* __restrict__ formant,
                                                   float* __restrict__ ring, uint16_t* __restrict__ ring_hi,
                                                   uint16_t* __restrict__ ring_lo, int slots, int B,
                                                   const int* __restrict__ frame_ptr) {
  __shared__ float ph[kCondStreams][256];
  __shared__ float ft[kCondStreams][kPitchFeatures];
  const int b0 = blockIdx.x * kCondStreams, c = threadIdx.x;
  const float bias = __ldg(be + c);                  // constants: before the dependency wait
  float wf[kPitchFeatures];
#pragma unroll
  for (int i = 0; i < kPitchFeatures; ++i) wf[i] = __ldg(Wf + i * kHidden + c);
  const int frame = *frame_ptr;                      // written only by the chain's advance kernel
  PdlWait();
  PdlLaunchDependents();
#pragma unroll
  for (int s = 0; s < kCondStreams; ++s) {
    const int b = b0 + s;
    if (c < P) ph[s][c] = b < B ? phone[static_cast<long long>(b) * P + c] : 0.f;
    if (c < kPitchFeatures) ft[s][c] = b < B ? feat[b * kPitchFeatures + c] : 0.f;
  }
  __syncthreads();
  float acc[kCondStreams];
#pragma unroll
  for (int s = 0; s < kCondStreams; ++s) acc[s] = bias;
#pragma unroll 1
  for (int i0 = 0; i0 < P; i0 += 32) {   // P is 128 or 256 (beatrice.h:17,20,23); 32 weight loads in flight per thread
    float w[32];
#pragma unroll
    for (int u = 0; u < 32; ++u) w[u] = __ldg(We + (i0 + u) * kHidden + c);
#pragma unroll
    for (int u = 0; u < 32; ++u) {
#pragma unroll
      for (int s = 0; s < kCondStreams; ++s) acc[s] = fmaf(ph[s][i0 + u], w[u], acc[s]);
    }
  }
#pragma unroll
  for (int s = 0; s < kCondStreams; ++s) {
    const int b = b0 + s;
    if (b >= B) break;
    const int qq = min(max(q[b], 0), bins - 1);
    float v = acc[s] + __ldg(pitch_emb + static_cast<long long>(qq) * kHidden + c);
    float fp = 0.f;
#pragma unroll
    for (int i = 0; i < kPitchFeatures; ++i) fp = fmaf(ft[s][i], wf[i], fp);
    v += fp;
    if (spk) v += spk[static_cast<long long>(b) * kHidden + c];
    if (formant) v += formant[static_cast<long long>(b) * kHidden + c];
    const long long o = (static_cast<long long>(b) * slots + (frame % slots)) * kHidden + c;
    if (ring) ring[o] = v;
    if (ring_hi) {   // bf16 hi [+ lo] planes for the tensor-core pre conv
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      ring_hi[o] = __bfloat16_as_ushort(h);
      if (ring_lo) ring_lo[o] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
    }
  }
}

// kNN-VQ against a 512 x C codebook: out = mean of the n nearest rows (squared L2, ties to
// the lower index); n = 0 or no codebook copies the input through.
__global__ void __launch_bounds__(kCodebookSize) vq_kernel(const float* __restrict__ phone_in,
                                                           float* __restrict__ phone_out,
                                                           const float* const* __restrict__ codebooks,
                                                           const int* __restrict__ n_neighbors, int C) {
  PdlWait();
  PdlLaunchDependents();
  __shared__ float ph[256];
  __shared__ float dist[kCodebookSize];
  __shared__ float red_v[kCodebookSize / 32];
  __shared__ int red_i[kCodebookSize / 32];
  __shared__ int chosen;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* cb = codebooks ? codebooks[b] : nullptr;
  const int n = cb ? min(max(n_neighbors[b], 0), kCodebookSize) : 0;
  if (tid < C) ph[tid] = phone_in[static_cast<long long>(b) * C + tid];
  __syncthreads();
  if (n == 0) {
    if (tid < C) phone_out[static_cast<long long>(b) * C + tid] = ph[tid];
    return;
  }
  {
    // One thread per codebook row, channels accumulated in order (the oracle's arithmetic: the neighbour CHOICE must be the
    // oracle's).  The rows come through shared memory in tiles of 16 channels so that global loads are coalesced -- a thread
    // walking its own 512-byte row took ~50 us per hop (the content lane then outlasted the vocoder at pipeline depth 2).
    __shared__ float tile[kCodebookSize][17];
    float nn = 0.f, dot = 0.f;
    for (int c0 = 0; c0 < C; c0 += 16) {
      __syncthreads();
#pragma unroll 4
      for (int i = tid; i < kCodebookSize * 16; i += kCodebookSize) {
        const int r = i >> 4, c = i & 15;
        tile[r][c] = __ldg(cb + static_cast<long long>(r) * C + c0 + c);
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float ev = tile[tid][c];
        nn = fmaf(ev, ev, nn);
        dot = fmaf(ev, ph[c0 + c], dot);
      }
    }
    dist[tid] = nn - 2.0f * dot;
  }
  float acc = 0.f;
  __syncthreads();
  for (int r = 0; r < n; ++r) {
    float v = dist[tid];
    int vi = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
      if (ov < v || (ov == v && oi < vi)) {
        v = ov;
        vi = oi;
      }
    }
    if ((tid & 31) == 0) {
      red_v[tid >> 5] = v;
      red_i[tid >> 5] = vi;
    }
    __syncthreads();
    if (tid < 32) {
      v = tid < kCodebookSize / 32 ? red_v[tid] : FLT_MAX;
      vi = tid < kCodebookSize / 32 ? red_i[tid] : INT_MAX;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
        if (ov < v || (ov == v && oi < vi)) {
          v = ov;
          vi = oi;
        }
      }
      if (tid == 0) {
        chosen = vi;
        dist[vi] = INFINITY;
      }
    }
    __syncthreads();
    if (tid < C) acc += __ldg(cb + static_cast<long long>(chosen) * C + tid);
    __syncthreads();
  }
  if (tid < C) phone_out[static_cast<long long>(b) * C + tid] = acc * (1.0f / static_cast<float>(n));
}

__device__ __forceinline__ void Project256Body(const float* __restrict__ W, const float* __restrict__ bias,
                                               const float* __restrict__ e, size_t e_stride, size_t src, float* __restrict__ out,
                                               int row) {
  __shared__ float ev[kHidden];
  const int o = threadIdx.x;
  ev[o] = e[src * e_stride + o];
  __syncthreads();
  float acc = __ldg(bias + o);
  for (int i = 0; i < kHidden; ++i) acc = fmaf(ev[i], __ldg(W + i * kHidden + o), acc);
  out[static_cast<long long>(row) * kHidden + o] = acc;
}
__global__ void __launch_bounds__(256) project256_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                         const float* __restrict__ e, size_t e_stride,
                                                         const int* __restrict__ e_index, float* __restrict__ out,
                                                         const int* __restrict__ out_index) {
  const int item = blockIdx.x;
  const size_t src = e_index ? static_cast<size_t>(e_index[item]) : static_cast<size_t>(item);
  Project256Body(W, bias, e, e_stride, src, out, out_index ? out_index[item] : item);
}
__global__ void __launch_bounds__(256) project256_items_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                               const float* __restrict__ e, size_t e_stride,
                                                               float* __restrict__ out, const __grid_constant__ SetterItems items) {
  const int item = blockIdx.x;
  Project256Body(W, bias, e, e_stride, static_cast<size_t>(items.src[item]), out, items.dst[item]);
}

// One stream per block: attention pool of the stream's key-value embedding [384][128] with a learned query, then the linear
// layer to the stage's FiLM (gamma | beta).  Runs at parameter-change rate, but a speaker change costs four of these (one per
// vocoder stage, one per hop) in front of the hop graph, so it is laid out for latency: scores with a warp per row (coalesced
// 512-byte rows, four rows in flight per warp), the pooled vector in three row thirds summed in a fixed order.
__device__ __forceinline__ void KvFilmBody(const float* __restrict__ kv, const float* __restrict__ query,
                                           const float* __restrict__ W, const float* __restrict__ bias, int C,
                                           float* __restrict__ film_row) {
  static_assert(kKvLength == 384 && kKvChannels == 128, "kv_film_kernel is laid out for a 384 x 128 embedding");
  __shared__ float qv[kKvChannels];
  __shared__ float p[kKvLength];
  __shared__ float part[3][kKvChannels];
  __shared__ float pooled[kKvChannels];
  __shared__ float red[kKvLength / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < kKvChannels) qv[tid] = __ldg(query + tid);
  __syncthreads();
  {
    // scores: warp w takes rows w, w + 12, ...; a lane holds 4 consecutive channels of the row
    const float4 q4 = *reinterpret_cast<const float4*>(qv + 4 * lane);
    constexpr int kWarps = kKvLength / 32;
#pragma unroll 1
    for (int r0 = warp; r0 < kKvLength; r0 += 4 * kWarps) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * kWarps;
        v[u] = r < kKvLength ? __ldg(reinterpret_cast<const float4*>(kv + static_cast<size_t>(r) * kKvChannels) + lane)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * kWarps;
        float d = fmaf(v[u].x, q4.x, fmaf(v[u].y, q4.y, fmaf(v[u].z, q4.z, v[u].w * q4.w)));
        d = WarpSum(d);
        if (lane == 0 && r < kKvLength) p[r] = d * (1.0f / sqrtf(static_cast<float>(kKvChannels)));
      }
    }
  }
  __syncthreads();
  const float s = p[tid];
  // block max
  float m = s;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < kKvLength / 32; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  const float ex = expf(s - m);
  float den = WarpSum(ex);
  if (lane == 0) red[warp] = den;
  __syncthreads();
  den = 0.f;
  for (int i = 0; i < kKvLength / 32; ++i) den += red[i];
  p[tid] = ex / den;
  __syncthreads();
  {
    // pooled[ch] = sum_i p[i] kv[i][ch]: thread (third g, channel ch) sums 128 rows, the thirds are added in order
    const int g = tid / kKvChannels, ch = tid - g * kKvChannels;
    const float* col = kv + static_cast<size_t>(g) * 128 * kKvChannels + ch;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int i = 0; i < 128; i += 4) {
      a0 = fmaf(p[g * 128 + i], __ldg(col + static_cast<size_t>(i) * kKvChannels), a0);
      a1 = fmaf(p[g * 128 + i + 1], __ldg(col + static_cast<size_t>(i + 1) * kKvChannels), a1);
      a2 = fmaf(p[g * 128 + i + 2], __ldg(col + static_cast<size_t>(i + 2) * kKvChannels), a2);
      a3 = fmaf(p[g * 128 + i + 3], __ldg(col + static_cast<size_t>(i + 3) * kKvChannels), a3);
    }
    part[g][ch] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  if (tid < kKvChannels) pooled[tid] = (part[0][tid] + part[1][tid]) + part[2][tid];
  __syncthreads();
  if (tid < 2 * C) {
    float a0 = __ldg(bias + tid), a1 = 0.f;
#pragma unroll 4
    for (int i = 0; i < kKvChannels; i += 2) {
      a0 = fmaf(pooled[i], __ldg(W + static_cast<size_t>(i) * 2 * C + tid), a0);
      a1 = fmaf(pooled[i + 1], __ldg(W + static_cast<size_t>(i + 1) * 2 * C + tid), a1);
    }
    film_row[tid] = a0 + a1;
  }
}
__global__ void __launch_bounds__(kKvLength) kv_film_kernel(const float* __restrict__ kv_base,
                                                            const int* __restrict__ kv_index, size_t kv_stride,
                                                            const float* __restrict__ query,
                                                            const float* __restrict__ W, const float* __restrict__ bias,
                                                            int C, float* __restrict__ film_base,
                                                            const int* __restrict__ out_index) {
  const int item = blockIdx.x;
  const float* kv = kv_base + (kv_index ? static_cast<size_t>(kv_index[item]) : 0) * kv_stride;
  const int row = out_index ? out_index[item] : item;
  KvFilmBody(kv, query, W, bias, C, film_base + static_cast<long long>(row) * 2 * C);
}
// the same for a by-value list of (key-value slot, stream, block): one launch serves every block of a hop's schedule
__global__ void __launch_bounds__(kKvLength) kv_film_items_kernel(const float* __restrict__ kv_base, size_t kv_stride,
                                                                  const __grid_constant__ KvBlockTable tab,
                                                                  const __grid_constant__ SetterItems items) {
  const int item = blockIdx.x, blk = items.blk[item];
  const int C = tab.C[blk];
  KvFilmBody(kv_base + static_cast<size_t>(items.src[item]) * kv_stride, tab.query[blk], tab.W[blk], tab.bias[blk], C,
             tab.film[blk] + static_cast<long long>(items.dst[item]) * 2 * C);
}
__global__ void vq_patch_items_kernel(int* __restrict__ vq_n, const float** __restrict__ codebook_ptrs,
                                      const float* __restrict__ codebooks, size_t cb_stride,
                                      const __grid_constant__ SetterItems items) {
  const int i = threadIdx.x;
  if (i < items.n) {
    vq_n[items.dst[i]] = items.blk[i];
    codebook_ptrs[items.dst[i]] = items.src[i] >= 0 ? codebooks + cb_stride * static_cast<size_t>(items.src[i]) : nullptr;
  }
}

__global__ void fill_kernel(float* p, float v, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
void LaunchConvGemm(const ConvDesc* d_descs, const ConvDesc& h0, int nz, int B, const int* d_frame,
                    cudaStream_t s) {
  const int M = B * h0.T, N = h0.N;
  if (N >= 64 || (N > 32 && N % 32 != 0)) {
    dim3 grid((M + 63) / 64, (N + 63) / 64, nz);
    conv_gemm_kernel<64, 64><<<grid, 128, 0, s>>>(d_descs, B, d_frame);
  } else if (N > 16) {
    dim3 grid((M + 127) / 128, (N + 31) / 32, nz);
    conv_gemm_kernel<128, 32><<<grid, 128, 0, s>>>(d_descs, B, d_frame);
  } else {
    dim3 grid((M + 255) / 256, (N + 15) / 16, nz);
    conv_gemm_kernel<256, 16><<<grid, 128, 0, s>>>(d_descs, B, d_frame);
  }
  B200_CHECK(cudaGetLastError());
}

void LaunchDirectConv(const ConvDesc* d_desc, const ConvDesc& h0, int B, const int* d_frame, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * h0.T * h0.N;
  LaunchPdl(direct_conv_kernel, dim3(static_cast<unsigned>((total + 127) / 128)), dim3(128), 0, s, 1, d_desc, B, d_frame);
  B200_CHECK(cudaGetLastError());
}

bool PostConvFused(const ConvDesc& h0) { return h0.k == 7 && h0.C_in == 16 && h0.N == 1 && h0.stride == 1 && h0.dil == 1; }
void LaunchPostConv(const ConvDesc* d_desc, const ConvDesc& h0, int B, const int* d_frame, const AdvanceFold& fold,
                    cudaStream_t s) {
  if (!PostConvFused(h0)) {
    LaunchDirectConv(d_desc, h0, B, d_frame, s);
    return;
  }
  const size_t smem = (7 * 16 + static_cast<size_t>(h0.T + 6) * kPostRowLd) * sizeof(float);
  LaunchPdl(post_conv_kernel, dim3(B), dim3(256), smem, s, 1, h0, B, d_frame, fold);
  B200_CHECK(cudaGetLastError());
}

bool Frontend0Supported(const ConvDesc& h0) {
  return h0.C_in == 1 && h0.N <= 32 && h0.k <= 16 && h0.dil == 1 && h0.x_T <= kInHop && h0.k - h0.stride <= 32 &&
         h0.n_x == 1 && h0.res == nullptr;
}
void LaunchFrontend0(const ConvDesc& h0, const float* staging, float* ring, int B, const int* d_frame, cudaStream_t s) {
  LaunchPdl(frontend0_kernel, dim3(B), dim3(256), 0, s, 1, h0, staging, ring, B, d_frame);
  B200_CHECK(cudaGetLastError());
}

void LaunchNorm(const NormDesc& d, int B, const int* d_frame, cudaStream_t s) {
  const int warps = B * d.T;
  LaunchPdl(channorm_gelu_kernel, dim3((warps * 32 + 127) / 128), dim3(128), 0, s, 1, d, B, d_frame);
  B200_CHECK(cudaGetLastError());
}

void LaunchIngest(const float* staging, float* ring, int slots, int T, int C, int B, const int* d_frame,
                  cudaStream_t s) {
  const long long total = static_cast<long long>(B) * T * C;
  LaunchPdl(ingest_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, 1, staging, ring, slots, T * C, total, d_frame);
  B200_CHECK(cudaGetLastError());
}

void LaunchAdvance(int* d_frame, cudaStream_t s) {
  LaunchPdl(advance_kernel, dim3(1), dim3(1), 0, s, 1, d_frame);
  B200_CHECK(cudaGetLastError());
}

void LaunchPitchArgmax(const float* head, int bins, const int* min_q, const int* max_q, int* q, float* feat, int B,
                       cudaStream_t s, const PitchParams* params, int* q_used) {
  LaunchPdl(pitch_argmax_kernel, dim3((B * 32 + 127) / 128), dim3(128), 0, s, 1, head, bins, min_q, max_q, q, feat, B, params,
            q_used);
  B200_CHECK(cudaGetLastError());
}

void LaunchPitchTransform(const int* q_in, const PitchParams* params, int bins, int* q_out, int B, cudaStream_t s) {
  LaunchPdl(pitch_transform_kernel, dim3((B + 127) / 128), dim3(128), 0, s, 1, q_in, params, bins, q_out, B);
  B200_CHECK(cudaGetLastError());
}

void LaunchCond(const float* phone, int P, const int* q, int bins, const float* feat, const float* We,
                const float* be, const float* pitch_emb, const float* Wf, const float* spk, const float* formant,
                float* ring, uint16_t* ring_hi, uint16_t* ring_lo, int slots, int B, const int* d_frame, cudaStream_t s) {
  LaunchPdl(cond_kernel, dim3((B + kCondStreams - 1) / kCondStreams), dim3(kHidden), 0, s, 1, phone, P, q, bins, feat, We, be,
            pitch_emb, Wf, spk, formant, ring, ring_hi, ring_lo, slots, B, d_frame);
  B200_CHECK(cudaGetLastError());
}

void LaunchVq(const float* phone_in, float* phone_out, const float* const* codebooks, const int* n_neighbors, int C,
              int B, cudaStream_t s) {
  LaunchPdl(vq_kernel, dim3(B), dim3(kCodebookSize), 0, s, 1, phone_in, phone_out, codebooks, n_neighbors, C);
  B200_CHECK(cudaGetLastError());
}

void LaunchProject256(const float* W, const float* b, const float* e, size_t e_stride, const int* e_index,
                      float* out, const int* out_index, int n_items, cudaStream_t s) {
  if (n_items <= 0) return;
  project256_kernel<<<n_items, kHidden, 0, s>>>(W, b, e, e_stride, e_index, out, out_index);
  B200_CHECK(cudaGetLastError());
}

void LaunchKvFilm(const float* kv_base, const int* kv_index, size_t kv_stride, const float* query, const float* W,
                  const float* b, int C, float* film_base, const int* out_index, int n_items, cudaStream_t s) {
  if (n_items <= 0) return;
  kv_film_kernel<<<n_items, kKvLength, 0, s>>>(kv_base, kv_index, kv_stride, query, W, b, C, film_base, out_index);
  B200_CHECK(cudaGetLastError());
}

void LaunchProject256Items(const float* W, const float* b, const float* e, size_t e_stride, float* out, const SetterItems& items,
                           cudaStream_t s) {
  if (items.n <= 0) return;
  project256_items_kernel<<<items.n, kHidden, 0, s>>>(W, b, e, e_stride, out, items);
  B200_CHECK(cudaGetLastError());
}
void LaunchKvFilmItems(const float* kv_base, size_t kv_stride, const KvBlockTable& tab, const SetterItems& items, cudaStream_t s) {
  if (items.n <= 0) return;
  kv_film_items_kernel<<<items.n, kKvLength, 0, s>>>(kv_base, kv_stride, tab, items);
  B200_CHECK(cudaGetLastError());
}
void LaunchVqPatchItems(int* vq_n, const float** codebook_ptrs, const float* codebooks, size_t cb_stride, const SetterItems& items,
                        cudaStream_t s) {
  if (items.n <= 0) return;
  vq_patch_items_kernel<<<1, kSetterItems, 0, s>>>(vq_n, codebook_ptrs, codebooks, cb_stride, items);
  B200_CHECK(cudaGetLastError());
}

void LaunchFill(float* p, float v, size_t n, cudaStream_t s) {
  if (n == 0) return;
  fill_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(p, v, n);
  B200_CHECK(cudaGetLastError());
}

}  // namespace b200
