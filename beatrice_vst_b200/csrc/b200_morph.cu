// Voice-morphing mode on the device: the spherical weighted average of up to eight speakers'
// embeddings that the reference call site computes on the host, per stream, whenever the morphing weights change
// (reference src/common/processor_core_2.cc:51-177 -- the additive embedding, 256-d, in one frame; the key-value
// embedding, 384 rows of 128, a quarter per frame over four frames -- with src/common/spherical_average.h:80-444).
//
// One WARP per average.  The algorithm is the reference's, step for step (SetWeights :164-216, then at most
// kSphAvgMaxNUpdates = 4 Update() :218-231, GetResult :233-241): the weighted chord mean as the start point, then
// L-BFGS (two correction pairs) on the sphere for the point whose tangent-space weighted mean of the inputs
// vanishes; the result is the same affine combination applied to the UN-normalised inputs.  Vectors live in
// registers, element l + 32 e in lane l; every inner product is a per-lane partial sum followed by a shuffle
// reduction, so sums are taken in a different order than the reference's sequential loops (agreement ~1e-7
// relative; tests hold it to 1e-6 against the reference header compiled in oracle/_ref).
#include <cfloat>

#include "b200_common.h"
#include "b200_kernels.h"

namespace b200 {
namespace {

constexpr int kMaxPts = 8;   // processor_core_2.h:26 kSphAvgMaxNSpeakers
constexpr int kMem = 2;      // spherical_average.h:134 num_memory
constexpr int kUpdates = 4;  // processor_core_2.h:90 kSphAvgMaxNUpdates

__device__ __forceinline__ float WarpSum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int E>
__device__ __forceinline__ float Dot(const float (&a)[E], const float (&b)[E]) {
  float s = 0.0f;
#pragma unroll
  for (int e = 0; e < E; ++e) s = fmaf(a[e], b[e], s);
  return WarpSum(s);
}
// spherical_average.h:307-325.  NB the reference writes an unqualified `abs(x)` there, which -- as the reference
// compiles on this platform (g++ / glibc: oracle/_ref) -- binds to the C library's abs(int): the angle is TRUNCATED to
// an integer before the threshold tests, so the series branch returns exactly 1 for every |x| < 1 and sin(x)/x is
// used from 1 radian on.  Results must be those of the reference as built here, so the truncation is reproduced
// (tests/test_morph.py compares with that build; a build whose <cmath> offers a global abs(float) would differ).
__device__ __forceinline__ float SincRef(float x) {
  const float t0 = FLT_EPSILON, t1 = sqrtf(t0), t2 = sqrtf(t1);
  const float ax = fabsf(truncf(x));
  if (ax >= t2) return sinf(x) / x;
  float y = 1.0f;
  if (ax >= t0) {
    const float x2 = x * x;
    y -= x2 / 6.0f;
    if (ax >= t1) y += x2 * x2 / 120.0f;
  }
  return y;
}

template <int M>
struct SphAvg {
  static constexpr int E = M / 32;
  float p[kMaxPts][E];   // normalised inputs
  float w[kMaxPts], v[kMaxPts];
  float q[E], g[E], d[E], s[kMem][E], t[kMem][E];
  float r[kMem], a[kMem], gamma;
  int n, mem;
  bool converged;

  __device__ void UpdateVGD() {   // :334-374
    float sum = 0.0f;
#pragma unroll
    for (int e = 0; e < E; ++e) g[e] = 0.0f;
#pragma unroll
    for (int i = 0; i < kMaxPts; ++i) {
      if (i < n) {
        float c = Dot<E>(p[i], q);
        c = fminf(fmaxf(c, -1.0f), 1.0f);
        const float theta = acosf(c);
        const float inv = 1.0f / (SincRef(theta) + FLT_EPSILON);
        sum += w[i] * c * inv;
        v[i] = w[i] * inv;
        const float an = -2.0f * v[i];
#pragma unroll
        for (int e = 0; e < E; ++e) g[e] = fmaf(an, p[i][e], g[e]);
      }
    }
    const float inv_sum = 1.0f / (sum + FLT_EPSILON);
#pragma unroll
    for (int i = 0; i < kMaxPts; ++i)
      if (i < n) v[i] *= inv_sum;
    const float mip = -Dot<E>(q, g);   // ProjectVectorToPlane(q, g)
#pragma unroll
    for (int e = 0; e < E; ++e) {
      g[e] = fmaf(mip, q[e], g[e]);
      d[e] = g[e];
    }
#pragma unroll
    for (int k = 0; k < kMem; ++k) {
      const int idx = (mem - k - 1 + kMem) % kMem;
      float dot = 0.0f;
      // s / t are indexed dynamically: select with a predicate so that they stay in registers
#pragma unroll
      for (int m2 = 0; m2 < kMem; ++m2)
        if (m2 == idx) dot = Dot<E>(s[m2], d);
      const float ak = (idx == 0 ? r[0] : r[1]) * dot;
      if (idx == 0) a[0] = ak; else a[1] = ak;
#pragma unroll
      for (int m2 = 0; m2 < kMem; ++m2)
        if (m2 == idx) {
#pragma unroll
          for (int e = 0; e < E; ++e) d[e] = fmaf(-ak, t[m2][e], d[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) d[e] *= gamma;
#pragma unroll
    for (int k = 0; k < kMem; ++k) {
      const int idx = (mem + k) % kMem;
      float dot = 0.0f;
#pragma unroll
      for (int m2 = 0; m2 < kMem; ++m2)
        if (m2 == idx) dot = Dot<E>(t[m2], d);
      const float b = (idx == 0 ? r[0] : r[1]) * dot;
      const float coef = (idx == 0 ? a[0] : a[1]) - b;
#pragma unroll
      for (int m2 = 0; m2 < kMem; ++m2)
        if (m2 == idx) {
#pragma unroll
          for (int e = 0; e < E; ++e) d[e] = fmaf(coef, s[m2][e], d[e]);
        }
    }
  }

  __device__ void Update() {   // :218-231 with UpdateQS :391-405, UpdateVGDT :376-389, UpdateGammaR :407-415
    if (converged) return;
    const float norm_d = sqrtf(Dot<E>(d, d));
    if (!(norm_d >= 8.0f * FLT_EPSILON)) {
      converged = true;
      return;
    }
#pragma unroll
    for (int m2 = 0; m2 < kMem; ++m2)
      if (m2 == mem) {
        // UpdateQS
#pragma unroll
        for (int e = 0; e < E; ++e) {
          s[m2][e] = q[e];
          q[e] -= d[e];
        }
        const float nq = sqrtf(Dot<E>(q, q));
        if (nq > 0.0f) {
          const float sc = 1.0f / nq;
#pragma unroll
          for (int e = 0; e < E; ++e) q[e] *= sc;
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
          s[m2][e] = q[e] - s[m2][e];
          t[m2][e] = g[e];   // UpdateVGDT, first line
        }
      }
    UpdateVGD();
#pragma unroll
    for (int m2 = 0; m2 < kMem; ++m2)
      if (m2 == mem) {
#pragma unroll
        for (int e = 0; e < E; ++e) t[m2][e] = g[e] - t[m2][e];
        const float mip = -Dot<E>(q, t[m2]);
#pragma unroll
        for (int e = 0; e < E; ++e) t[m2][e] = fmaf(mip, q[e], t[m2][e]);
        // UpdateGammaR
        gamma = Dot<E>(s[m2], t[m2]);
        const float rr = 1.0f / gamma;
        if (m2 == 0) r[0] = rr; else r[1] = rr;
        gamma /= Dot<E>(t[m2], t[m2]);
      }
    mem = (mem + 1 >= kMem) ? 0 : mem + 1;
  }
};

// jobs[j]: one stream's weights; item i of job j averages row (job.item0 + i) of the speakers job.idx[0..n).
// table: [speaker][rows_per_speaker][M]; dst: [dst_row][rows_per_speaker][M].
template <int M>
__global__ void __launch_bounds__(128) sph_avg_kernel(const float* __restrict__ table, long long speaker_stride,
                                                      const MorphJob* __restrict__ jobs, int items_per_job, int n_items,
                                                      float* __restrict__ dst, long long dst_stride) {
  constexpr int E = M / 32;
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= n_items) return;
  const MorphJob job = jobs[item / items_per_job];
  const int row = job.item0 + item % items_per_job;
  SphAvg<M> A;
  A.n = job.n;
  // SetWeights (:164-216): weights in arg-sorted order, already cut at the first zero by the host
  float wsum = 0.0f;
#pragma unroll
  for (int i = 0; i < kMaxPts; ++i) {
    A.w[i] = i < A.n ? job.w[i] : 0.0f;
    A.v[i] = 0.0f;
    if (i < A.n) wsum += A.w[i];
  }
  A.converged = true;
  if (A.n > 0 && wsum > 0.0f) {
    const float sc = 1.0f / wsum;
#pragma unroll
    for (int i = 0; i < kMaxPts; ++i) A.w[i] *= sc;
#pragma unroll
    for (int e = 0; e < E; ++e) A.q[e] = 0.0f;
#pragma unroll
    for (int i = 0; i < kMaxPts; ++i) {
      if (i < A.n) {
        const float* src = table + job.idx[i] * speaker_stride + static_cast<long long>(row) * M;
#pragma unroll
        for (int e = 0; e < E; ++e) A.p[i][e] = src[lane + 32 * e];
        const float nrm = sqrtf(Dot<E>(A.p[i], A.p[i]));     // Initialize(): NormalizeVector (:157-159, :286-296)
        if (nrm > 0.0f) {
          const float s1 = 1.0f / nrm;
#pragma unroll
          for (int e = 0; e < E; ++e) A.p[i][e] *= s1;
        }
#pragma unroll
        for (int e = 0; e < E; ++e) A.q[e] = (i == 0) ? A.w[0] * A.p[0][e] : fmaf(A.w[i], A.p[i][e], A.q[e]);
      }
    }
    const float nq = sqrtf(Dot<E>(A.q, A.q));
    if (nq > 0.0f) {
      const float s1 = 1.0f / nq;
#pragma unroll
      for (int e = 0; e < E; ++e) A.q[e] *= s1;
      A.converged = false;
    }
  }
  if (!A.converged) {
    A.mem = 0;
    A.gamma = 1.0f;
#pragma unroll
    for (int m2 = 0; m2 < kMem; ++m2) {
      A.r[m2] = A.a[m2] = 0.0f;
#pragma unroll
      for (int e = 0; e < E; ++e) A.s[m2][e] = A.t[m2][e] = 0.0f;
    }
    A.UpdateVGD();
    for (int j = 0; j < kUpdates; ++j) {
      A.Update();
      if (A.converged) break;
    }
  }
  // GetResult (:233-241): the same combination of the RAW inputs; no inputs / zero weights -> zeros
  float y[E];
#pragma unroll
  for (int e = 0; e < E; ++e) y[e] = 0.0f;
#pragma unroll
  for (int i = 0; i < kMaxPts; ++i) {
    if (i < A.n) {
      const float* src = table + job.idx[i] * speaker_stride + static_cast<long long>(row) * M;
#pragma unroll
      for (int e = 0; e < E; ++e) y[e] = (i == 0) ? A.v[0] * src[lane + 32 * e] : fmaf(A.v[i], src[lane + 32 * e], y[e]);
    }
  }
  float* out = dst + job.dst_row * dst_stride + static_cast<long long>(row) * M;
#pragma unroll
  for (int e = 0; e < E; ++e) out[lane + 32 * e] = y[e];
}

}  // namespace

void LaunchSphAvg(int M, const float* table, long long speaker_stride, const MorphJob* d_jobs, int n_jobs,
                  int items_per_job, float* dst, long long dst_stride, cudaStream_t s) {
  if (n_jobs <= 0 || items_per_job <= 0) return;
  const int n_items = n_jobs * items_per_job;
  const int blocks = (n_items + 3) / 4;
  if (M == 256) sph_avg_kernel<256><<<blocks, 128, 0, s>>>(table, speaker_stride, d_jobs, items_per_job, n_items, dst, dst_stride);
  else if (M == 128) sph_avg_kernel<128><<<blocks, 128, 0, s>>>(table, speaker_stride, d_jobs, items_per_job, n_items, dst, dst_stride);
  else Fail(-107, "spherical average: unsupported width", __FILE__, __LINE__);
  B200_CHECK(cudaGetLastError());
}

}  // namespace b200
