// tcgen05 (5th-generation tensor core) form of the causal-conv implicit GEMM, sm_100a only.
//
//   D[128 rows x BN cols] (fp32, in TMEM)  +=  A[128 x 64] (bf16, smem)  *  W[64 x BN] (bf16, smem)
//
// One CTA = one 128-row tile of the (stream, time) axis x one BN-wide tile of output channels.
// K is walked in chunks of 64 (one tap x 64 input channels, or several taps when C_in < 64).
// Per chunk and pipeline stage:
//   * the weight tile arrives by ONE TMA bulk copy (cp.async.bulk, mbarrier complete_tx) from a
//     host-prepacked image that already has the UMMA canonical K-major / no-swizzle layout
//       [K/8 panels][rows][8 elements]   (core matrix = 8 rows x 16 B contiguous)
//   * the 128 threads gather their own activation row for that tap from the fp32 ring buffers
//     (sum of up to three branches, scale, LeakyReLU), convert to bf16 and write one 16-byte
//     vector per panel -- conflict free because consecutive threads own consecutive rows;
//   * one elected thread issues 4 tcgen05.mma (K = 16 each) and tcgen05.commit's the stage's
//     "empty" mbarrier, so the gather of the next chunk overlaps the MMAs of this one.
// Epilogue: tcgen05.ld (32 lanes x 32-bit x 16 columns per warp) -> bias / FiLM / residual /
// activation -> fp32 ring store, identical to the CUDA-core kernel's epilogue.
//
// kSplit adds the hi/lo bf16 decomposition of both operands (x = hi + lo): three MMAs
// hi*hi + hi*lo + lo*hi per K step recover ~16 mantissa bits with fp32 accumulation.
#include <cuda_bf16.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "b200_common.h"
#include "b200_kernels.h"
#include "b200_tc_common.cuh"

namespace b200 {
namespace {

constexpr int kTcM = 128;   // rows per CTA == TMEM lanes
constexpr int kTcKC = 64;   // K elements per chunk
// 8 worker warps (producers, then epilogue) + 1 MMA-issuing warp.  Two workers per SM sub-partition:
// the producer and epilogue code is latency bound per warp (one warp per scheduler cannot hide its
// own dependent-issue stalls), so splitting it over twice the warps nearly halves it.
constexpr int kTcWorkers = 256;
constexpr int kTcThreads = kTcWorkers + 32;
// bytes of one 8-element K panel of the activation tile: 128 rows x 16 B, +16 so that the eight
// panels of one row fall into different shared-memory bank groups (the cp.async producer writes
// the eight 16-byte pieces of a row from eight adjacent lanes)
constexpr int kPanelA = kTcM * 16 + 16;

// kGather = false compiles the cp.async-only producer (no fp32 gather code): ~half the registers,
// so three to four CTAs fit on an SM for the small-tile layers.
template <bool kSplit, int kStages, bool kGather>
__global__ void __launch_bounds__(kTcThreads, 2) conv_gemm_tc_kernel(const __grid_constant__ ConvDesc d0,
                                                           const ConvDesc* __restrict__ descs, int B,
                                                           const int* __restrict__ frame_ptr, int KS) {
  constexpr int kOperands = kSplit ? 2 : 1;
  extern __shared__ __align__(1024) uint8_t smem[];

  // descriptor of z = 0 rides in the kernel parameters (no global round trip before the pipeline
  // can start); the rare z-batched launch reads the others from the device array (constant data)
  ConvDesc d = d0;
  if (blockIdx.z != 0) d = descs[blockIdx.z];
  const int BN = d.tc_bn;
  const int tid = threadIdx.x, lane = tid & 31;
  // warp index broadcast from lane 0: provably warp-uniform, so the role branches below are uniform
  // control flow and the single-thread MMA / TMA loops can live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int M = B * d.T;
  // K split: a cluster of KS CTAs (consecutive blockIdx.x) shares one output tile; CTA `rank` walks the
  // chunk range [rank, rank + 1) * n_chunks / KS, then the partial accumulators are reduce-scattered
  // through distributed shared memory and each CTA finishes BN / KS columns of the tile.
  const int rank = KS > 1 ? static_cast<int>(ClusterCtaRank()) : 0;
  const int m0 = (blockIdx.x / KS) * kTcM, n0 = blockIdx.y * BN;
  const int C_in = d.C_in, N = d.N;

  const uint32_t a_bytes = 8 * kPanelA;               // one bf16 plane of the activation tile
  const uint32_t w_bytes = static_cast<uint32_t>(BN) * kTcKC * 2;
  const uint32_t stage_bytes = (a_bytes + w_bytes) * kOperands;
  // pipeline area; never smaller than the epilogue's staging tile [128][BN / KS + 4] fp32 (see the launcher)
  const uint32_t tile_bytes = (static_cast<uint32_t>(kTcM) * (BN / KS + 4) * 4 + 127) / 128 * 128;
  const uint32_t pipe_bytes = kStages * stage_bytes > tile_bytes ? kStages * stage_bytes : tile_bytes;
  uint8_t* tail = smem + pipe_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);  // full[kStages], empty[kStages], done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2);
  float* bias_s = reinterpret_cast<float*>(tail + 128);   // BN floats
  const uint32_t bar_full = SmemAddr(bars), bar_empty = SmemAddr(bars + kStages), bar_done = SmemAddr(bars + 2 * kStages),
                 bar_box = SmemAddr(bars + 2 * kStages + 1);   // peers' partial sums have landed in the inbox
  const uint32_t box_off = static_cast<uint32_t>(tail - smem) + 128 + 1024 + 1024;   // [KS-1][128 rows][BN/KS + 4] fp32
  // developer timeline (d.trace): [0..7] phases, [8+4c..] per chunk: producer issue / landed, MMA full / commit
  long long* trace = reinterpret_cast<long long*>(tail + 128 + 1024);
  const bool tracing = d0.trace != 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define B200_TR(slot) do { if (tracing) trace[slot] = clock64(); } while (0)
  if (tracing && tid == 0) trace[0] = clock64();
  const uint32_t smem_base = SmemAddr(smem);

  // TMEM columns: power of two >= 32 covering BN fp32 accumulator columns
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(BN)) tmem_cols <<= 1;

  for (int i = tid; i < BN; i += kTcThreads) {
    const int col = blockIdx.y * BN + i;
    bias_s[i] = (d.bias && col < d.N) ? __ldg(d.bias + col) : 0.0f;
  }
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      MbarInit(bar_full + 8 * i, kTcWorkers + 1);   // every worker thread + the weight TMA's expect_tx arrive
      MbarInit(bar_empty + 8 * i, 1);
    }
    MbarInit(bar_done, 1);
    MbarInit(bar_box, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemAddr(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  TcFenceBefore();
  __syncthreads();
  if (KS > 1) ClusterSyncAll();   // every CTA's barriers exist before a peer may signal them
  TcFenceAfter();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // same value in every lane, provably
  if (tid == 0) B200_TR(1);

  // ---- activation rows (index arithmetic only: done before the dependency wait) ----
  // gather mode: thread == row (fp32 loads, conversion in registers).
  // cp.async mode: lane group of 8 == one row, lane & 7 == K panel, so one warp instruction moves
  // four whole 128-byte row segments (coalesced); every worker serves rows rg + 32 i, i < 4.
  const int x_L = d.x_slots * d.x_T;
  const int m = m0 + (tid & 127);
  const bool row_ok = m < M;
  const int whalf = (tid >> 7) & 1;   // which half of the split work (K half / column half) a worker takes
  long long xbase = 0;
  int xu0 = 0;
  if (row_ok) {
    const int b = m / d.T, t = m - b * d.T;
    xbase = static_cast<long long>(b) * x_L * C_in;
    xu0 = t * d.stride + d.stride - 1;
  }
  const int g_l16 = tid & 15, g_rg = (tid >> 4) & 15;   // gather mode: lane group of 16 == one row
  long long g_base[8];
  int g_u0[8];
  if constexpr (kGather) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int mi = m0 + g_rg + 16 * i;
      g_base[i] = -1;
      g_u0[i] = 0;
      if (mi < M) {
        const int b = mi / d.T, t = mi - b * d.T;
        g_base[i] = static_cast<long long>(b) * x_L * C_in;
        g_u0[i] = t * d.stride + d.stride - 1;
      }
    }
  }
  const int a_pl = tid & 7, a_rg = (tid >> 3) & 31;
  long long a_base[4];
  int a_u0[4];
  if constexpr (!kGather) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int mi = m0 + a_rg + 32 * i;
      a_base[i] = -1;
      a_u0[i] = 0;
      if (mi < M) {
        const int b = mi / d.T, t = mi - b * d.T;
        a_base[i] = static_cast<long long>(b) * x_L * C_in;
        a_u0[i] = t * d.stride + d.stride - 1;
      }
    }
  }

  // ---- weights of the first kStages chunks: constants, fetched while the predecessor drains ----
  const int n_sub = d.C_in >= kTcKC ? d.C_in / kTcKC : 1;
  const int tpc = d.C_in >= kTcKC ? 1 : kTcKC / d.C_in;       // taps per chunk
  const int n_chunks = d.C_in >= kTcKC ? d.k * n_sub : (d.k + tpc - 1) / tpc;
  const int c_begin = rank * n_chunks / KS, c_end = (rank + 1) * n_chunks / KS;   // this CTA's K range
  const uint8_t* w_hi = static_cast<const uint8_t*>(d.w_tc) + static_cast<size_t>(blockIdx.y) * n_chunks * w_bytes;
  const uint8_t* w_lo = kSplit ? static_cast<const uint8_t*>(d.w_tc_lo) + static_cast<size_t>(blockIdx.y) * n_chunks * w_bytes
                               : nullptr;
  if (tid == 0) {
    for (int c = c_begin; c < c_begin + kStages && c < c_end; ++c) {
      const int st = c - c_begin;
      const uint32_t w_s = smem_base + st * stage_bytes + a_bytes * kOperands;
      MbarExpectTx(bar_full + 8 * st, w_bytes * kOperands);
      TmaBulkLoadKeep(w_s, w_hi + static_cast<size_t>(c) * w_bytes, w_bytes, bar_full + 8 * st);
      if (kSplit) TmaBulkLoadKeep(w_s + w_bytes, w_lo + static_cast<size_t>(c) * w_bytes, w_bytes, bar_full + 8 * st);
    }
  }
  // The hop counter is written only by the advance kernel that ends a chain, and no conv kernel directly
  // follows one (a PDL kernel overlaps at most its immediate predecessor): safe to read before the wait.
  const int frame = *frame_ptr;
  PdlWait();   // from here on: data written by the predecessor (activations)
  PdlLaunchDependents();   // the successor's prologue may overlap this kernel's main loop
  if (tid == 0) B200_TR(2);
  if (tracing && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(*reinterpret_cast<unsigned long long*>(&trace[120])));

  const int x_cur = (frame % d.x_slots) * d.x_T;

  // ---- chunk enumeration (must match PackWeightsTc) ----
  const int cw = C_in >= kTcKC ? kTcKC : C_in;            // channels gathered per tap
  const int lcw = 31 - __clz(cw);                         // cw is 16, 32 or 64
  const uint32_t idesc = MakeIdesc(BN);
  const bool async_a = !kGather;          // activations pre-rounded to bf16 by their producer (d.xh)

  // element offsets of the (up to four) taps of chunk c for this thread's row
  auto tap_offsets = [&](int c, long long* ro) {
    const int j0 = C_in >= kTcKC ? c / n_sub : c * tpc;
    const int ci0 = C_in >= kTcKC ? (c - j0 * n_sub) * kTcKC : 0;
#pragma unroll
    for (int tl = 0; tl < 4; ++tl) {
      const int j = min(j0 + tl, d.k - 1);
      int r = x_cur + xu0 - (d.k - 1 - j) * d.dil;
      if (r < 0) r += x_L;
      ro[tl] = xbase + static_cast<long long>(r) * C_in + ci0;
    }
  };
  // stage c % kStages: wait until its previous MMAs retired, then start the weight TMA and
  // (async mode) the 16-byte cp.async's that drop this thread's row straight into the panels
  auto issue = [&](int c) {
    const int s = (c - c_begin) % kStages, round = (c - c_begin) / kStages;
    if (round > 0) MbarWait(bar_empty + 8 * s, (round - 1) & 1);
    const uint32_t st_base = smem_base + s * stage_bytes;
    const uint32_t w_hi_s = st_base + a_bytes * kOperands;
    if (tid == 0 && round > 0) {   // round 0 was issued before PdlWait()
      MbarExpectTx(bar_full + 8 * s, w_bytes * kOperands);
      TmaBulkLoadKeep(w_hi_s, w_hi + static_cast<size_t>(c) * w_bytes, w_bytes, bar_full + 8 * s);
      if (kSplit) TmaBulkLoadKeep(w_hi_s + w_bytes, w_lo + static_cast<size_t>(c) * w_bytes, w_bytes, bar_full + 8 * s);
    }
    if constexpr (!kGather) {
      const int j0 = C_in >= kTcKC ? c / n_sub : c * tpc;
      const int ci0 = C_in >= kTcKC ? (c - j0 * n_sub) * kTcKC : 0;
      const int ch = a_pl * 8;
      const int tl = ch >> lcw, cc = ch & (cw - 1);
      const int back = (d.k - 1 - min(j0 + tl, d.k - 1)) * d.dil;
      const uint32_t dst = st_base + a_pl * kPanelA + a_rg * 16;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int r = x_cur + a_u0[i] - back;
        if (r < 0) r += x_L;
        const bool ok = a_base[i] >= 0;
        const long long a = (ok ? a_base[i] : 0) + static_cast<long long>(r) * C_in + ci0 + cc;
        CpAsync16(dst + i * 512, d.xh + a, ok);
        if (kSplit) CpAsync16(dst + a_bytes + i * 512, d.xl + a, ok);
      }
    }
  };

  // ---- warp-specialised main loop ----
  // warps 0-7 are PRODUCERS: they fill pipeline stages and
  // arrive on the stage's "full" mbarrier when their own bytes have landed -- no block barrier.
  // warp 8 is the MMA ISSUER: one elected lane waits on "full", issues the tcgen05.mma's of the
  // chunk and commits the stage's "empty" mbarrier.  The roles only meet through mbarriers, so
  // the tensor pipe is fed back-to-back while the producers run up to kStages chunks ahead.
  constexpr int kRetire = kStages >= 3 ? kStages - 2 : 0;   // cp.async groups left in flight
  if (warp < 8) {
    for (int c = c_begin; c < c_end; ++c) {
      const int lc = c - c_begin;
      const int s = lc % kStages;
      issue(c);
      if (tid == 0 && lc < 24) B200_TR(8 + 4 * lc);
      if constexpr (!kGather) {
        CpAsyncCommit();
        if (lc >= kRetire) {
          CpAsyncWait<kRetire>();   // this thread's part of chunk c-kRetire has landed
          FenceProxyAsync();        // ... and is visible to the tensor core (async proxy)
          MbarArrive(bar_full + 8 * ((lc - kRetire) % kStages));
          if (tid == 0 && lc - kRetire < 24) B200_TR(8 + 4 * (lc - kRetire) + 1);
        }
      } else {
        // register path (inputs that exist only as fp32 rings, e.g. the mean of the three MRF branches):
        // 16 lanes cover the 64 K-elements (256 B) of one row, so a warp instruction reads two whole row
        // segments (coalesced); every worker serves rows g_rg + 16 i, i < 8, in two batches of four so
        // that all loads of a batch are in flight before the first use.
        const uint32_t a_dst = smem_base + s * stage_bytes + (g_l16 >> 1) * kPanelA + (g_l16 & 1) * 8;
        const int j0 = C_in >= kTcKC ? c / n_sub : c * tpc;
        const int ci0 = C_in >= kTcKC ? (c - j0 * n_sub) * kTcKC : 0;
        const int ch = g_l16 * 4;
        const int tl = ch >> lcw, cc = ch & (cw - 1);
        const int back = (d.k - 1 - min(j0 + tl, d.k - 1)) * d.dil;
#pragma unroll
        for (int bt = 0; bt < 2; ++bt) {
          float4 x0[4], x1[4], x2[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ii = bt * 4 + i;
            int r = x_cur + g_u0[ii] - back;
            if (r < 0) r += x_L;
            const bool ok = g_base[ii] >= 0;
            const long long a = (ok ? g_base[ii] : 0) + static_cast<long long>(r) * C_in + ci0 + cc;
            x0[i] = ok ? __ldg(reinterpret_cast<const float4*>(d.x[0] + a)) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (d.n_x > 1) {
              x1[i] = ok ? __ldg(reinterpret_cast<const float4*>(d.x[1] + a)) : make_float4(0.f, 0.f, 0.f, 0.f);
              x2[i] = ok ? __ldg(reinterpret_cast<const float4*>(d.x[2] + a)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float v[4] = {x0[i].x, x0[i].y, x0[i].z, x0[i].w};
            if (d.n_x > 1) {
              v[0] = ((v[0] + x1[i].x) + x2[i].x) * d.in_scale;
              v[1] = ((v[1] + x1[i].y) + x2[i].y) * d.in_scale;
              v[2] = ((v[2] + x1[i].z) + x2[i].z) * d.in_scale;
              v[3] = ((v[3] + x1[i].w) + x2[i].w) * d.in_scale;
            }
            if (d.in_act == kActLrelu) {   // the only input activation of spec M0
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.0f ? v[e] : 0.1f * v[e];
            }
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v[0]), h1 = __float2bfloat16_rn(v[1]),
                                h2 = __float2bfloat16_rn(v[2]), h3 = __float2bfloat16_rn(v[3]);
            const uint32_t hi0 = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
            const uint32_t hi1 = static_cast<uint32_t>(__bfloat16_as_ushort(h2)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h3)) << 16);
            const uint32_t dst = a_dst + (g_rg + 16 * (bt * 4 + i)) * 16;
            asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(dst), "r"(hi0), "r"(hi1) : "memory");
            if (kSplit) {
              const __nv_bfloat16 l0 = __float2bfloat16_rn(v[0] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v[1] - __bfloat162float(h1)),
                                  l2 = __float2bfloat16_rn(v[2] - __bfloat162float(h2)), l3 = __float2bfloat16_rn(v[3] - __bfloat162float(h3));
              const uint32_t lo0 = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
              const uint32_t lo1 = static_cast<uint32_t>(__bfloat16_as_ushort(l2)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l3)) << 16);
              asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(dst + a_bytes), "r"(lo0), "r"(lo1) : "memory");
            }
          }
        }
        FenceProxyAsync();
        MbarArrive(bar_full + 8 * s);
        if (tid == 0 && lc < 24) B200_TR(8 + 4 * lc + 1);
      }
    }
    if (!kGather && kRetire > 0) {   // drain: the last kRetire chunks
      CpAsyncWait<0>();
      FenceProxyAsync();
      const int nl = c_end - c_begin;
      for (int lc = (nl > kRetire ? nl - kRetire : 0); lc < nl; ++lc) MbarArrive(bar_full + 8 * (lc % kStages));
    }
  } else {   // warp 8 walks the issue loop converged; one elected lane issues each MMA
    for (int c = c_begin; c < c_end; ++c) {
      const int lc = c - c_begin;
      const int s = lc % kStages, round = lc / kStages;
      const int j0 = C_in >= kTcKC ? c / n_sub : c * tpc;
      const int taps = C_in >= kTcKC ? 1 : min(tpc, d.k - j0);
      const uint32_t st_base = smem_base + s * stage_bytes;
      const uint32_t a_hi = st_base, w_hi_s = st_base + a_bytes * kOperands;
      const uint32_t a_lo = st_base + a_bytes, w_lo_s = w_hi_s + w_bytes;
      MbarWait(bar_full + 8 * s, round & 1);   // 128 row arrivals + the weight TMA's bytes
      TcFenceAfter();
      if (lc < 24 && lane == 0) B200_TR(8 + 4 * lc + 2);
      const int ksteps = (taps * cw) >> 4;
      for (int kk = 0; kk < ksteps; ++kk) {
        const uint64_t ah = MakeDesc(a_hi + 2 * kk * kPanelA, kPanelA, 128);
        const uint64_t wh = MakeDesc(w_hi_s + 2 * kk * BN * 16, BN * 16, 128);
        const uint32_t acc = (lc > 0 || kk > 0) ? 1u : 0u;
        MmaW(tmem_base, ah, wh, idesc, acc);
        if (kSplit) {
          const uint64_t al = MakeDesc(a_lo + 2 * kk * kPanelA, kPanelA, 128);
          const uint64_t wl = MakeDesc(w_lo_s + 2 * kk * BN * 16, BN * 16, 128);
          MmaW(tmem_base, ah, wl, idesc, 1u);
          MmaW(tmem_base, al, wh, idesc, 1u);
        }
      }
      MmaCommitW(bar_empty + 8 * s);            // frees the stage when these MMAs retire
      if (lc < 24 && lane == 0) B200_TR(8 + 4 * lc + 3);
      if (c == c_end - 1) MmaCommitW(bar_done);
    }
  }
  __syncwarp();   // re-converge the MMA warp (its other 31 lanes skipped the loop)

  // ---- epilogue: TMEM -> registers -> bias / FiLM / residual / activation -> rings ----
  // Every row of the tile is owned by one thread (TMEM lane == thread).  Bias sits in shared
  // memory since kernel start; the residual tile is pulled into the (now idle) pipeline
  // buffers with one batch of cp.async so its DRAM latency is paid once, not per 16 columns.
  // Worker warps w and w + 4 share TMEM lane quarter w & 3 (== tile rows 32 (w & 3) ..) and split that
  // quarter's columns between them: whalf 0 takes the first half of the 16-column groups, whalf 1 the rest.
  if (warp < 8) {
  const int row = tid & 127;
  const int mm = m0 + row;
  const bool out_ok = mm < M;
  int ob = 0, ot = 0;
  if (out_ok) {
    ob = mm / d.T;
    ot = mm - ob * d.T;
  }
  const int y_L = d.y_slots * d.y_T;
  const int y_cur = (frame % d.y_slots) * d.y_T;
  const int yh_L = d.yh_slots * d.y_T;
  const int yh_cur = d.yh ? (frame % d.yh_slots) * d.y_T : 0;
  const int res_L = d.res_slots * d.res_T;
  const int res_cur = d.res ? (frame % d.res_slots) * d.res_T : 0;
  float* out_row = d.y ? d.y + (static_cast<long long>(ob) * y_L + y_cur) * d.y_C + static_cast<long long>(ot) * N : nullptr;
  const long long oh = (static_cast<long long>(ob) * yh_L + yh_cur) * d.y_C + static_cast<long long>(ot) * N;
  const float* res_row = d.res ? d.res + (static_cast<long long>(ob) * res_L + res_cur + ot) * N : nullptr;
  const float* film_row = d.film ? d.film + static_cast<long long>(ob) * 2 * d.film_C : nullptr;
  // The (now idle) pipeline buffers become an fp32 staging tile [128 rows][CW + 4] of this CTA's own
  // columns: the residual lands there by cp.async, each thread replaces it in place with its finished
  // values, and the rows then leave for HBM as whole contiguous segments (coalesced), not as one
  // 16-byte piece per thread.  The launcher sizes shared memory so the tile always fits.
  const int res_ld = BN / KS + 4;   // floats; +4 keeps the per-thread 16-byte accesses conflict free
  const bool res_in_smem = d.res != nullptr;

  if (tid == 0) B200_TR(3);
  MbarWait(bar_done, 0);
  TcFenceAfter();
  if (tid == 0) B200_TR(4);
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  // the BN / KS columns this CTA finishes, and this worker's share [c_lo, c_hi) of them
  const int CW = BN / KS;
  const int col_begin = rank * CW, col_end = col_begin + CW;
  const int n16 = CW >> 4, n16_lo = (n16 + 1) >> 1;
  const int c_lo = col_begin + (whalf ? n16_lo * 16 : 0), c_hi = whalf ? col_end : col_begin + n16_lo * 16;
  const uint32_t box_row_bytes = static_cast<uint32_t>(CW + 4) * 4;   // +16 B: conflict-free 16-byte row accesses
  float* res_s = reinterpret_cast<float*>(smem) + row * res_ld;
  if (res_in_smem) {
    const uint32_t dst = smem_base + row * res_ld * 4;
#pragma unroll 1
    for (int cc = c_lo - col_begin; cc < c_hi - col_begin; cc += 4)
      CpAsync16(dst + cc * 4, res_row + n0 + col_begin + cc, out_ok && (n0 + col_begin + cc) < N);
    CpAsyncCommit();
  }
  if (KS > 1) {
    // reduce-scatter, push half: the columns peer q finishes go to q's inbox; q's mbarrier counts the bytes
    if (tid == 0) MbarExpectTx(bar_box, static_cast<uint32_t>(KS - 1) * kTcM * CW * 4);
#pragma unroll 1
    for (int q = 1; q < KS; ++q) {
      const int pr = (rank + q) % KS;
      const int slot = rank < pr ? rank : rank - 1;
      const uint32_t dst = MapToCta(smem_base + box_off + (slot * kTcM + row) * box_row_bytes, pr);
      const uint32_t rbar = MapToCta(bar_box, pr);
#pragma unroll 1
      for (int h = 0; h < CW; h += 16) {
        if ((((q - 1) * n16 + (h >> 4)) & 1) != whalf) continue;   // (peer, group) pairs alternate between the two workers of a row
        uint32_t raw[16];
        TmemLd16(t_lane + pr * CW + h, raw);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          StAsync16(dst + (h + 4 * e) * 4, __uint_as_float(raw[4 * e]), __uint_as_float(raw[4 * e + 1]),
                    __uint_as_float(raw[4 * e + 2]), __uint_as_float(raw[4 * e + 3]), rbar);
      }
    }
    MbarWait(bar_box, 0);
  }
  if (res_in_smem) CpAsyncWait<0>();   // own row only: no block-level barrier needed
  if (tid == 0) B200_TR(5);
  const uint8_t* box_mine = smem + box_off + static_cast<size_t>(row) * box_row_bytes;
  // (epilogue loops are deliberately NOT unrolled: this code runs once per launch with a cold
  //  instruction cache, so every extra copy of the body is another exposed fetch from L2)
#pragma unroll 1
  for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
    uint32_t rr[16];
    TmemLd16(t_lane + c0, rr);   // whole warp, even when some rows are past M
    if (tid == 0 && c0 == col_begin) B200_TR(7);
    if (KS > 1) {   // own partial + the peers' partials, in rank order (deterministic)
      float acc[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = 0.0f;
#pragma unroll 1
      for (int src = 0; src < KS; ++src) {
        if (src == rank) {
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[e] += __uint_as_float(rr[e]);
        } else {
          const int slot = src < rank ? src : src - 1;
          const float4* bp = reinterpret_cast<const float4*>(box_mine + static_cast<size_t>(slot) * kTcM * box_row_bytes +
                                                             (c0 - col_begin) * 4);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 f = bp[e];
            acc[4 * e] += f.x; acc[4 * e + 1] += f.y; acc[4 * e + 2] += f.z; acc[4 * e + 3] += f.w;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) rr[e] = __float_as_uint(acc[e]);
    }
    if (!out_ok || n0 + c0 >= N) continue;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int col = n0 + c0 + 4 * g;
      float4 v = make_float4(__uint_as_float(rr[4 * g]), __uint_as_float(rr[4 * g + 1]), __uint_as_float(rr[4 * g + 2]),
                             __uint_as_float(rr[4 * g + 3]));
      if (col < N) {
        const float4 bz = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * g);
        v.x += bz.x; v.y += bz.y; v.z += bz.z; v.w += bz.w;
        if (film_row) {
          const int fc = col % d.film_C;
          const float4 ga = __ldg(reinterpret_cast<const float4*>(film_row + fc));
          const float4 be = __ldg(reinterpret_cast<const float4*>(film_row + d.film_C + fc));
          v.x = v.x * (1.0f + ga.x) + be.x;
          v.y = v.y * (1.0f + ga.y) + be.y;
          v.z = v.z * (1.0f + ga.z) + be.z;
          v.w = v.w * (1.0f + ga.w) + be.w;
        }
        if (res_row) {
          const float4 rz = *reinterpret_cast<const float4*>(res_s + (c0 - col_begin) + 4 * g);
          v.x += rz.x; v.y += rz.y; v.z += rz.z; v.w += rz.w;
        }
        if (d.out_act != kActNone) {
          v.x = ActTc(v.x, d.out_act);
          v.y = ActTc(v.y, d.out_act);
          v.z = ActTc(v.z, d.out_act);
          v.w = ActTc(v.w, d.out_act);
        }
      }
      *reinterpret_cast<float4*>(res_s + (c0 - col_begin) + 4 * g) = v;   // staged; leaves below
    }

  }
  // ---- write-out: warps w and w + 4 stream 16 each of their quarter's 32 staged rows, a whole row
  //      segment per lane group (both column halves of a row must be staged: worker-wide barrier) ----
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (tid == 0) B200_TR(104);
  {
    const float* wst = reinterpret_cast<const float*>(smem) + static_cast<size_t>((warp & 3) * 32) * res_ld;
    const int r_begin = (warp >> 2) * 16, r_end = r_begin + 16;
    const int colg = n0 + col_begin;
    if (d.y) {
      const int lpr = CW >> 2;                   // lanes per row (16 bytes each); CW <= 128
      const int rpp = 32 / lpr;                  // rows per pass
      const int sub = lane / lpr, c4 = (lane - sub * lpr) * 4;
      const unsigned long long my = reinterpret_cast<unsigned long long>(out_row);
#pragma unroll 1
      for (int r0 = r_begin; r0 < r_end; r0 += rpp) {
        const int j = (r0 + sub) & 31;
        const unsigned long long base = __shfl_sync(0xffffffffu, my, j);
        const int ok = __shfl_sync(0xffffffffu, out_ok ? 1 : 0, j);
        if (ok && sub < rpp && r0 + sub < r_end && colg + c4 < N)   // sub >= rpp: spare lanes when CW / 4 does not divide 32
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + colg + c4) =
              *reinterpret_cast<const float4*>(wst + j * res_ld + c4);
      }
    }
    if (d.yh) {   // bf16 consumer copy (hi [+ lo] planes), N a multiple of 8 wherever one exists
      const int lpr = CW >> 3;                   // lanes per row (8 channels = 16 bytes of bf16 each)
      const int rpp = 32 / lpr;
      const int sub = lane / lpr, c8 = (lane - sub * lpr) * 8;
#pragma unroll 1
      for (int r0 = r_begin; r0 < r_end; r0 += rpp) {
        const int j = (r0 + sub) & 31;
        const long long off = __shfl_sync(0xffffffffu, oh, j);
        const int ok = __shfl_sync(0xffffffffu, out_ok ? 1 : 0, j);
        if (ok && sub < rpp && r0 + sub < r_end && colg + c8 + 8 <= N) {
          float hv[8];
          const float4 a = *reinterpret_cast<const float4*>(wst + j * res_ld + c8);
          const float4 b = *reinterpret_cast<const float4*>(wst + j * res_ld + c8 + 4);
          hv[0] = a.x; hv[1] = a.y; hv[2] = a.z; hv[3] = a.w; hv[4] = b.x; hv[5] = b.y; hv[6] = b.z; hv[7] = b.w;
          if (d.yh_act == kActLrelu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) hv[e] = hv[e] > 0.0f ? hv[e] : 0.1f * hv[e];
          }
          uint4 h0, l0;
          Pack8<true>(hv, &h0, &l0);
          *reinterpret_cast<uint4*>(d.yh + off + colg + c8) = h0;
          if (d.yl) *reinterpret_cast<uint4*>(d.yl + off + colg + c8) = l0;
        }
      }
    }
  }
  if (tid == 0) B200_TR(6);
  }
  TcFenceBefore();
  __syncthreads();
  // (no exit-time cluster barrier: a CTA leaves only after every byte addressed to its inbox has been
  //  counted by its own mbarrier, and it sends nothing after its last st.async)
  if (tracing && tid == 0) {
    const long long t0 = trace[0];
    unsigned long long t_end_ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end_ns));
    printf("[tc trace] wall: dependency wait returned at %llu ns, CTA 0 done at %llu ns (+%llu)\n",
           static_cast<unsigned long long>(trace[120]), t_end_ns, t_end_ns - static_cast<unsigned long long>(trace[120]));
    printf("[tc trace] KS %d grid (%d,%d,%d) BN %d C_in %d k %d N %d T %d stages %d gather %d chunks %d | setup %lld pdl_wait %lld producers_done %lld mma_done %lld res_loaded %lld first_tmem_ld %lld staged %lld epi_done %lld end %lld\n",
           KS, gridDim.x, gridDim.y, gridDim.z, BN, C_in, d.k, N, d.T, kStages, kGather ? 1 : 0, n_chunks, trace[1] - t0, trace[2] - t0,
           trace[3] - t0, trace[4] - t0, trace[5] - t0, trace[7] - t0, trace[104] - t0, trace[6] - t0, clock64() - t0);
    for (int c = 0; c < c_end - c_begin && c < 24; ++c)
      printf("[tc trace]   chunk %d: issued %lld landed %lld | mma_full %lld mma_committed %lld\n", c, trace[8 + 4 * c] - t0,
             trace[9 + 4 * c] - t0, trace[10 + 4 * c] - t0, trace[11 + 4 * c] - t0);
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

bool KSplitDisabled() {
  static const bool off = [] {
    const char* e = std::getenv("BEATRICE_B200_NO_KSPLIT");
    return e && e[0] == '1';
  }();
  return off;
}

size_t TcStageBytes(bool split, int bn) {
  return (static_cast<size_t>(8) * kPanelA + static_cast<size_t>(bn) * kTcKC * 2) * (split ? 2 : 1);
}
// Pipeline depth.  A launch that fits in one wave (<= 148 CTAs) is latency bound: as deep as the K
// loop is long (<= 4) within the SM's shared memory.  Multi-wave launches (the C = 16 / 32 stages:
// hundreds of CTAs, 3-6 chunks) prefer occupancy: keep the CTA small enough for 2-3 per SM.
int TcStages(bool split, int bn, int n_chunks, int n_ctas) {
  const size_t stage = TcStageBytes(split, bn);
  const size_t cap = 222 * 1024;
  int s;
  if (n_ctas <= 148) {
    s = static_cast<int>(cap / stage);
  } else if (n_chunks <= 6) {
    s = (3 * stage <= 74 * 1024) ? 3 : 2;
  } else {
    s = static_cast<int>((110 * 1024) / stage);
    if (s < 3) s = static_cast<int>(cap / stage) >= 3 ? 3 : 2;
  }
  if (s > 4) s = 4;
  if (s > n_chunks) s = n_chunks;
  if (s < 2) s = 2;
  return s;
}

template <bool kSplit, int kStages, bool kGather>
void LaunchTcT(const ConvDesc& h0, const ConvDesc* d_descs, dim3 grid, size_t smem, int B, const int* d_frame, int ks,
               cudaStream_t s) {
  static bool attr_set[64] = {};
  int dev = 0;
  B200_CHECK(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    B200_CHECK(cudaFuncSetAttribute(conv_gemm_tc_kernel<kSplit, kStages, kGather>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
    attr_set[dev & 63] = true;
  }
  LaunchPdl(conv_gemm_tc_kernel<kSplit, kStages, kGather>, grid, dim3(kTcThreads, 1, 1), smem, s, ks, h0, d_descs, B, d_frame, ks);
}

}  // namespace

size_t PackWeightsTc(const float* w, int k, int C_in, int N, int bn_cap, int* bn_out, int* kc_out, uint16_t* hi,
                     uint16_t* lo) {
  // N tile: the whole (16-padded) N when it fits under the cap, else tiles of `bn_cap` columns.
  // Small caps trade activation re-reads for more CTAs on the layers that have few rows.
  const int n16 = (N + 15) / 16 * 16;
  if (bn_cap > 256) bn_cap = 256;
  const int bn = n16 <= bn_cap ? n16 : bn_cap;
  const int n_tiles = (N + bn - 1) / bn;
  const int n_sub = C_in >= kTcKC ? C_in / kTcKC : 1;
  const int tpc = C_in >= kTcKC ? 1 : kTcKC / C_in;
  const int n_chunks = C_in >= kTcKC ? k * n_sub : (k + tpc - 1) / tpc;
  if (bn_out) *bn_out = bn;
  if (kc_out) *kc_out = kTcKC;
  const size_t block = static_cast<size_t>(kTcKC) * bn;
  const size_t total = static_cast<size_t>(n_tiles) * n_chunks * block;
  if (!hi) return total * sizeof(uint16_t);
  for (int nt = 0; nt < n_tiles; ++nt)
    for (int c = 0; c < n_chunks; ++c) {
      const size_t base = (static_cast<size_t>(nt) * n_chunks + c) * block;
      for (int p = 0; p < kTcKC / 8; ++p)
        for (int n = 0; n < bn; ++n)
          for (int e = 0; e < 8; ++e) {
            const int kidx = p * 8 + e;
            int j, ci;
            if (C_in >= kTcKC) {
              j = c / n_sub;
              ci = (c - j * n_sub) * kTcKC + kidx;
            } else {
              j = c * tpc + kidx / C_in;
              ci = kidx % C_in;
            }
            const int col = nt * bn + n;
            float val = 0.f;
            if (j < k && col < N) val = w[(static_cast<size_t>(j) * C_in + ci) * N + col];
            const uint16_t h = Bf16Rn(val);
            const size_t o = base + (static_cast<size_t>(p) * bn + n) * 8 + e;
            hi[o] = h;
            if (lo) lo[o] = Bf16Rn(val - Bf16ToF(h));
          }
    }
  return total * sizeof(uint16_t);
}

void LaunchConvGemmTc(const ConvDesc* d_descs, const ConvDesc& h0, int nz, int B, const int* d_frame, bool split,
                      cudaStream_t s) {
  const int M = B * h0.T;
  const int bn = h0.tc_bn;
  const int n_tiles = (h0.N + bn - 1) / bn;
  const int kmax = nz > 1 ? 11 : h0.k;   // z-batched launches are the MRF branches k = 3, 7, 11
  const int n_chunks = h0.C_in >= kTcKC ? kmax * (h0.C_in / kTcKC) : (kmax + kTcKC / h0.C_in - 1) / (kTcKC / h0.C_in);
  const int m_tiles = (M + kTcM - 1) / kTcM;
  // K split over a cluster (see the kernel): worth it when the launch is a handful of CTAs walking a
  // long K loop -- the latency-bound layers with few rows per hop
  int ks = 1;
  const bool gather = h0.xh == nullptr;
  if (nz == 1 && !KSplitDisabled()) {
    // a gathered chunk (fp32 rings of three branches, summed and rounded in registers) costs ~2.7 us per CTA,
    // a cp.async chunk ~0.6 us: gathered launches split down to one chunk per CTA
    const int min_chunks = gather ? 1 : 2;
    // One wave only.  A B200 has 148 SMs in GPCs of 16-20: clusters of 2 pack all of them, clusters of 4 strand
    // 16 SMs (B300_MICROARCH.md, "CTAS_ACTIVE = {1: 148, 2: 148, 4: 132}") -- this 132 is that number, not an SM count.
    auto one_wave = [](int cluster) { return cluster >= 4 ? 132 : 148; };
    for (int cand : {4, 2}) {
      if (n_chunks >= min_chunks * cand && (bn / cand) % 16 == 0 && m_tiles * n_tiles * cand <= one_wave(cand)) {
        ks = cand;
        break;
      }
    }
  }
  dim3 grid(m_tiles * ks, n_tiles, nz);
  const int local_chunks = (n_chunks + ks - 1) / ks;
  const size_t inbox = ks > 1 ? static_cast<size_t>(ks - 1) * kTcM * (bn / ks + 4) * 4 : 0;
  int stages = TcStages(split, bn, local_chunks, static_cast<int>(grid.x * grid.y * grid.z));
  while (stages > 2 && stages * TcStageBytes(split, bn) + 128 + 2048 + inbox > 227 * 1024) --stages;
  const size_t tile = (static_cast<size_t>(kTcM) * (bn / ks + 4) * 4 + 127) / 128 * 128;   // epilogue staging tile
  size_t pipe = stages * TcStageBytes(split, bn);
  if (pipe < tile) pipe = tile;
  const size_t smem = pipe + 128 + 1024 + 1024 + inbox;   // + barriers, bias, trace, inbox
#define B200_TC_DISPATCH(SPLIT, GATHER)                                                    \
  do {                                                                                     \
    if (stages == 4) LaunchTcT<SPLIT, 4, GATHER>(h0, d_descs, grid, smem, B, d_frame, ks, s);      \
    else if (stages == 3) LaunchTcT<SPLIT, 3, GATHER>(h0, d_descs, grid, smem, B, d_frame, ks, s); \
    else LaunchTcT<SPLIT, 2, GATHER>(h0, d_descs, grid, smem, B, d_frame, ks, s);                  \
  } while (0)
  if (split) {
    if (gather) B200_TC_DISPATCH(true, true);
    else B200_TC_DISPATCH(true, false);
  } else {
    if (gather) B200_TC_DISPATCH(false, true);
    else B200_TC_DISPATCH(false, false);
  }
#undef B200_TC_DISPATCH
  B200_CHECK(cudaGetLastError());
}

}  // namespace b200
