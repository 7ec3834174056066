// Cluster form of the fused MRF branch kernel (b200_mrf.cu): one thread-block CLUSTER of NC CTAs
// runs one branch of one vocoder stage for a group of S streams, K-SPLIT across the cluster.
//
// Why: at C = 128 (stage 0, 5 rows per stream per hop, dilated history of 50 steps) the bf16
// hi+lo input panels of one 128-row group are ~0.6 MB -- no single CTA can hold them, and one CTA
// per (group, branch) would leave 118 of 148 SMs idle.  Here CTA `rank` of the cluster owns the
// channel slice [rank*C/NC, (rank+1)*C/NC) of every activation:
//   * its X / Y shared-memory panels, its conv histories and its slice of the weight stream hold
//     only those input channels, so the implicit GEMM of a conv is split along K;
//   * every CTA accumulates a PARTIAL D[128 rows x C] over its K slice in TMEM (tcgen05.mma, taps
//     are descriptor row shifts exactly as in the single-CTA kernel);
//   * the partials are reduce-scattered through DISTRIBUTED SHARED MEMORY: the epilogue warps read
//     the columns that belong to peer q from TMEM and push them (st.shared::cluster, fp32) into
//     q's inbox, signal q's mbarrier (release.cluster) and wait for their own inbox; the sum is
//     taken in rank order (deterministic), bias / residual / LeakyReLU applied, and the result is
//     -- by construction -- exactly the K slice this CTA needs as input of the next conv.
// Nothing but the stage input u, the branch output and the conv histories touches HBM.
#include <cstdio>
#include <cstring>
#include <vector>

#include "b200_common.h"
#include "b200_mrf.h"
#include "b200_tc_common.cuh"

namespace b200 {
namespace {

// weight ring: stages x K steps per chunk.  Every chunk costs the MMA warp an mbarrier wait and a commit,
// so at C = 128 (8 KB per K step, shared memory nearly full) 3 stages of 2 K steps beat 4 stages of 1.
__host__ __device__ constexpr int NstForC(int C) { return C >= 128 ? 3 : 4; }
// warps 0-7: epilogue.  Warps w and w + 4 share TMEM lane quarter w & 3 (tile rows 32 (w & 3) ..) and
// split its work items -- (tile, 16-channel group) and (peer, group) pairs -- by parity: the epilogue
// is latency bound per warp, two warps per SM sub-partition nearly halve it.
constexpr int kEpiWarps = 8;
constexpr int kQuarters = 4;
constexpr int kWarpMma = 8, kWarpW = 9, kWarpH = 10;
constexpr int kThreads = 11 * 32;
constexpr int kBars = 48;        // mbarrier slots reserved at the front of shared memory
constexpr int kHdr = 8 * kBars + 16;

__host__ __device__ constexpr int NkForC(int C) { return C >= 64 ? 2 : 4; }

template <int C, int NC, bool kSplit>
__global__ void __launch_bounds__(kThreads, 1) mrf_cluster_kernel(const __grid_constant__ MrfStageParams p) {
  constexpr int Cs = C / NC;                  // channels owned by this CTA (K slice and output slice)
  constexpr int Gs = Cs / 16;                 // own 16-channel groups == K steps per tap
  constexpr int PANs = Cs / 8;                // own 8-channel K panels
  constexpr int PAN = C / 8;                  // panels of the whole activation (history image)
  constexpr int NK = NkForC(C);               // K steps per weight chunk
  constexpr int kNst = NstForC(C);
  constexpr int P = kSplit ? 2 : 1;
  constexpr uint32_t kKstepBytes = P * C * 32;
  constexpr uint32_t kChunkBytes = NK * kKstepBytes;
  constexpr uint32_t kBoxRow = (Cs + 4) * 4;  // inbox row pitch (bytes); +16 keeps 16-byte row accesses conflict free
  static_assert(Cs % 16 == 0, "channel slice must be a multiple of one K step");
  extern __shared__ __align__(1024) uint8_t smem[];

  unsigned long long t_start_ns = 0;
  if (p.trace < 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start_ns));
  const int tid = threadIdx.x, lane = tid & 31;
  // warp index broadcast from lane 0: provably warp-uniform, so the role branches below are uniform
  // control flow and the single-thread MMA / TMA loops can live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int rank = static_cast<int>(ClusterCtaRank());
  const MrfBranchDesc& br = p.br[p.y2br[blockIdx.y]];
  const int k = br.k, T = p.T, S = p.S, MT = p.MT;
  const int group = blockIdx.x / NC;
  const int HX = (k - 1) * 5, HY = k - 1;            // history rows (time steps) in front of X / Y
  const int rows_valid = S * T;
  const int rows8 = (rows_valid + 7) & ~7;
  const int RX = HX * S + rows8, RY = HY * S + rows8;
  const int frame = *p.frame;

  // ---- shared memory carve-up ----
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = SmemAddr(bars);
  const uint32_t bar_w_full = bar0, bar_w_empty = bar0 + 8 * kNst, bar_hist = bar0 + 16 * kNst, bar_free = bar_hist + 16,
                 bar_box_full = bar_free + 16, bar_box_free = bar_box_full + 8 * kQuarters,
                 bar_in = bar_box_free + 8 * kQuarters, bar_acc = bar_in + 8 * MT * Gs;
  const int n_bars = 2 * kNst + 4 + 2 * kQuarters + MT * Gs + MT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * kBars);
  volatile uint32_t* in_cnt = reinterpret_cast<volatile uint32_t*>(smem + 8 * kBars + 4);   // += 1 per epilogue warp per conv input
  volatile uint32_t* acc_cnt = reinterpret_cast<volatile uint32_t*>(smem + 8 * kBars + 8);  // += 1 per conv whose MMAs retired
  float* bias_s = reinterpret_cast<float*>(smem + kHdr);          // [6][Cs]
  const uint32_t x_off = (kHdr + 6 * Cs * 4 + 127) / 128 * 128;
  const uint32_t x_pstride = static_cast<uint32_t>(RX) * 16, y_pstride = static_cast<uint32_t>(RY) * 16;
  const uint32_t x_plane = PANs * x_pstride, y_plane = PANs * y_pstride;
  const uint32_t y_off = x_off + P * x_plane;
  const uint32_t box_off = (y_off + P * y_plane + 127) / 128 * 128;
  const uint32_t box_slot = static_cast<uint32_t>(rows_valid) * kBoxRow;
  const uint32_t w_off = (box_off + (NC - 1) * box_slot + 2048 + 127) / 128 * 128;
  const uint32_t smem_base = SmemAddr(smem);
  const uint32_t x_base = smem_base + x_off, y_base = smem_base + y_off, box_base = smem_base + box_off,
                 w_base = smem_base + w_off;

  // developer trace (BEATRICE_B200_MRF_TRACE=1): clock64 stamps of one CTA, printed at exit
  long long* trace = reinterpret_cast<long long*>(smem + w_off + kNst * kChunkBytes);
  const bool tracing = p.trace != 0 && blockIdx.x == static_cast<unsigned>(p.trace - 1) && blockIdx.y == 0;
#define B200_TR(i, slot) do { if (tracing) trace[(i) * 16 + (slot)] = clock64(); } while (0)

  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(2 * MT * C + MT * Cs)) tmem_cols <<= 1;

  for (int i = tid; i < 6 * Cs; i += kThreads) bias_s[i] = __ldg(br.bias + (i / Cs) * C + rank * Cs + (i % Cs));
  if (tid == 0) {
    *in_cnt = 0;
    *acc_cnt = 0;
    for (int i = 0; i < n_bars; ++i) {
      const uint32_t b = bar0 + 8 * i;
      uint32_t count = 1;
      if (b >= bar_in && b < bar_acc) count = kQuarters;     // the four warps that produce a group
      if (b >= bar_box_free && b < bar_in) count = 2 * (NC - 1);   // both warps of the quarter, in every peer
      // box_full: one local arrive.expect_tx per phase (count 1)
      MbarInit(b, count);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWarpMma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemAddr(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tracing && tid < 128) trace[tid] = 0;
  TcFenceBefore();
  __syncthreads();
  if (tracing && tid == 0) trace[7 * 16] = clock64();
  ClusterSyncAll();   // every CTA's mbarriers exist before any peer signals them
  TcFenceAfter();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // same value in every lane, provably

  const size_t hist_unit = static_cast<size_t>(p.n_groups) * P * PAN * S * 8 * (k - 1);   // elements per unit dilation

  if (warp < kEpiWarps) {
    // =========================== epilogue warps ===========================
    const int q4 = warp & 3, whalf = warp >> 2, rtid = tid & 127;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
    const uint32_t x_col0 = 2 * MT * C;   // fp32 residual stream of the own channel slice
    // A lane quarter whose 32 rows lie entirely past the group's rows (S * T = 80 at stage 0: quarter 3 is empty) has
    // nothing to push or pull.  It must then stay OUT of the box_free / box_full hand-shake altogether: with no data
    // dependency on its peers nothing stops them from running a whole conv ahead, and a second round of box_free
    // arrivals landing before such a warp has observed the first one flips the barrier's parity back under it --
    // the wait then never returns (seen as a stall of the whole cluster about once per 3e4 hops under
    // tools/pipe_stress.py; the watchdog records pointed at exactly these warps).  Every CTA of a cluster has the
    // same rows, so the decision is the same in all of them and the barrier counts of the active quarters hold.
    const bool q_active = rows_valid > q4 * 32;
    // Everything up to here (barriers, TMEM, bias, and in the other warps the weight ring and the
    // history loads) touches nothing the preceding kernel -- the upsampler that writes u -- produces.
    PdlWait();
    PdlLaunchDependents();
    // ---- prologue: u -> TMEM (fp32 residual stream) and lrelu(u) -> X new rows ----
    for (int m = 0; m < MT; ++m) {
      const int r = m * 128 + rtid;
      const int t = r / S, s = r - t * S;
      const int b = group * S + s;
      const bool exists = r < rows_valid;
      const bool valid = exists && b < p.B;
      const float* urow = p.u + (static_cast<size_t>(b) * p.u_slots * T + (frame % p.u_slots) * T + t) * C + rank * Cs;
      const uint32_t srow = x_base + static_cast<uint32_t>(HX * S + r) * 16;
      const float* frow = p.film ? p.film + static_cast<size_t>(b) * 2 * C + rank * Cs : nullptr;
#pragma unroll 1
      for (int g = 0; g < Gs; ++g) {
        if (((m * Gs + g) & 1) != whalf) continue;
        float v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) f = __ldg(reinterpret_cast<const float4*>(urow + 16 * g + 4 * q));
          if (valid && frow) {   // FiLM of this vocoder stage (per stream): u * (1 + gamma) + beta
            const float4 ga = __ldg(reinterpret_cast<const float4*>(frow + 16 * g + 4 * q));
            const float4 be = __ldg(reinterpret_cast<const float4*>(frow + C + 16 * g + 4 * q));
            f.x = f.x * (1.0f + ga.x) + be.x;
            f.y = f.y * (1.0f + ga.y) + be.y;
            f.z = f.z * (1.0f + ga.z) + be.z;
            f.w = f.w * (1.0f + ga.w) + be.w;
          }
          v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
        }
        uint32_t raw[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) raw[e] = __float_as_uint(v[e]);
        TmemSt16(t_lane + x_col0 + m * Cs + 16 * g, raw);
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = v[e] > 0.0f ? v[e] : 0.1f * v[e];
        uint4 h0, l0, h1, l1;
        Pack8<kSplit>(v, &h0, &l0);
        Pack8<kSplit>(v + 8, &h1, &l1);
        if (exists) {
          const uint32_t a0 = srow + (2 * g) * x_pstride;
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0), "r"(h0.x), "r"(h0.y), "r"(h0.z), "r"(h0.w) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + x_pstride), "r"(h1.x), "r"(h1.y), "r"(h1.z), "r"(h1.w) : "memory");
          if (kSplit) {
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + x_plane), "r"(l0.x), "r"(l0.y), "r"(l0.z), "r"(l0.w) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + x_plane + x_pstride), "r"(l1.x), "r"(l1.y), "r"(l1.z), "r"(l1.w) : "memory");
          }
        }
        TmemStWait();
        FenceProxyAsync();
        TcFenceBefore();
        __syncwarp();
        if (lane == 0) MbarArrive(bar_in + 8 * (m * Gs + g));
      }
    }
    __syncwarp();
    if (lane == 0) SmemAddRelease(in_cnt);
    if (tid == 0) B200_TR(7, 1);
    // ---- the six convs ----
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
      const bool is_c1 = (i & 1) == 0;
      const bool last = i == 5;
      // destination of lrelu(.): the OTHER buffer (c1 -> Y, c2 -> X)
      const uint32_t d_base = is_c1 ? y_base : x_base;
      const uint32_t d_pstride = is_c1 ? y_pstride : x_pstride, d_plane = is_c1 ? y_plane : x_plane;
      const int d_hmax = is_c1 ? HY : HX;
      const float* bias = bias_s + i * Cs;
      if (i >= 1 && !last) MbarWait(bar_free + 8 * ((i + 1) & 1), ((i - 1) >> 1) & 1);
      // the peers have consumed what this warp pushed for conv i-1
      // (only quarters that own rows take part in the exchange -- see q_active)
      if (i >= 1 && q_active) MbarWaitClusterDbg(bar_box_free + 8 * q4, (i - 1) & 1, 200000u + 212u, static_cast<uint32_t>(i));
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int r = m * 128 + rtid;
        const int t = r / S, s = r - t * S;
        const int b = group * S + s;
        const bool exists = r < rows_valid;
        const bool valid = exists && b < p.B;
        MbarWait(bar_acc + 8 * m, i & 1);
        TcFenceAfter();
        if (tid == 0 && m == 0) B200_TR(i, 0);
        if (m == MT - 1 && tid == 0) atomicAdd(const_cast<uint32_t*>(acc_cnt), 1u);   // conv i's MMAs have all retired
        const uint32_t dcol = t_lane + ((i & 1) * MT + m) * C;
        // ---- reduce-scatter, push half: the columns of every peer go to that peer's inbox; the peer's
        //      mbarrier counts the bytes (st.async complete_tx), so no fence and no arrive is needed ----
        {
          const int w_rows = min(max(rows_valid - (m * 128 + q4 * 32), 0), 32);   // rows of this lane quarter in tile m
          if (lane == 0 && whalf == 0 && q_active) MbarExpectTx(bar_box_full + 8 * q4, static_cast<uint32_t>(NC - 1) * w_rows * Cs * 4);
        }
        // this warp's (peer, 16-column group) items are those of its parity; two at a time, so that two TMEM loads are in
        // flight under one wait and eight remote stores follow back to back
        if (q_active) {
          constexpr int kItems = (NC - 1) * Gs;
          auto target = [&](int it, uint32_t* dst, uint32_t* rbar, uint32_t* tcol) {
            const int q = 1 + it / Gs, h = it - (q - 1) * Gs;
            const int pr = (rank + q) % NC;
            const int slot = rank < pr ? rank : rank - 1;
            *dst = MapToCta(box_base + slot * box_slot + static_cast<uint32_t>(exists ? r : 0) * kBoxRow, pr) + 64 * h;
            *rbar = MapToCta(bar_box_full + 8 * q4, pr);
            *tcol = dcol + pr * Cs + 16 * h;
          };
#pragma unroll 1
          for (int it = whalf; it < kItems; it += 4) {
            uint32_t d0, b0, c0, d1 = 0, b1 = 0, c1 = 0;
            target(it, &d0, &b0, &c0);
            const bool two = it + 2 < kItems;
            uint32_t r0[16], r1[16];
            if (two) {
              target(it + 2, &d1, &b1, &c1);
              TmemLd16x2(c0, r0, c1, r1);
            } else {
              TmemLd16(c0, r0);
            }
            if (exists) {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                StAsync16(d0 + 16 * e, __uint_as_float(r0[4 * e]), __uint_as_float(r0[4 * e + 1]), __uint_as_float(r0[4 * e + 2]),
                          __uint_as_float(r0[4 * e + 3]), b0);
              if (two) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  StAsync16(d1 + 16 * e, __uint_as_float(r1[4 * e]), __uint_as_float(r1[4 * e + 1]), __uint_as_float(r1[4 * e + 2]),
                            __uint_as_float(r1[4 * e + 3]), b1);
              }
            }
          }
        }
        if (tid == 0 && m == 0) B200_TR(i, 1);
        // ---- pull half: own columns + the peers' partials, summed in rank order ----
        if (q_active) MbarWaitDbg(bar_box_full + 8 * q4, (i * MT + m) & 1, 200000u + 252u, static_cast<uint32_t>(i));
        if (tid == 0 && m == 0) B200_TR(i, 2);
        const uint32_t xcol = t_lane + x_col0 + m * Cs;
        const uint32_t srow = d_base + static_cast<uint32_t>(d_hmax * S + r) * 16;
        const uint8_t* box_row = smem + box_off + static_cast<size_t>(exists ? r : 0) * kBoxRow;
        float* orow = br.out + (static_cast<size_t>(b) * br.out_slots * T + (frame % br.out_slots) * T + t) * C + rank * Cs;
#pragma unroll 1
        for (int g = 0; g < Gs; ++g) {
          if (((m * Gs + g) & 1) != whalf) continue;
          uint32_t raw[16], xr[16];
          // own partial columns and (c2) the fp32 residual slice: both loads in flight, one wait
          if (!is_c1) TmemLd16x2(dcol + rank * Cs + 16 * g, raw, xcol + 16 * g, xr);
          else TmemLd16(dcol + rank * Cs + 16 * g, raw);
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = 0.0f;
#pragma unroll
          for (int src = 0; src < NC; ++src) {
            if (src == rank) {
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] += __uint_as_float(raw[e]);
            } else {
              const int slot = src < rank ? src : src - 1;
              const float4* bp = reinterpret_cast<const float4*>(box_row + slot * box_slot + 64 * g);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 f = bp[e];
                v[4 * e] += f.x; v[4 * e + 1] += f.y; v[4 * e + 2] += f.z; v[4 * e + 3] += f.w;
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias + 16 * g + 4 * q);
            v[4 * q] += b4.x;
            v[4 * q + 1] += b4.y;
            v[4 * q + 2] += b4.z;
            v[4 * q + 3] += b4.w;
          }
          if (!is_c1) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] += __uint_as_float(xr[e]);
            if (!last) {
#pragma unroll
              for (int e = 0; e < 16; ++e) xr[e] = __float_as_uint(v[e]);
              TmemSt16(xcol + 16 * g, xr);
            } else if (valid) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(orow + 16 * g + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
          }
          if (!last) {
            if (!valid) {
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = 0.0f;
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = v[e] > 0.0f ? v[e] : 0.1f * v[e];
            uint4 h0, l0, h1, l1;
            Pack8<kSplit>(v, &h0, &l0);
            Pack8<kSplit>(v + 8, &h1, &l1);
            if (exists) {
              const uint32_t a0 = srow + (2 * g) * d_pstride;
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0), "r"(h0.x), "r"(h0.y), "r"(h0.z), "r"(h0.w) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + d_pstride), "r"(h1.x), "r"(h1.y), "r"(h1.z), "r"(h1.w) : "memory");
              if (kSplit) {
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + d_plane), "r"(l0.x), "r"(l0.y), "r"(l0.z), "r"(l0.w) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0 + d_plane + d_pstride), "r"(l1.x), "r"(l1.y), "r"(l1.z), "r"(l1.w) : "memory");
              }
            }
            if (!is_c1) TmemStWait();
            FenceProxyAsync();
            TcFenceBefore();
            __syncwarp();
            if (lane == 0) MbarArrive(bar_in + 8 * (m * Gs + g));
          }
        }
      }
      if (tid == 0) B200_TR(i, 3);
      if (!last) {
        // this warp's inbox rows are consumed: the peers may push conv i+1
        __syncwarp();
        if (lane == 0 && q_active) {
          for (int q = 1; q < NC; ++q) MbarArriveCluster(MapToCta(bar_box_free + 8 * q4, (rank + q) % NC));
        }
        __syncwarp();
        if (lane == 0) SmemAddRelease(in_cnt);
      }
    }
  } else if (warp == kWarpMma) {
    // =========================== MMA issuer ===========================
    {
      // lean issue loop (see b200_mrf.cu): constant descriptor high words, running low words, ring counters
      const uint32_t idesc = MakeIdesc(C);
      const uint64_t w_tmpl = MakeDesc(0, C * 16, 128);
      const uint32_t w_tmpl_lo = static_cast<uint32_t>(w_tmpl), w_hi32 = static_cast<uint32_t>(w_tmpl >> 32);
      uint32_t stage = 0, within = 0, wphase = 0;
      uint32_t w_lo = w_tmpl_lo + ((w_base >> 4) & 0x3FFFu);   // a CTA of a cluster sees its window at a rank-dependent base: keep 14 bits
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        const int buf = i & 1, dil = ConvDil(i);
        const int hmax = buf ? HY : HX;
        const uint32_t bbase = buf ? y_base : x_base, pstride = buf ? y_pstride : x_pstride, plane = buf ? y_plane : x_plane;
        const uint64_t a_tmpl = MakeDesc(0, pstride, 128);
        const uint32_t a_tmpl_lo = static_cast<uint32_t>(a_tmpl), a_hi32 = static_cast<uint32_t>(a_tmpl >> 32);
        const uint32_t tap_step = static_cast<uint32_t>(dil * S);
        const uint32_t plane16 = plane >> 4, group_step = (2 * pstride) >> 4;
        MbarWait(bar_hist + 8 * buf, (i >> 1) & 1);
        if (lane == 0) B200_TR(i, 4);
#pragma unroll 1
        for (int m = 0; m < MT; ++m) {
          const uint32_t dcol = tmem_base + ((i & 1) * MT + m) * C;
          uint32_t acc = 0u;
          uint32_t a_group = a_tmpl_lo + ((bbase >> 4) & 0x3FFFu) + static_cast<uint32_t>((hmax - (k - 1) * dil) * S + 128 * m);
#pragma unroll 1
          for (int g = 0; g < Gs; ++g) {
            MbarWait(bar_in + 8 * (m * Gs + g), i & 1);
            TcFenceAfter();
            if (m == 0 && g == 0) if (lane == 0) B200_TR(i, 5);
            if (m == 0 && g == Gs - 1) if (lane == 0) B200_TR(i, 6);
            uint32_t a_lo = a_group;
#pragma unroll 1
            for (int j = 0; j < k; ++j) {
              if (within == 0) {
                MbarWait(bar_w_full + 8 * stage, wphase);
                TcFenceAfter();
              }
              MmaW2(dcol, a_lo, a_hi32, w_lo, w_hi32, idesc, acc);
              if (kSplit) {
                MmaW2(dcol, a_lo, a_hi32, w_lo + ((C * 32) >> 4), w_hi32, idesc, 1u);    // x_hi * W_lo
                MmaW2(dcol, a_lo + plane16, a_hi32, w_lo, w_hi32, idesc, 1u);            // x_lo * W_hi
              }
              acc = 1u;
              a_lo += tap_step;
              w_lo += kKstepBytes >> 4;
              ++within;
              if (within == NK || (g == Gs - 1 && j == k - 1)) {
                MmaCommitW(bar_w_empty + 8 * stage);
                within = 0;
                ++stage;
                w_lo = w_tmpl_lo + (((w_base + stage * kChunkBytes) >> 4) & 0x3FFFu);
                if (stage == kNst) {
                  stage = 0;
                  wphase ^= 1u;
                  w_lo = w_tmpl_lo + ((w_base >> 4) & 0x3FFFu);
                }
              }
            }
            a_group += group_step;
          }
          MmaCommitW(bar_acc + 8 * m);
          if (m == MT - 1) if (lane == 0) B200_TR(i, 7);
        }
      }
    }
    __syncwarp();
  } else if (warp == kWarpW) {
    // =========================== weight producer ===========================
    if (ElectOneSync()) {
      const int ksteps = k * Gs;                     // own K slice
      const int ksteps_conv = k * (C / 16);          // whole conv
      const int chunks = (ksteps + NK - 1) / NK;
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(br.w);
      uint32_t cc = 0;
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        const uint8_t* wconv = wsrc + (static_cast<size_t>(i) * ksteps_conv + static_cast<size_t>(rank) * ksteps) * kKstepBytes;
#pragma unroll 1
        for (int m = 0; m < MT; ++m) {
#pragma unroll 1
          for (int c = 0; c < chunks; ++c) {
            const uint32_t stage = cc % kNst, round = cc / kNst;
            if (round > 0) MbarWait(bar_w_empty + 8 * stage, (round - 1) & 1);
            const int n = min(NK, ksteps - c * NK);
            const uint32_t bytes = n * kKstepBytes;
            MbarExpectTx(bar_w_full + 8 * stage, bytes);
            TmaBulkLoadKeep(w_base + stage * kChunkBytes, wconv + static_cast<size_t>(c) * kChunkBytes, bytes, bar_w_full + 8 * stage);
            ++cc;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kWarpH) {
    // =========================== history mover ===========================
    if (ElectOneSync()) {
      auto hist_ptr = [&](int i, int H) {
        // conv i block: [group][plane][panel][H*S rows][8]; this CTA moves its own panels only
        return br.hist + hist_unit * DilPrefix(i) + static_cast<size_t>(group) * P * PAN * H * S * 8;
      };
      auto load_hist = [&](int i) {
        const int buf = i & 1, H = (k - 1) * ConvDil(i);
        const int hmax = buf ? HY : HX;
        const uint32_t bbase = buf ? y_base : x_base, pstride = buf ? y_pstride : x_pstride, plane = buf ? y_plane : x_plane;
        const uint32_t bytes = static_cast<uint32_t>(H) * S * 16;
        const uint16_t* src = hist_ptr(i, H);
        MbarExpectTx(bar_hist + 8 * buf, bytes * P * PANs);
        for (int pl = 0; pl < P; ++pl)
          for (int pn = 0; pn < PANs; ++pn)
            TmaBulkLoad(bbase + pl * plane + pn * pstride + static_cast<uint32_t>((hmax - H) * S) * 16,
                        src + static_cast<size_t>(pl * PAN + rank * PANs + pn) * H * S * 8, bytes, bar_hist + 8 * buf);
      };
      load_hist(0);
      load_hist(1);
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        const int buf = i & 1, H = (k - 1) * ConvDil(i);
        const int hmax = buf ? HY : HX;
        const uint32_t bbase = buf ? y_base : x_base, pstride = buf ? y_pstride : x_pstride, plane = buf ? y_plane : x_plane;
        // input of conv i complete: history landed + every new row written by the epilogue warps
        MbarWait(bar_hist + 8 * buf, (i >> 1) & 1);
        SpinUntil(in_cnt, static_cast<uint32_t>(kEpiWarps * (i + 1)), 200000u + 458u);
        __threadfence_block();
        FenceProxyAsync();
        {
          const uint32_t bytes = static_cast<uint32_t>(H) * S * 16;
          uint16_t* dst = const_cast<uint16_t*>(hist_ptr(i, H));
          for (int pl = 0; pl < P; ++pl)
            for (int pn = 0; pn < PANs; ++pn)
              TmaBulkStore(dst + static_cast<size_t>(pl * PAN + rank * PANs + pn) * H * S * 8,
                           bbase + pl * plane + pn * pstride + static_cast<uint32_t>((hmax + T - H) * S) * 16, bytes);
          BulkCommit();
          BulkWaitRead0();
        }
        B200_TR(i, 8);
        // conv i's MMAs done reading the buffer
        SpinUntil(acc_cnt, static_cast<uint32_t>(i + 1), 200000u + 473u);
        if (i + 2 < 6) load_hist(i + 2);
        MbarArrive(bar_free + 8 * buf);
        B200_TR(i, 9);
      }
      BulkWait0();
      B200_TR(7, 2);
    }
    __syncwarp();
  }

  TcFenceBefore();
  __syncthreads();
  if (tracing && tid == 0) {
    const long long t0 = trace[7 * 16];
    printf("[mrfc trace] C=%d NC=%d rank=%d k=%d S=%d MT=%d  prologue_done=%lld hist_drained=%lld end=%lld (cycles after init)\n", C, NC, rank, k, S,
           MT, trace[7 * 16 + 1] - t0, trace[7 * 16 + 2] - t0, clock64() - t0);
    for (int i = 0; i < 6; ++i)
      printf("[mrfc trace]  conv %d: hist_ready %lld in_g0 %lld in_gLast %lld mma_issued %lld | acc %lld pushed %lld box_full %lld epi_done %lld | tail_stored %lld hist_next %lld\n",
             i, trace[i * 16 + 4] - t0, trace[i * 16 + 5] - t0, trace[i * 16 + 6] - t0, trace[i * 16 + 7] - t0, trace[i * 16 + 0] - t0,
             trace[i * 16 + 1] - t0, trace[i * 16 + 2] - t0, trace[i * 16 + 3] - t0, trace[i * 16 + 8] - t0, trace[i * 16 + 9] - t0);
  }
  // (no exit-time cluster barrier: every byte and every signal addressed to this CTA has been consumed by
  //  its epilogue warps before they get here -- inbox bytes are counted by box_full, and the last
  //  box_free arrivals, for conv 4, were awaited before the pushes of conv 5)
  if (p.trace < 0 && tid == 0) {   // developer aid: residency window of every CTA (BEATRICE_B200_MRF_TRACE=-1)
    unsigned long long t_end_ns;
    uint32_t smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end_ns));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    printf("[%s cta] C %d bx %d by %d sm %u start_ns %llu end_ns %llu\n", "mrfc", C, blockIdx.x, blockIdx.y, smid, t_start_ns, t_end_ns);
  }
  if (warp == kWarpMma) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

template <int C, int NC, bool kSplit>
void LaunchClusterT(const MrfStageParams& p, size_t smem, cudaStream_t s) {
  static bool attr_set[64] = {};
  int dev = 0;
  B200_CHECK(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    B200_CHECK(cudaFuncSetAttribute(mrf_cluster_kernel<C, NC, kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev & 63] = true;
  }
  LaunchMaybePdl(!p.late_launch, mrf_cluster_kernel<C, NC, kSplit>, dim3(p.n_groups * NC, p.n_branches, 1), dim3(kThreads, 1, 1), smem, s,
                 NC, p);
}

}  // namespace

size_t MrfClusterSmemBytes(int C, int NC, int T, int S, bool split) {
  const int P = split ? 2 : 1, Cs = C / NC, PANs = Cs / 8;
  const int k = 11;   // the launch is sized for its largest branch
  const int rows_valid = S * T, rows8 = (rows_valid + 7) & ~7;
  const size_t RX = static_cast<size_t>((k - 1) * 5) * S + rows8, RY = static_cast<size_t>(k - 1) * S + rows8;
  size_t off = (kHdr + 6 * Cs * 4 + 127) / 128 * 128;
  off += P * PANs * RX * 16;
  off += P * PANs * RY * 16;
  off = (off + 127) / 128 * 128;
  off += static_cast<size_t>(NC - 1) * rows_valid * (Cs + 4) * 4 + 2048;   // inbox (+ slack the last tile's MMA may read into)
  off = (off + 127) / 128 * 128;
  off += static_cast<size_t>(NstForC(C)) * NkForC(C) * P * C * 32;
  off += 1024;   // developer trace area
  return off;
}

bool MrfClusterSupported(int C, int NC, int T, int S, bool split) {
  if (NC < 2 || NC > 8 || C % (16 * NC) != 0) return false;
  // instantiated forms; the C = 64 forms read the planar weight image, which PackMrfWeights emits only for
  // C = 128 in split mode (C <= 64 is packed for the single-CTA kernel's concatenated-N MMAs)
  if (!((C == 128 && NC == 4) || (!split && C == 64 && (NC == 2 || NC == 4)))) return false;
  const int MT = (S * T + 127) / 128;
  const int Cs = C / NC;
  if (2 * MT * C + MT * Cs > 512) return false;
  if (2 * NstForC(C) + 4 + 2 * kQuarters + MT * (Cs / 16) + MT > kBars) return false;
  return MrfClusterSmemBytes(C, NC, T, S, split) <= 227 * 1024;
}

void LaunchMrfStageCluster(const MrfStageParams& p, int C, int NC, bool split, cudaStream_t s) {
  const size_t smem = MrfClusterSmemBytes(C, NC, p.T, p.S, split);
#define B200_MRFC_CASE(CC, NN)                                      \
  if (C == CC && NC == NN) {                                        \
    if (split) LaunchClusterT<CC, NN, true>(p, smem, s);            \
    else LaunchClusterT<CC, NN, false>(p, smem, s);                 \
    return;                                                         \
  }
  B200_MRFC_CASE(128, 4)
  B200_MRFC_CASE(64, 2)
  B200_MRFC_CASE(64, 4)
#undef B200_MRFC_CASE
  Fail(-106, "cluster MRF kernel has no form for this width / cluster size", __FILE__, __LINE__);
}

void SetSpinDebugMrfc(unsigned long long* dev_ptr) { SetSpinDebugPtr(dev_ptr); }

}  // namespace b200
