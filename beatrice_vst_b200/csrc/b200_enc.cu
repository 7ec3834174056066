// Fused residual stack of the encoders (content encoder: 6 blocks at C = 256, pitch estimator: 3 blocks at
// C = 128), tcgen05 / TMEM / TMA + thread-block clusters, sm_100a only.
//
//     for r in blocks:  x = x + b_r + Conv_{k=3, dil=d_r}( GELU( ChanNorm(x) * gamma_r + beta_r ) )
//
// One hop adds ONE row per stream to this stack, so a block is a GEMM with 256 rows (streams) and
// K = 3C: far too little work to be anything but latency bound, and as separate launches (norm + conv
// per block) the stack was twelve dependent kernels.  Here one CLUSTER of NC = C/32 CTAs carries a tile of
// 32 streams through all blocks without leaving the SMs:
//
//  * TRANSPOSED GEMM.  D^T[C_out x 32 streams] = W^T[C_out x K] * G^T[K x 32]: the output channels fill
//    the 128 MMA rows (C/128 MMAs of M = 128, N = 32 per K step), the streams are the narrow N.
//  * K-SPLIT over the cluster.  CTA `rank` owns input channels [32 rank, 32 rank + 32) of every
//    activation: its slice of x (fp32, shared memory), of g (bf16 hi/lo K-panels with the block's causal
//    history in front: a tap is a row shift of the B descriptor), of the conv histories (TMA bulk copies)
//    and of the weight stream (TMA ring).  Its partial D^T covers all C output channels.
//  * REDUCE-SCATTER through distributed shared memory.  TMEM lane quarter q of M tile m holds output
//    channels 128 m + 32 q .. +32 == the slice CTA 4m + q owns: warp (m, q) reads its quarter and pushes it
//    (st.async, the receiver's mbarrier counts the bytes) into that CTA's inbox; sums in rank order.
//  * ChanNorm statistics: per stream over all C channels = over the cluster.  Every CTA reduces its 32
//    channels (exact two-pass inside a thread, Chan's parallel combination above), the 8-byte partials are
//    exchanged through DSMEM, every CTA combines them in rank order -> identical mean / rstd everywhere.
//  * Everything else (bias, residual add, normalise, GELU, bf16 hi/lo split) is elementwise on the CTA's
//    own [32 channels x 32 streams] slice, spread over all 256 worker threads.
//
// The same machinery runs every other one-row-per-hop conv of the model as further block kinds of the chain
// (ResStackParams in b200_enc.h): the last front-end layer in front of the stack (kind 1: k = 2, stride 2,
// GELU), the content encoder's 1x1 head behind it (kind 2), and -- as a chain of its own at C = 256 -- the
// vocoder's conditioning (kind 4: 1x1 phone embedding, then pitch-embedding row + feature projection +
// speaker / formant embeddings in the epilogue) followed by its `pre` conv (kind 3: k = 7, six history rows).
// The inbox is single-buffered; what keeps a fast CTA from pushing block r+1 into a peer that still reads
// block r is the statistics exchange of block r+1, which blocks without a ChanNorm run for that purpose alone
// (see the worker loop).
#include <cstdio>
#include <cstring>
#include <vector>

#include "b200_common.h"
#include "b200_enc.h"
#include "b200_tc_common.cuh"

namespace b200 {
namespace {

constexpr int kRs = 32;            // streams per cluster tile (the MMA's N)
constexpr int kCs = 32;            // channels owned by a CTA
constexpr int kHmax = 8;           // history time steps kept in front of the new row: 2 * max dilation
constexpr int kGpRows = (kHmax + 1) * kRs;
constexpr int kWorkers = 256;      // 8 warps: warp w <-> (M tile w / 4, TMEM lane quarter w % 4)
constexpr int kWarpMma = 8, kWarpW = 9, kWarpH = 10;
constexpr int kEncThreads = 11 * 32;
constexpr int kNstW = 6;           // weight ring stages, one K step (16 KB at C = 256) each: a whole block's slice in flight
constexpr int kBoxPitch = kRs + 4; // floats per inbox row (one channel x 32 streams), +16 B against bank conflicts

// Chan's pairwise combination of (count, mean, M2) when every partial has the same count g and the running
// value already holds k of them: nb / nt = 1 / (k + 1), n nb / nt = g k / (k + 1) -- compile-time constants
// once the loop over k is unrolled (no divisions on the critical path).
__device__ __forceinline__ void ChanCombineEq(int k, float g, float& mean, float& m2, float mb, float m2b) {
  const float inv = 1.0f / static_cast<float>(k + 1);
  const float delta = mb - mean;
  mean = fmaf(delta, inv, mean);
  m2 = m2 + m2b + delta * delta * (g * static_cast<float>(k) * inv);
}

// shared-memory layout (bytes), identical on host and device
template <int C>
struct EncSmem {
  static constexpr int NC = C / kCs;
  static constexpr int P = 2;                                   // hi + lo planes (the encoders always run split-bf16)
  static constexpr uint32_t kBars = 0;                          // 32 mbarriers
  static constexpr uint32_t kMisc = 256;                        // tmem slot, counters
  static constexpr uint32_t kXbuf = 512;                        // fp32 x slice [32 ch][32 streams]
  static constexpr uint32_t kStat = kXbuf + kCs * kRs * 4;      // [8 groups][32 streams][2] local partials, then [32][2] mean/rstd
  static constexpr uint32_t kStatBox = kStat + 8 * kRs * 8 + kRs * 8;   // [NC][32 streams][2] cluster partials
  static constexpr uint32_t kGp = (kStatBox + NC * kRs * 8 + 1023) / 1024 * 1024;
  static constexpr uint32_t kGpPanel = kGpRows * 16;            // one 8-channel K panel
  static constexpr uint32_t kGpPlane = (kCs / 8) * kGpPanel;
  static constexpr uint32_t kGpBuf = P * kGpPlane;              // one of two buffers (block parity)
  // inbox [NC slots][32 ch][kBoxPitch] fp32, single-buffered: a peer can push block r+1 only after it has this
  // CTA's statistics of block r+1, which are sent after the last read of block r's partials (a block without
  // ChanNorm that is not the first of its chain runs the exchange for this hand-shake alone)
  static constexpr uint32_t kBox = kGp + 2 * kGpBuf;
  static constexpr uint32_t kBoxSlot = kCs * kBoxPitch * 4;
  static constexpr uint32_t kBoxBuf = NC * kBoxSlot;
  static constexpr uint32_t kW = (kBox + kBoxBuf + 1023) / 1024 * 1024;
  static constexpr uint32_t kKstep = P * 2 * C * 16;            // one K step of weights: [plane][2 panels][C rows][8]
  static constexpr uint32_t kTotal = kW + kNstW * kKstep;
};

template <int C>
__global__ void __launch_bounds__(kEncThreads, 1) enc_res_stack_kernel(const __grid_constant__ ResStackParams p) {
  using L = EncSmem<C>;
  constexpr int NC = L::NC, P = L::P;
  constexpr int MTc = C / 128;                 // M tiles of output channels
  constexpr int Gs = kCs / 16;                 // K steps per tap (own slice)
  constexpr uint32_t kTmemCols = 64;           // MTc * 32 accumulator columns (power of two >= 32)
  extern __shared__ __align__(1024) uint8_t smem[];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int rank = static_cast<int>(ClusterCtaRank());
  const int tile = blockIdx.x / NC;
  const uint32_t smem_base = SmemAddr(smem);

  // barriers
  const uint32_t bar0 = smem_base + L::kBars;
  const uint32_t bar_w_full = bar0, bar_w_empty = bar0 + 8 * kNstW, bar_hist = bar0 + 16 * kNstW, bar_free = bar_hist + 16,
                 bar_in = bar_free + 16, bar_acc = bar_in + 8, bar_box = bar_acc + 8, bar_stat = bar_box + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kMisc);
  volatile uint32_t* in_cnt = reinterpret_cast<volatile uint32_t*>(smem + L::kMisc + 4);    // += 1 per worker warp per block input
  volatile uint32_t* acc_cnt = reinterpret_cast<volatile uint32_t*>(smem + L::kMisc + 8);   // += 1 per block whose MMAs retired
  float* xbuf = reinterpret_cast<float*>(smem + L::kXbuf);        // [lc][s]
  float* stat_loc = reinterpret_cast<float*>(smem + L::kStat);    // [cg][s][2], then mean/rstd at + 8*32*2
  float* stat_box = reinterpret_cast<float*>(smem + L::kStatBox); // [src rank][s][2]

  if (tid == 0) {
    *in_cnt = 0;
    *acc_cnt = 0;
    for (int i = 0; i < 2 * kNstW + 4; ++i) MbarInit(bar0 + 8 * i, 1);   // w_full, w_empty, hist[2], free[2]  (20 + 4 more fit in 256 B)
    MbarInit(bar_in, kWorkers / 32);
    MbarInit(bar_acc, 1);
    MbarInit(bar_box, 1);
    MbarInit(bar_stat, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWarpMma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemAddr(tmem_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  TcFenceBefore();
  __syncthreads();
  ClusterSyncAll();   // every CTA's barriers exist before a peer signals them
  TcFenceAfter();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // ---- per-block geometry (see ResStackParams) ----
  auto blk_taps = [&](int r) { return p.kind[r] == 0 ? 3 : (p.kind[r] == 1 ? 2 : (p.kind[r] == 3 ? 7 : 1)); };
  auto blk_rows = [&](int r) { return p.kind[r] == 2 ? p.head_n : C; };                  // output channels == A rows
  auto blk_hist = [&](int r) { return p.kind[r] == 0 ? 2 * p.dil[r] : (p.kind[r] == 3 ? 6 : 0); };   // history time steps
  auto blk_kstep = [&](int r) { return static_cast<uint32_t>(P * 2 * blk_rows(r) * 16); };   // bytes of one K step of weights
  // history image of block r: [tile][rank*P*4 + plane*4 + panel][H*32 rows][8]
  auto hist_ptr = [&](int r) {
    size_t off = 0;
    for (int i = 0; i < r; ++i) off += static_cast<size_t>(p.n_tiles) * NC * P * (kCs / 8) * blk_hist(i) * kRs * 8;
    return p.hist + off + (static_cast<size_t>(tile) * NC + rank) * P * (kCs / 8) * blk_hist(r) * kRs * 8;
  };

  __shared__ long long trace[8 * 8];
  const bool tracing = p.trace != 0 && blockIdx.x == 0;
#define B200_ETR(r, slot) do { if (tracing && tid == 0) trace[(r) * 8 + (slot)] = clock64(); } while (0)
  const long long t_origin = clock64();

  if (warp < kWorkers / 32) {
    // =========================== worker warps ===========================
    const int s = tid & 31, cg = tid >> 5;            // stream in the tile, group of 4 own channels
    const int b = tile * kRs + s;
    const bool valid = b < p.B;
    const int wm = warp >> 2, wq = warp & 3;          // (M tile, lane quarter) this warp drains
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    PdlWait();
    PdlLaunchDependents();
    // x slice of this CTA: [lc][s] (a kind-1 first block computes it instead)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int lc = cg * 4 + i;
      const int xc = p.x_in_C > 0 ? p.x_in_C : C;     // a kind-4 first block may have fewer input channels than C
      const int ch = rank * kCs + lc;
      xbuf[lc * kRs + s] = (valid && p.kind[0] != 1 && ch < xc) ? __ldg(p.x_in + static_cast<size_t>(b) * xc + ch) : 0.0f;
    }
    int stat_phase = 0;   // bar_stat completes one phase per kind-0 block
#pragma unroll 1
    for (int r = 0; r < p.n_blk; ++r) {
      const int buf = r & 1, kind = p.kind[r];
      const uint32_t gp = smem_base + L::kGp + buf * L::kGpBuf;
      // ---- ChanNorm statistics of x (block input), exact two-pass per thread, Chan above ----
      // per-block parameters of this thread's four channels: in flight while the statistics are exchanged
      const float4 ga = __ldg(reinterpret_cast<const float4*>(p.gamma + static_cast<size_t>(r) * C + rank * kCs + cg * 4));
      const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + static_cast<size_t>(r) * C + rank * kCs + cg * 4));
      const float4 bi = __ldg(reinterpret_cast<const float4*>(p.bias + static_cast<size_t>(r) * C + rank * kCs + cg * 4));
      float xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = xbuf[(cg * 4 + i) * kRs + s];
      B200_ETR(r, 0);
      float mean = 0.f, rstd = 0.f;
      // The exchange doubles as the cluster-wide hand-shake that makes the single inbox safe: a CTA sends its
      // partials of block r only after every peer's statistics of block r arrived, and a peer sends those only
      // after its last read of block r-1's inbox.  A head that follows another block has no statistics of its
      // own, so it runs the same exchange for the hand-shake alone.
      if (kind == 0 || ((kind == 2 || kind == 3) && r > 0)) {
      {
        const float m4 = ((xv[0] + xv[1]) + (xv[2] + xv[3])) * 0.25f;
        float q4 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) q4 = fmaf(xv[i] - m4, xv[i] - m4, q4);
        stat_loc[(cg * kRs + s) * 2] = m4;
        stat_loc[(cg * kRs + s) * 2 + 1] = q4;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (cg == 0) {   // 32 threads: this CTA's partial over its 32 channels, pushed to every peer
        float mean = stat_loc[s * 2], m2 = stat_loc[s * 2 + 1];
#pragma unroll
        for (int g = 1; g < 8; ++g) ChanCombineEq(g, 4.f, mean, m2, stat_loc[(g * kRs + s) * 2], stat_loc[(g * kRs + s) * 2 + 1]);
        stat_box[(rank * kRs + s) * 2] = mean;
        stat_box[(rank * kRs + s) * 2 + 1] = m2;
        if (lane == 0) MbarExpectTx(bar_stat, static_cast<uint32_t>(NC - 1) * kRs * 8);
        const uint32_t src = smem_base + L::kStatBox + (rank * kRs + s) * 8;
#pragma unroll 1
        for (int q = 1; q < NC; ++q) {
          const int pr = (rank + q) % NC;
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1,%2}, [%3];" ::"r"(MapToCta(src, pr)),
                       "f"(mean), "f"(m2), "r"(MapToCta(bar_stat, pr))
                       : "memory");
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // own partial visible to every worker
      MbarWait(bar_stat, stat_phase & 1);
      ++stat_phase;
      B200_ETR(r, 1);
      {
        float m2;
        mean = stat_box[s * 2];
        m2 = stat_box[s * 2 + 1];
#pragma unroll
        for (int src = 1; src < NC; ++src)
          ChanCombineEq(src, static_cast<float>(kCs), mean, m2, stat_box[(src * kRs + s) * 2], stat_box[(src * kRs + s) * 2 + 1]);
        rstd = rsqrtf(m2 * (1.0f / static_cast<float>(C)) + 1e-5f);
      }
      }   // statistics / hand-shake
      // ---- g = GELU(norm * gamma + beta) -> bf16 hi/lo into the new rows of the B panels ----
      if (r >= 2) MbarWait(bar_free + 8 * buf, ((r - 2) >> 1) & 1);   // the history mover is done with this buffer's block r-2
      if (kind == 1) {
        // last front-end layer: tap j reads input row j of the hop (k = 2, stride 2): rows -> time blocks kHmax-1, kHmax
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint2 hv = make_uint2(0u, 0u), lv = make_uint2(0u, 0u);
          if (valid) {
            const size_t o = (static_cast<size_t>(b) * 2 + j) * C + rank * kCs + cg * 4;
            hv = __ldg(reinterpret_cast<const uint2*>(p.fin_h + o));
            lv = __ldg(reinterpret_cast<const uint2*>(p.fin_l + o));
          }
          const uint32_t dst = gp + (cg >> 1) * L::kGpPanel + static_cast<uint32_t>((kHmax - 1 + j) * kRs + s) * 16 + (cg & 1) * 8;
          asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(dst), "r"(hv.x), "r"(hv.y) : "memory");
          asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(dst + L::kGpPlane), "r"(lv.x), "r"(lv.y) : "memory");
        }
      } else {
        float g[4];
        if (kind == 0) {
          g[0] = GeluFast((xv[0] - mean) * rstd * ga.x + be.x);
          g[1] = GeluFast((xv[1] - mean) * rstd * ga.y + be.y);
          g[2] = GeluFast((xv[2] - mean) * rstd * ga.z + be.z);
          g[3] = GeluFast((xv[3] - mean) * rstd * ga.w + be.w);
        } else {   // head: the operand is x itself
          g[0] = xv[0]; g[1] = xv[1]; g[2] = xv[2]; g[3] = xv[3];
        }
        if (!valid) g[0] = g[1] = g[2] = g[3] = 0.0f;
        // packed conversions (cvt.rn.bf16x2.f32); the residual against the bf16 read back as the upper half of an fp32
        const uint32_t hi0 = CvtBf16x2(g[0], g[1]), hi1 = CvtBf16x2(g[2], g[3]);
        const uint32_t lo0 = CvtBf16x2(g[0] - __uint_as_float(hi0 << 16), g[1] - __uint_as_float(hi0 & 0xffff0000u));
        const uint32_t lo1 = CvtBf16x2(g[2] - __uint_as_float(hi1 << 16), g[3] - __uint_as_float(hi1 & 0xffff0000u));
        // channel lc = 4 cg + i -> panel cg / 2, byte (cg & 1) * 8 inside the 16-byte row entry; row = newest time step
        const uint32_t dst = gp + (cg >> 1) * L::kGpPanel + static_cast<uint32_t>(kHmax * kRs + s) * 16 + (cg & 1) * 8;
        asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(dst), "r"(hi0), "r"(hi1) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(dst + L::kGpPlane), "r"(lo0), "r"(lo1) : "memory");
      }
      FenceProxyAsync();
      __syncwarp();
      if (lane == 0) {
        MbarArrive(bar_in);
        SmemAddRelease(in_cnt);
      }
      B200_ETR(r, 2);
      // ---- the block's MMAs run (warp 8); then reduce-scatter the partial D^T ----
      MbarWait(bar_acc, r & 1);
      TcFenceAfter();
      B200_ETR(r, 3);
      const int m_tiles = blk_rows(r) / 128;
      const int owners = blk_rows(r) / kCs;           // ranks that own 32 of this block's output channels
      if (tid == 0) {
        SmemAddRelease(acc_cnt);
        MbarExpectTx(bar_box, rank < owners ? static_cast<uint32_t>(NC - 1) * kCs * kRs * 4 : 0u);
      }
      if (wm < m_tiles) {
        const int dest = wm * 4 + wq;                 // owner of output channels 128 wm + 32 wq .. + 32
        const int slot = rank;                        // inbox slots are indexed by source rank
        const uint32_t row = smem_base + L::kBox + slot * L::kBoxSlot + static_cast<uint32_t>(lane) * kBoxPitch * 4;
        uint32_t raw[32];
        TmemLd16(t_lane + wm * kRs, raw);
        TmemLd16(t_lane + wm * kRs + 16, raw + 16);
        if (dest == rank) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + 16 * e), "r"(raw[4 * e]), "r"(raw[4 * e + 1]),
                         "r"(raw[4 * e + 2]), "r"(raw[4 * e + 3])
                         : "memory");
        } else {
          const uint32_t rrow = MapToCta(row, dest), rbar = MapToCta(bar_box, dest);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            StAsync16(rrow + 16 * e, __uint_as_float(raw[4 * e]), __uint_as_float(raw[4 * e + 1]), __uint_as_float(raw[4 * e + 2]),
                      __uint_as_float(raw[4 * e + 3]), rbar);
        }
      }
      TcFenceBefore();
      asm volatile("bar.sync 1, 256;" ::: "memory");   // own partial stored (and every TMEM read done before the next MMAs)
      B200_ETR(r, 4);
      MbarWait(bar_box, r & 1);
      B200_ETR(r, 5);
      // ---- x += bias + sum of the partials in rank order ----
      {
        const float* box = reinterpret_cast<const float*>(smem + L::kBox);
        const float bb[4] = {bi.x, bi.y, bi.z, bi.w};
        float yo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int lc = cg * 4 + i;
          float y = 0.f;
#pragma unroll
          for (int src = 0; src < NC; ++src) y += box[(src * kCs + lc) * kBoxPitch + s];
          yo[i] = y + bb[i];
          if (kind == 0) xbuf[lc * kRs + s] = xv[i] + yo[i];
          else if (kind == 1) xbuf[lc * kRs + s] = GeluFast(yo[i]);
          else if (kind == 3) xbuf[lc * kRs + s] = yo[i];
        }
        if (kind == 4) {   // conditioning: + pitch embedding row + feature projection + speaker (+ formant) embeddings
          const int ch = rank * kCs + cg * 4;
          float4 add = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) {
            const int qq = min(max(__ldg(p.emb_q + b), 0), p.emb_bins - 1);
            add = __ldg(reinterpret_cast<const float4*>(p.emb_pitch + static_cast<size_t>(qq) * C + ch));
            float4 fp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int f = 0; f < 4; ++f) {
              const float ft = __ldg(p.emb_feat + b * 4 + f);
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.emb_wf + static_cast<size_t>(f) * C + ch));
              fp.x = fmaf(ft, w4.x, fp.x); fp.y = fmaf(ft, w4.y, fp.y); fp.z = fmaf(ft, w4.z, fp.z); fp.w = fmaf(ft, w4.w, fp.w);
            }
            add.x += fp.x; add.y += fp.y; add.z += fp.z; add.w += fp.w;   // (acc + pitch) + feat, like cond_kernel
          }
          float v4[4] = {yo[0] + add.x, yo[1] + add.y, yo[2] + add.z, yo[3] + add.w};
          if (valid && p.emb_spk) {
            const float4 sp = *reinterpret_cast<const float4*>(p.emb_spk + static_cast<size_t>(b) * C + ch);
            v4[0] += sp.x; v4[1] += sp.y; v4[2] += sp.z; v4[3] += sp.w;
          }
          if (valid && p.emb_formant) {
            const float4 fm = *reinterpret_cast<const float4*>(p.emb_formant + static_cast<size_t>(b) * C + ch);
            v4[0] += fm.x; v4[1] += fm.y; v4[2] += fm.z; v4[3] += fm.w;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) xbuf[(cg * 4 + i) * kRs + s] = v4[i];
        }
        if (kind == 2 && rank < owners && valid) {
          const size_t o = static_cast<size_t>(b) * p.head_n + rank * kCs + cg * 4;
          *reinterpret_cast<float4*>(p.head_out + o) = make_float4(yo[0], yo[1], yo[2], yo[3]);
          if (p.head_out2) *reinterpret_cast<float4*>(p.head_out2 + o) = make_float4(yo[0], yo[1], yo[2], yo[3]);
        }
      }
      // (every thread re-reads only its own xbuf entries at the top of the next block: no barrier needed here)
      B200_ETR(r, 6);
    }
    if (tracing && tid == 0) {
      printf("[enc trace] C=%d blocks %d (cycles after kernel start)\n", C, p.n_blk);
      for (int r = 0; r < p.n_blk; ++r)
        printf("[enc trace]  block %d: start %lld stats %lld g_written %lld acc %lld pushed %lld box_full %lld summed %lld\n", r,
               trace[r * 8] - t_origin, trace[r * 8 + 1] - t_origin, trace[r * 8 + 2] - t_origin, trace[r * 8 + 3] - t_origin,
               trace[r * 8 + 4] - t_origin, trace[r * 8 + 5] - t_origin, trace[r * 8 + 6] - t_origin);
    }
    // ---- stack output: fp32 x for the record, bf16 hi/lo for the head conv ----
    if (valid && p.x_out) {
      float xo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xo[i] = xbuf[(cg * 4 + i) * kRs + s];
      const int n_slots = p.out_slots > 1 ? p.out_slots : 1;
      const int cur = n_slots > 1 ? (*p.frame % n_slots) : 0;
      const size_t o = (static_cast<size_t>(b) * n_slots + cur) * C + rank * kCs + cg * 4;
      *reinterpret_cast<float4*>(p.x_out + o) = make_float4(xo[0], xo[1], xo[2], xo[3]);
      if (p.out_act == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) xo[i] = xo[i] > 0.0f ? xo[i] : 0.1f * xo[i];
      }
      if (p.xh_out) {
        uint2 hv, lv;
        hv.x = CvtBf16x2(xo[0], xo[1]);
        hv.y = CvtBf16x2(xo[2], xo[3]);
        *reinterpret_cast<uint2*>(p.xh_out + o) = hv;
        if (p.xl_out) {
          lv.x = CvtBf16x2(xo[0] - __uint_as_float(hv.x << 16), xo[1] - __uint_as_float(hv.x & 0xffff0000u));
          lv.y = CvtBf16x2(xo[2] - __uint_as_float(hv.y << 16), xo[3] - __uint_as_float(hv.y & 0xffff0000u));
          *reinterpret_cast<uint2*>(p.xl_out + o) = lv;
        }
      }
    }
  } else if (warp == kWarpMma) {
    // =========================== MMA issuer (warp-converged, elected issue) ===========================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kRs >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    uint32_t cc = 0;
#pragma unroll 1
    for (int r = 0; r < p.n_blk; ++r) {
      const int buf = r & 1, kind = p.kind[r], dil = p.dil[r];
      const int taps = blk_taps(r), rows_a = blk_rows(r), m_tiles = rows_a / 128;
      const uint32_t gp = smem_base + L::kGp + buf * L::kGpBuf;
      MbarWait(bar_hist + 8 * buf, (r >> 1) & 1);
      MbarWait(bar_in, r & 1);
      TcFenceAfter();
      int ks = 0;
#pragma unroll 1
      for (int j = 0; j < taps; ++j) {
        // time block the tap reads: residual block t - (2 - j) dil; front conv: hop row j; head: the current row
        const int tb = kind == 0 ? kHmax - (2 - j) * dil : (kind == 1 ? kHmax - 1 + j : (kind == 3 ? kHmax - (6 - j) : kHmax));
        const uint32_t row0 = static_cast<uint32_t>(tb * kRs);
#pragma unroll 1
        for (int h = 0; h < Gs; ++h) {
          const uint32_t stage = cc % kNstW;
          MbarWait(bar_w_full + 8 * stage, (cc / kNstW) & 1);
          TcFenceAfter();
          const uint32_t w_s = smem_base + L::kW + stage * L::kKstep;
          const uint32_t b_hi = gp + (2 * h) * L::kGpPanel + row0 * 16;
          const uint64_t bh = MakeDesc(b_hi, L::kGpPanel, 128);
          const uint64_t bl = MakeDesc(b_hi + L::kGpPlane, L::kGpPanel, 128);
#pragma unroll 1
          for (int m = 0; m < m_tiles; ++m) {
            const uint64_t ah = MakeDesc(w_s + m * 128 * 16, rows_a * 16, 128);
            const uint64_t al = MakeDesc(w_s + 2 * rows_a * 16 + m * 128 * 16, rows_a * 16, 128);
            const uint32_t dcol = tmem_base + m * kRs;
            MmaW(dcol, ah, bh, idesc, ks > 0 ? 1u : 0u);
            MmaW(dcol, ah, bl, idesc, 1u);
            MmaW(dcol, al, bh, idesc, 1u);
          }
          ++ks;
          MmaCommitW(bar_w_empty + 8 * stage);
          ++cc;
        }
      }
      MmaCommitW(bar_acc);
    }
    __syncwarp();
  } else if (warp == kWarpW) {
    // =========================== weight producer ===========================
    if (ElectOneSync()) {
      uint32_t cc = 0;
      size_t w_off = 0;   // byte offset of block r in the image: [block][rank][tap][K step]
#pragma unroll 1
      for (int r = 0; r < p.n_blk; ++r) {
        const int ksteps = blk_taps(r) * Gs;   // own K slice of the block
        const uint32_t kb = blk_kstep(r);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + w_off + static_cast<size_t>(rank) * ksteps * kb;
#pragma unroll 1
        for (int c = 0; c < ksteps; ++c) {
          const uint32_t stage = cc % kNstW, round = cc / kNstW;
          if (round > 0) MbarWait(bar_w_empty + 8 * stage, (round - 1) & 1);
          MbarExpectTx(bar_w_full + 8 * stage, kb);
          TmaBulkLoadKeep(smem_base + L::kW + stage * L::kKstep, wsrc + static_cast<size_t>(c) * kb, kb, bar_w_full + 8 * stage);
          ++cc;
        }
        w_off += static_cast<size_t>(NC) * ksteps * kb;
      }
    }
    __syncwarp();
  } else if (warp == kWarpH) {
    // =========================== history mover ===========================
    if (ElectOneSync()) {
      auto load_hist = [&](int r) {
        const int buf = r & 1, H = blk_hist(r);
        if (H == 0) {   // no history (front conv, head): just complete the buffer's phase
          MbarArrive(bar_hist + 8 * buf);
          return;
        }
        const uint32_t gp = smem_base + L::kGp + buf * L::kGpBuf;
        const uint32_t bytes = static_cast<uint32_t>(H) * kRs * 16;
        const uint16_t* src = hist_ptr(r);
        MbarExpectTx(bar_hist + 8 * buf, bytes * P * (kCs / 8));
        for (int pl = 0; pl < P; ++pl)
          for (int pn = 0; pn < kCs / 8; ++pn)
            TmaBulkLoad(gp + pl * L::kGpPlane + pn * L::kGpPanel + static_cast<uint32_t>((kHmax - H) * kRs) * 16,
                        src + static_cast<size_t>(pl * (kCs / 8) + pn) * H * kRs * 8, bytes, bar_hist + 8 * buf);
      };
      load_hist(0);
      if (p.n_blk > 1) load_hist(1);
#pragma unroll 1
      for (int r = 0; r < p.n_blk; ++r) {
        const int buf = r & 1, H = blk_hist(r);
        const uint32_t gp = smem_base + L::kGp + buf * L::kGpBuf;
        MbarWait(bar_hist + 8 * buf, (r >> 1) & 1);
        SpinUntil(in_cnt, static_cast<uint32_t>((kWorkers / 32) * (r + 1)), 300000u + 492u);   // the new row of block r is in the panels
        __threadfence_block();
        FenceProxyAsync();
        if (H > 0) {
          const uint32_t bytes = static_cast<uint32_t>(H) * kRs * 16;
          uint16_t* dst = const_cast<uint16_t*>(hist_ptr(r));
          for (int pl = 0; pl < P; ++pl)
            for (int pn = 0; pn < kCs / 8; ++pn)
              TmaBulkStore(dst + static_cast<size_t>(pl * (kCs / 8) + pn) * H * kRs * 8,
                           gp + pl * L::kGpPlane + pn * L::kGpPanel + static_cast<uint32_t>((kHmax + 1 - H) * kRs) * 16, bytes);
          BulkCommit();
          BulkWaitRead0();
        }
        SpinUntil(acc_cnt, static_cast<uint32_t>(r + 1), 300000u + 505u);   // block r's MMAs are done reading the buffer
        if (r + 2 < p.n_blk) load_hist(r + 2);
        MbarArrive(bar_free + 8 * buf);
      }
      BulkWait0();
    }
    __syncwarp();
  }

  TcFenceBefore();
  __syncthreads();
  if (warp == kWarpMma) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

template <int C>
void LaunchResStackT(const ResStackParams& p, cudaStream_t s) {
  static bool attr_set[64] = {};
  int dev = 0;
  B200_CHECK(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    B200_CHECK(cudaFuncSetAttribute(enc_res_stack_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(EncSmem<C>::kTotal)));
    attr_set[dev & 63] = true;
  }
  constexpr int NC = C / kCs;
  LaunchPdl(enc_res_stack_kernel<C>, dim3(p.n_tiles * NC, 1, 1), dim3(kEncThreads, 1, 1), EncSmem<C>::kTotal, s, NC, p);
}

}  // namespace

bool ResStackSupported(int C, int n_res, const int* dil) {
  if (C != 128 && C != 256) return false;
  if (n_res < 1 || n_res > 6) return false;
  for (int r = 0; r < n_res; ++r)
    if (dil[r] < 1 || 2 * dil[r] > kHmax) return false;
  return (C == 256 ? EncSmem<256>::kTotal : EncSmem<128>::kTotal) <= 227 * 1024;
}

int ResStackTiles(int B) { return (B + kRs - 1) / kRs; }

size_t ResStackHistElems(int C, int n_res, const int* dil, int B) {
  size_t n = 0;
  for (int r = 0; r < n_res; ++r) n += static_cast<size_t>(ResStackTiles(B)) * (C / kCs) * 2 * (kCs / 8) * (2 * dil[r]) * kRs * 8;
  return n;
}

// One MrfHistBlock per residual block (layout [tile][rank*8 + plane*4 + panel][H][32 streams][8], which is the
// [group][planes_panels][H][S][8] form mrf_zero_stream_kernel clears per stream).
void ResStackHistBlocks(int C, int n_res, const int* dil, int B, uint16_t* base, std::vector<MrfHistBlock>* out) {
  size_t off = 0;
  for (int r = 0; r < n_res; ++r) {
    MrfHistBlock hb;
    hb.base = base + off;
    hb.planes_panels = (C / kCs) * 2 * (kCs / 8);
    hb.H = 2 * dil[r];
    hb.S = kRs;
    hb.pad_ = 0;
    out->push_back(hb);
    off += static_cast<size_t>(ResStackTiles(B)) * hb.planes_panels * hb.H * kRs * 8;
  }
}

size_t PackChainWeights(const ChainLayer* layers, int n_layers, int C, uint16_t* out) {
  // [block r][rank][tap j][K step h][plane][2 panels][n_out rows n][8]   element = W_r[j][32 rank + 16 h + 8 pp + e][n]
  const int NC = C / kCs, Gs = kCs / 16;
  size_t total = 0;
  for (int r = 0; r < n_layers; ++r) total += static_cast<size_t>(NC) * layers[r].taps * Gs * (static_cast<size_t>(2) * 2 * layers[r].n_out * 8);
  if (!out) return total;
  size_t base = 0;
  for (int r = 0; r < n_layers; ++r) {
    const int taps = layers[r].taps, N = layers[r].n_out;
    const size_t kstep = static_cast<size_t>(2) * 2 * N * 8;
    for (int rk = 0; rk < NC; ++rk)
      for (int j = 0; j < taps; ++j)
        for (int h = 0; h < Gs; ++h) {
          uint16_t* blk = out + base + ((static_cast<size_t>(rk) * taps + j) * Gs + h) * kstep;
          for (int pp = 0; pp < 2; ++pp)
            for (int n = 0; n < N; ++n)
              for (int e = 0; e < 8; ++e) {
                const int ci = rk * kCs + 16 * h + 8 * pp + e;
                const float val = layers[r].w[(static_cast<size_t>(j) * C + ci) * N + n];
                const uint16_t hi = Bf16Rn(val);
                const size_t o = (static_cast<size_t>(pp) * N + n) * 8 + e;
                blk[o] = hi;
                blk[static_cast<size_t>(2) * N * 8 + o] = Bf16Rn(val - Bf16ToF(hi));
              }
        }
    base += static_cast<size_t>(NC) * taps * Gs * kstep;
  }
  return total;
}

void LaunchResStack(const ResStackParams& p, int C, cudaStream_t s) {
  // the single-inbox hand-shake (see the kernel) covers residual blocks, heads and the pre conv anywhere, the
  // front conv and the conditioning block only as the first block of a chain
  for (int r = 1; r < p.n_blk; ++r)
    if (p.kind[r] == 1 || p.kind[r] == 4) Fail(-103, "chain block of kind 1 / 4 must be the first of its chain", __FILE__, __LINE__);
  if (C == 256) LaunchResStackT<256>(p, s);
  else if (C == 128) LaunchResStackT<128>(p, s);
  else Fail(-104, "residual-stack kernel has no form for this width", __FILE__, __LINE__);
  B200_CHECK(cudaGetLastError());
}

void SetSpinDebugEnc(unsigned long long* dev_ptr) { SetSpinDebugPtr(dev_ptr); }

}  // namespace b200
