// Fused MRF branch kernel (vocoder stages with C <= 64), tcgen05 / TMEM / TMA, sm_100a only.
//
// One CTA runs ONE branch (kernel size k in {3,7,11}) of one vocoder stage for a group of S
// streams, all six convolutions of the branch back to back:
//     x0 = u;  for d in {1,3,5}:  a = c1_d(lrelu(x)) ; x = x + c2_d(lrelu(a))        (oracle: Generate(), MRF loop)
// Nothing but the stage input u, the branch output and the conv histories touches HBM.
//
//  * Rows are TIME-MAJOR inside the group:  row = t * S + s.  The input of every conv sits in shared
//    memory as bf16 K-panels [C/8][rows][8] (UMMA canonical K-major, no swizzle, 16 B per row per
//    panel) with the causal history in front of the hop's new rows, so tap j of a dilated conv is
//    the SAME buffer at a row offset of (k-1-j)*dil*S rows: an implicit GEMM with no im2col and no
//    per-tap data movement -- the tap is a shift of the descriptor's start address.
//  * Accumulators live in TMEM.  The residual stream x stays in TMEM as fp32 for the whole chain:
//    c2's MMAs accumulate straight onto it.  The epilogue warps read 16 columns at a time
//    (tcgen05.ld), add the bias, write x back (tcgen05.st), and store lrelu(.) as bf16 (hi [+ lo])
//    panels for the next conv; each 16-channel group is handed to the MMA warp through its own
//    mbarrier, so the next conv starts on channel group 0 while the epilogue is still on group 1.
//  * Weights stream through a 4-stage ring of TMA bulk copies (one warp), K ordered
//    [channel group][tap] to match that hand-off.  A third control warp moves the conv histories:
//    TMA bulk loads global -> shared before a conv, bulk stores of the new tail shared -> global
//    after its input is complete.
//  * Precision: plain bf16 (1 MMA per K step) or split bf16 (x = hi + lo for both operands,
//    hi*hi + hi*lo + lo*hi, fp32 accumulate).  In split mode the weight rows of a K step are packed
//    [W_hi ; W_lo] so that ONE MMA of width N = 2C yields x_hi*W_hi (accumulator columns [0,C)) and
//    x_hi*W_lo (columns [C,2C)), and a second MMA of width C adds x_lo*W_hi onto columns [0,C): two
//    MMAs instead of three, and the x_hi panel -- shared-memory reads bound these narrow-N MMAs --
//    is read once instead of twice.  The epilogue adds the two column halves.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "b200_common.h"
#include "b200_mrf.h"
#include "b200_tc_common.cuh"

namespace b200 {
namespace {

// Weight ring: K steps per chunk and stages.  Every chunk costs the MMA warp one mbarrier wait (~90
// cycles even when already complete) and one tcgen05.commit, so chunks are as large as the CTA's
// shared-memory budget allows (C = 64: one CTA per SM, 16 KB chunks; C = 32: two per SM, 8 KB; C = 16:
// four per SM, 6 KB) while the bytes in flight still cover the ring's ~1.6k-cycle round trip.
__host__ __device__ constexpr int NkFor(int C) { return C >= 128 ? 1 : (C <= 16 ? 6 : 4); }
__host__ __device__ constexpr int NstFor(int C) { return C >= 64 ? 4 : 3; }
// Epilogue warps: 8 where a tile has more than one work item -- (tile, 16-channel group) pairs --
// per conv, else 4.  Warps w and w + 4 share TMEM lane quarter w & 3 (tile rows 32 (w & 3) ..) and take
// the items of one parity each: the epilogue is latency bound per warp, two warps per SM
// sub-partition nearly halve it.  Then one MMA-issuing, one weight and one history warp.
__host__ __device__ constexpr int EpiWarpsFor(int C) { return C >= 32 ? 8 : 4; }
__host__ __device__ constexpr int ThreadsFor(int C) { return (EpiWarpsFor(C) + 3) * 32; }
constexpr int kQuarters = 4;

template <int C>
struct MrfCfg {
  static constexpr int kG = C / 16;                       // 16-channel groups == K steps per tap
  static constexpr int kPan = C / 8;                      // 8-channel K panels
  static constexpr int kNk = NkFor(C);                    // K steps per weight chunk
};

template <int C, bool kSplit>
__global__ void __launch_bounds__(ThreadsFor(C), C <= 16 ? 4 : (C <= 32 ? 2 : 1)) mrf_branch_kernel(const __grid_constant__ MrfStageParams p) {
  constexpr int kEpiWarps = EpiWarpsFor(C), kThreads = ThreadsFor(C);
  constexpr int kWarpMma = kEpiWarps, kWarpW = kEpiWarps + 1, kWarpH = kEpiWarps + 2;
  using Cfg = MrfCfg<C>;
  constexpr int G = Cfg::kG, PAN = Cfg::kPan, NK = Cfg::kNk;
  constexpr int P = kSplit ? 2 : 1;
  constexpr int DW = kSplit ? 2 * C : C;          // accumulator columns per 128-row tile (see "Precision")
  constexpr uint32_t kKstepBytes = P * C * 32;
  constexpr uint32_t kChunkBytes = NK * kKstepBytes;
  constexpr int kNst = NstFor(C);
  extern __shared__ __align__(1024) uint8_t smem[];

  unsigned long long t_start_ns = 0;
  if (p.trace < 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start_ns));
  const int tid = threadIdx.x, lane = tid & 31;
  // warp index broadcast from lane 0: provably warp-uniform, so the role branches below are uniform
  // control flow and the single-thread MMA / TMA loops can live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const MrfBranchDesc& br = p.br[p.y2br[blockIdx.y]];
  const int k = br.k, T = p.T, S = p.S, MT = p.MT;
  const int group = blockIdx.x;
  const int HX = (k - 1) * 5, HY = k - 1;            // history rows (time steps) in front of X / Y
  // rows of the X / Y panels: history + the hop's rows (8-row granules).  The last tile's MMA may read up
  // to 127 rows past that -- into the next panel, plane or buffer, all mapped shared memory; rows of an
  // operand only ever reach the same rows of the accumulator, and those rows are never read back.
  const int rows8 = (S * T + 7) & ~7;
  const int RX = HX * S + rows8, RY = HY * S + rows8;
  const int frame = *p.frame;

  // ---- shared memory carve-up ----
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // [0,NST) w_full  [NST,2NST) w_empty  [2NST,+2) hist_full  [+2,+4) buf_free  then in_ready[MT*G], acc_ready[MT]
  const uint32_t bar0 = SmemAddr(bars);
  const uint32_t bar_w_full = bar0, bar_w_empty = bar0 + 8 * kNst, bar_hist = bar0 + 16 * kNst,
                 bar_free = bar_hist + 16, bar_in = bar_free + 16, bar_acc = bar_in + 8 * MT * G;
  const int n_bars = 2 * kNst + 4 + MT * G + MT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * 40);
  // hand-off counters for the history mover (monotonic, so a late reader can never alias a phase)
  volatile uint32_t* in_cnt = reinterpret_cast<volatile uint32_t*>(smem + 8 * 40 + 4);   // += 1 per epilogue warp per conv input
  volatile uint32_t* acc_cnt = reinterpret_cast<volatile uint32_t*>(smem + 8 * 40 + 8);  // += 1 per conv whose MMAs retired
  float* bias_s = reinterpret_cast<float*>(smem + 8 * 40 + 16);          // [6][C]
  const uint32_t x_off = (8 * 40 + 16 + 6 * C * 4 + 127) / 128 * 128;
  const uint32_t x_pstride = static_cast<uint32_t>(RX) * 16, y_pstride = static_cast<uint32_t>(RY) * 16;
  const uint32_t x_plane = PAN * x_pstride, y_plane = PAN * y_pstride;
  const uint32_t y_off = x_off + P * x_plane;
  const uint32_t w_off = (y_off + P * y_plane + 127) / 128 * 128;
  const uint32_t smem_base = SmemAddr(smem);
  const uint32_t x_base = smem_base + x_off, y_base = smem_base + y_off, w_base = smem_base + w_off;

  // developer trace (BEATRICE_B200_MRF_TRACE=1 + blockIdx.x): clock64 stamps of one CTA, printed at exit
  long long* trace = reinterpret_cast<long long*>(smem + w_off + kNst * kChunkBytes);
  const bool tracing = p.trace != 0 && blockIdx.x == static_cast<unsigned>(p.trace - 1) && blockIdx.y == 0;
#define B200_TR(i, slot) do { if (tracing) trace[(i) * 16 + (slot)] = clock64(); } while (0)
  if (tracing && tid < 128) trace[tid] = 0;

  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(2 * MT * DW)) tmem_cols <<= 1;

  for (int i = tid; i < 6 * C; i += kThreads) bias_s[i] = __ldg(br.bias + i);
  if (tid == 0) {
    *in_cnt = 0;
    *acc_cnt = 0;
    for (int i = 0; i < n_bars; ++i) {
      const uint32_t b = bar0 + 8 * i;
      const bool is_in = b >= bar_in && b < bar_acc;
      MbarInit(b, is_in ? kQuarters : 1);   // the four warps that produce a group
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWarpMma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemAddr(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  TcFenceBefore();
  __syncthreads();
  TcFenceAfter();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // same value in every lane, provably
  if (tracing && tid == 0) trace[7 * 16] = clock64();

  const int rows_valid = S * T;
  const size_t hist_unit = static_cast<size_t>(p.n_groups) * P * PAN * S * 8 * (k - 1);   // elements per unit dilation

  if (warp < kEpiWarps) {
    // =========================== epilogue warps ===========================
    const int q4 = warp & 3, whalf = warp >> 2, rtid = tid & 127;   // whalf is 0 when there are only 4 epilogue warps
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
    // Everything up to here (barriers, TMEM, bias, and in the other warps the weight ring and the
    // history loads) touches nothing the preceding kernel -- the upsampler that writes u -- produces.
    if (p.pdl_mode == 0) PdlWait();   // pdl_mode 1: u is already complete, see MrfStageParams
    PdlLaunchDependents();
    // ---- prologue: u -> TMEM (fp32 residual stream) and lrelu(u) -> X new rows ----
    for (int m = 0; m < MT; ++m) {
      const int r = m * 128 + rtid;
      const int t = r / S, s = r - t * S;
      const int b = group * S + s;
      const bool exists = r < rows_valid;
      const bool valid = exists && b < p.B;
      const float* urow = p.u + (static_cast<size_t>(b) * p.u_slots * T + (frame % p.u_slots) * T + t) * C;
      const uint32_t srow = x_base + static_cast<uint32_t>(HX * S + r) * 16;
      const float* frow = p.film ? p.film + static_cast<size_t>(b) * 2 * C : nullptr;
#pragma unroll 1
      for (int g = 0; g < G; ++g) {
        if (kEpiWarps == 8 && ((m * G + g) & 1) != whalf) continue;   // the partner warp's item
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
        TmemSt16(t_lane + m * DW + 16 * g, raw);
        if (kSplit) {   // the x_hi*W_lo half of the residual tile starts every c2 at zero
#pragma unroll
          for (int e = 0; e < 16; ++e) raw[e] = 0u;
          TmemSt16(t_lane + m * DW + C + 16 * g, raw);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = v[e] > 0.0f ? v[e] : 0.1f * v[e];
        uint4 h0, l0, h1, l1;
        Pack8<kSplit>(v, &h0, &l0);
        Pack8<kSplit>(v + 8, &h1, &l1);
        if (exists) {   // rows past S*T do not exist in the panels (the next panel starts there)
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
        if (lane == 0) MbarArrive(bar_in + 8 * (m * G + g));
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
      const float* bias = bias_s + i * C;
      if (i >= 1 && !last) MbarWait(bar_free + 8 * ((i + 1) & 1), ((i - 1) >> 1) & 1);
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
        if (tid == 0 && m == MT - 1) B200_TR(i, 1);
        if (m == MT - 1 && tid == 0) atomicAdd(const_cast<uint32_t*>(acc_cnt), 1u);   // conv i's MMAs have all retired
        const uint32_t tcol = t_lane + (is_c1 ? (MT + m) * DW : m * DW);
        const uint32_t srow = d_base + static_cast<uint32_t>(d_hmax * S + r) * 16;
        float* orow = br.out + (static_cast<size_t>(b) * br.out_slots * T + (frame % br.out_slots) * T + t) * C;
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
          if (kEpiWarps == 8 && ((m * G + g) & 1) != whalf) continue;   // the partner warp's item
          uint32_t raw[16];
          TmemLd16(tcol + 16 * g, raw);
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(raw[e]);
          if (kSplit) {   // + the x_hi*W_lo half
            TmemLd16(tcol + C + 16 * g, raw);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] += __uint_as_float(raw[e]);
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] += bias[16 * g + e];
          if (!is_c1) {
            if (!last) {
#pragma unroll
              for (int e = 0; e < 16; ++e) raw[e] = __float_as_uint(v[e]);
              TmemSt16(tcol + 16 * g, raw);
              if (kSplit) {
#pragma unroll
                for (int e = 0; e < 16; ++e) raw[e] = 0u;
                TmemSt16(tcol + C + 16 * g, raw);
              }
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
            if (lane == 0) MbarArrive(bar_in + 8 * (m * G + g));
          }
        }
      }
      if (tid == 0) B200_TR(i, 3);
      if (!last) {
        __syncwarp();
        if (lane == 0) SmemAddRelease(in_cnt);
      }
    }
  } else if (warp == kWarpMma) {
    // =========================== MMA issuer ===========================
    {
      // The issue loop is one warp's dependent instruction stream, so it is kept lean: descriptors are a constant
      // high word plus a running low word (start address in 16-byte units), bumped per tap / per K step, and
      // the weight ring position is carried in (stage, within, phase) counters -- no division or multiply per step.
      const uint32_t idesc = MakeIdesc(C), idesc_cat = MakeIdesc(2 * C);
      constexpr uint32_t kWLbo = kSplit ? 2 * C * 16 : C * 16;          // weight K-panel stride
      const uint64_t w_tmpl = MakeDesc(0, kWLbo, 128);
      const uint32_t w_tmpl_lo = static_cast<uint32_t>(w_tmpl), w_hi32 = static_cast<uint32_t>(w_tmpl >> 32);
      uint32_t stage = 0, within = 0, wphase = 0;
      uint32_t w_lo = w_tmpl_lo + ((w_base >> 4) & 0x3FFFu);
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        const int buf = i & 1, dil = ConvDil(i);
        const int hmax = buf ? HY : HX;
        const uint32_t bbase = buf ? y_base : x_base, pstride = buf ? y_pstride : x_pstride, plane = buf ? y_plane : x_plane;
        const uint64_t a_tmpl = MakeDesc(0, pstride, 128);
        const uint32_t a_tmpl_lo = static_cast<uint32_t>(a_tmpl), a_hi32 = static_cast<uint32_t>(a_tmpl >> 32);
        const uint32_t tap_step = static_cast<uint32_t>(dil * S);       // rows between taps, == 16-byte units
        const uint32_t plane16 = plane >> 4, group_step = (2 * pstride) >> 4;
        MbarWait(bar_hist + 8 * buf, (i >> 1) & 1);
        if (lane == 0) B200_TR(i, 4);
#pragma unroll 1
        for (int m = 0; m < MT; ++m) {
          const uint32_t dcol = tmem_base + (buf == 0 ? (MT + m) * DW : m * DW);
          uint32_t acc = buf == 0 ? 0u : 1u;
          // tap 0 of channel group 0: rows (hmax - (k-1) dil) S + 128 m of panel 0
          uint32_t a_group = a_tmpl_lo + ((bbase >> 4) & 0x3FFFu) + static_cast<uint32_t>((hmax - (k - 1) * dil) * S + 128 * m);
#pragma unroll 1
          for (int g = 0; g < G; ++g) {
            MbarWait(bar_in + 8 * (m * G + g), i & 1);
            TcFenceAfter();
            if (m == 0 && g == 0) if (lane == 0) B200_TR(i, 5);
            if (m == 0 && g == G - 1) if (lane == 0) B200_TR(i, 6);
            uint32_t a_lo = a_group;
#pragma unroll 1
            for (int j = 0; j < k; ++j) {
              if (within == 0) {
                MbarWait(bar_w_full + 8 * stage, wphase);
                TcFenceAfter();
              }
              if (kSplit) {
                // weight rows of the K step: [W_hi ; W_lo], 2C rows per 8-element K panel
                MmaW2(dcol, a_lo, a_hi32, w_lo, w_hi32, idesc_cat, acc);          // x_hi*W_hi -> cols [0,C), x_hi*W_lo -> cols [C,2C)
                MmaW2(dcol, a_lo + plane16, a_hi32, w_lo, w_hi32, idesc, 1u);     // x_lo*W_hi -> cols [0,C) (first C rows of the same tile)
              } else {
                MmaW2(dcol, a_lo, a_hi32, w_lo, w_hi32, idesc, acc);
              }
              acc = 1u;
              a_lo += tap_step;
              w_lo += kKstepBytes >> 4;
              ++within;
              if (within == NK || (g == G - 1 && j == k - 1)) {
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
          if (m == 0) if (lane == 0) B200_TR(i, 2);
          if (m == MT - 1) if (lane == 0) B200_TR(i, 7);
        }
      }
    }
    __syncwarp();
  } else if (warp == kWarpW) {
    // =========================== weight producer ===========================
    if (ElectOneSync()) {
      const int ksteps = k * G;
      const int chunks = (ksteps + NK - 1) / NK;
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(br.w);
      uint32_t cc = 0;
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        const uint8_t* wconv = wsrc + static_cast<size_t>(i) * ksteps * kKstepBytes;
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
        // conv i block: [group][plane][panel][H*S rows][8]
        return br.hist + hist_unit * DilPrefix(i) + static_cast<size_t>(group) * P * PAN * H * S * 8;
      };
      auto load_hist = [&](int i) {
        const int buf = i & 1, H = (k - 1) * ConvDil(i);
        const int hmax = buf ? HY : HX;
        const uint32_t bbase = buf ? y_base : x_base, pstride = buf ? y_pstride : x_pstride, plane = buf ? y_plane : x_plane;
        const uint32_t bytes = static_cast<uint32_t>(H) * S * 16;
        const uint16_t* src = hist_ptr(i, H);
        MbarExpectTx(bar_hist + 8 * buf, bytes * P * PAN);
        for (int pl = 0; pl < P; ++pl)
          for (int pn = 0; pn < PAN; ++pn)
            TmaBulkLoad(bbase + pl * plane + pn * pstride + static_cast<uint32_t>((hmax - H) * S) * 16,
                        src + static_cast<size_t>(pl * PAN + pn) * H * S * 8, bytes, bar_hist + 8 * buf);
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
        SpinUntil(in_cnt, static_cast<uint32_t>(kEpiWarps * (i + 1)), 100000u + 433u);
        __threadfence_block();
        FenceProxyAsync();
        {
          const uint32_t bytes = static_cast<uint32_t>(H) * S * 16;
          uint16_t* dst = const_cast<uint16_t*>(hist_ptr(i, H));
          for (int pl = 0; pl < P; ++pl)
            for (int pn = 0; pn < PAN; ++pn)
              TmaBulkStore(dst + static_cast<size_t>(pl * PAN + pn) * H * S * 8,
                           bbase + pl * plane + pn * pstride + static_cast<uint32_t>((hmax + T - H) * S) * 16, bytes);
          BulkCommit();
          BulkWaitRead0();
        }
        B200_TR(i, 8);
        // conv i's MMAs done reading the buffer
        SpinUntil(acc_cnt, static_cast<uint32_t>(i + 1), 100000u + 448u);
        if (i + 2 < 6) load_hist(i + 2);
        MbarArrive(bar_free + 8 * buf);
        B200_TR(i, 9);
      }
      BulkWait0();
      B200_TR(7, 2);
    }
    __syncwarp();
  }

  if (p.pdl_mode == 1) PdlWait();   // the stage is complete only when the first launch of the pair is, too
  TcFenceBefore();
  __syncthreads();
  if (tracing && tid == 0) {
    const long long t0 = trace[7 * 16];
    printf("[mrf trace] C=%d k=%d S=%d MT=%d  prologue_done=%lld hist_drained=%lld end=%lld (cycles after init)\n", C, k, S, MT,
           trace[7 * 16 + 1] - t0, trace[7 * 16 + 2] - t0, clock64() - t0);
    for (int i = 0; i < 6; ++i)
      printf("[mrf trace]  conv %d: hist_ready %lld in_g0 %lld in_gLast %lld mma_issued_m0 %lld mma_issued %lld | acc_m0 %lld acc_mLast %lld epi_done %lld | tail_stored %lld hist_next %lld\n",
             i, trace[i * 16 + 4] - t0, trace[i * 16 + 5] - t0, trace[i * 16 + 6] - t0, trace[i * 16 + 2] - t0, trace[i * 16 + 7] - t0,
             trace[i * 16 + 0] - t0, trace[i * 16 + 1] - t0, trace[i * 16 + 3] - t0, trace[i * 16 + 8] - t0, trace[i * 16 + 9] - t0);
  }
  if (p.trace < 0 && tid == 0) {   // developer aid: residency window of every CTA (BEATRICE_B200_MRF_TRACE=-1)
    unsigned long long t_end_ns;
    uint32_t smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end_ns));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    printf("[%s cta] C %d bx %d by %d sm %u start_ns %llu end_ns %llu\n", "mrf", C, blockIdx.x, blockIdx.y, smid, t_start_ns, t_end_ns);
  }
  if (warp == kWarpMma) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

__global__ void mrf_zero_stream_kernel(const MrfHistBlock* __restrict__ blocks, int b) {
  const MrfHistBlock hb = blocks[blockIdx.x];
  const int group = b / hb.S, s = b - group * hb.S;
  // [group][plane*panel][H rows][S][8]
  uint4* base = reinterpret_cast<uint4*>(hb.base) + static_cast<size_t>(group) * hb.planes_panels * hb.H * hb.S;
  const int n = hb.planes_panels * hb.H;
  for (int i = threadIdx.x; i < n; i += blockDim.x) base[static_cast<size_t>(i) * hb.S + s] = make_uint4(0, 0, 0, 0);
}

template <int C, bool kSplit>
void LaunchMrfT(const MrfStageParams& p, size_t smem, cudaStream_t s) {
  static bool attr_set[64] = {};
  int dev = 0;
  B200_CHECK(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    B200_CHECK(cudaFuncSetAttribute(mrf_branch_kernel<C, kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev & 63] = true;
  }
  LaunchPdl(mrf_branch_kernel<C, kSplit>, dim3(p.n_groups, p.n_branches, 1), dim3(ThreadsFor(C), 1, 1), smem, s, 1, p);
}


}  // namespace

size_t MrfSmemBytes(int C, int T, int S, bool split, int kmax) {
  const int MT = (S * T + 127) / 128;
  const int P = split ? 2 : 1, PAN = C / 8;
  const int k = kmax;   // the launch is sized for its largest branch
  const size_t rows8 = (static_cast<size_t>(S) * T + 7) & ~static_cast<size_t>(7);
  const size_t RX = static_cast<size_t>((k - 1) * 5) * S + rows8, RY = static_cast<size_t>(k - 1) * S + rows8;
  (void)MT;
  size_t off = (8 * 40 + 16 + 6 * C * 4 + 127) / 128 * 128;
  off += P * PAN * RX * 16;
  off += P * PAN * RY * 16;
  off = (off + 127) / 128 * 128;
  off += static_cast<size_t>(NstFor(C)) * NkFor(C) * P * C * 32;   // also absorbs the last tile's MMA over-read (< 2 KB)
  off += 1024;   // developer trace area
  return off;
}

bool MrfFusedSupported(int C, int T, int S, bool split, int kmax) {
  if (C != 16 && C != 32 && C != 64 && C != 128) return false;
  const int MT = (S * T + 127) / 128;
  if (2 * MT * (split ? 2 * C : C) > 512) return false;
  if (2 * NstFor(C) + 4 + MT * (C / 16) + MT > 40) return false;
  return MrfSmemBytes(C, T, S, split, kmax) <= 227 * 1024;
}

size_t MrfHistElems(int C, int k, int S, int n_groups, bool split) {
  return static_cast<size_t>(n_groups) * (split ? 2 : 1) * (C / 8) * S * 8 * (k - 1) * 12;
}

size_t PackMrfWeights(const float* const w[6], int k, int C, bool split, bool concat_rows, uint16_t* out) {
  // per conv, K step ks = g*k + j, element = W[j][16g + 8p + e][n]:
  //   planar (bf16 mode, and split mode at C = 128 / the cluster kernel): [plane][2 panels][C rows (n)][8]
  //   concatenated (split mode at C <= 64 / the single-CTA kernel):       [2 panels][2C rows: hi n, then lo n][8]
  const int G = C / 16, P = split ? 2 : 1;
  const bool concat = split && concat_rows;
  const size_t kstep = static_cast<size_t>(P) * 2 * C * 8;
  const size_t total = 6 * static_cast<size_t>(k) * G * kstep;
  if (!out) return total;
  for (int i = 0; i < 6; ++i)
    for (int g = 0; g < G; ++g)
      for (int j = 0; j < k; ++j) {
        uint16_t* blk = out + (static_cast<size_t>(i) * k * G + static_cast<size_t>(g) * k + j) * kstep;
        for (int pp = 0; pp < 2; ++pp)
          for (int n = 0; n < C; ++n)
            for (int e = 0; e < 8; ++e) {
              const int ci = 16 * g + 8 * pp + e;
              const float val = w[i][(static_cast<size_t>(j) * C + ci) * C + n];
              const uint16_t h = Bf16Rn(val);
              if (concat) {
                blk[(static_cast<size_t>(pp) * 2 * C + n) * 8 + e] = h;
                blk[(static_cast<size_t>(pp) * 2 * C + C + n) * 8 + e] = Bf16Rn(val - Bf16ToF(h));
              } else {
                const size_t o = (static_cast<size_t>(pp) * C + n) * 8 + e;
                blk[o] = h;
                if (split) blk[static_cast<size_t>(2) * C * 8 + o] = Bf16Rn(val - Bf16ToF(h));
              }
            }
      }
  return total;
}

void LaunchMrfStage(const MrfStageParams& p, int C, bool split, cudaStream_t s) {
  int kmax = 3;
  for (int y = 0; y < p.n_branches; ++y) kmax = std::max(kmax, p.br[p.y2br[y]].k);
  const size_t smem = MrfSmemBytes(C, p.T, p.S, split, kmax);
#define B200_MRF_CASE(CC)                                  \
  case CC:                                                 \
    if (split) LaunchMrfT<CC, true>(p, smem, s);           \
    else LaunchMrfT<CC, false>(p, smem, s);                \
    break
  switch (C) {
    B200_MRF_CASE(16);
    B200_MRF_CASE(32);
    B200_MRF_CASE(64);
    B200_MRF_CASE(128);
    default:
      Fail(-105, "fused MRF kernel has no form for this width", __FILE__, __LINE__);
  }
#undef B200_MRF_CASE
  B200_CHECK(cudaGetLastError());
}

void LaunchMrfZeroStream(const MrfHistBlock* d_blocks, int n_blocks, int b, cudaStream_t s) {
  if (n_blocks <= 0) return;
  mrf_zero_stream_kernel<<<n_blocks, 128, 0, s>>>(d_blocks, b);
  B200_CHECK(cudaGetLastError());
}

void SetSpinDebugMrf(unsigned long long* dev_ptr) { SetSpinDebugPtr(dev_ptr); }

}  // namespace b200
