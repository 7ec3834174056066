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

// kUps: the stage's upsampler runs in the prologue (MrfUpsDesc); a template parameter so that the plain form carries none of it
template <int C, bool kSplit, bool kUps>
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
  // branches this CTA runs one after the other (MrfStageParams::yseq); the shared-memory layout is that of the
  // largest of them, a smaller branch keeps its (shorter) histories right in front of the hop's rows
  const int nb = p.ylen[blockIdx.y];
  int kmax = 3;
  for (int bi = 0; bi < nb; ++bi) kmax = max(kmax, p.br[p.yseq[blockIdx.y][bi]].k);
  const int T = p.T, S = p.S, MT = p.MT;
  const int group = blockIdx.x;
  const int HX = (kmax - 1) * 5, HY = kmax - 1;      // history rows (time steps) in front of X / Y
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
                 bar_free = bar_hist + 16, bar_in = bar_free + 16, bar_acc = bar_in + 8 * MT * G,
                 bar_ups_in = bar_acc + 8 * MT, bar_ups_acc = bar_ups_in + 8;   // fused upsampler: input panels written / MMAs retired
  const int n_bars = 2 * kNst + 4 + MT * G + MT + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * 40);
  // hand-off counters for the history mover (monotonic, so a late reader can never alias a phase)
  volatile uint32_t* in_cnt = reinterpret_cast<volatile uint32_t*>(smem + 8 * 40 + 4);   // += 1 per epilogue warp per conv input
  volatile uint32_t* acc_cnt = reinterpret_cast<volatile uint32_t*>(smem + 8 * 40 + 8);  // += 1 per conv whose MMAs retired
  float* bias_s = reinterpret_cast<float*>(smem + 8 * 40 + 16);          // [branch of this CTA][6][C], then the upsampler's [C]
  float* bias_u = bias_s + p.nb_max * 6 * C;
  float* film_s = bias_u + C;                                            // [S][2C] gamma | beta of the group's streams (fused upsampler)
  const uint32_t x_off = (8 * 40 + 16 + ((p.nb_max * 6 + 1) * C + (kUps ? p.S * 2 * C : 0)) * 4 + 127) / 128 * 128;
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

  for (int bi = 0; bi < nb; ++bi) {
    const float* bsrc = p.br[p.yseq[blockIdx.y][bi]].bias;
    for (int i = tid; i < 6 * C; i += kThreads) bias_s[bi * 6 * C + i] = __ldg(bsrc + i);
  }
  if (kUps) {
    for (int i = tid; i < C; i += kThreads) bias_u[i] = __ldg(p.ups.bias + i);
    // FiLM parameters are written by the setters, outside and in front of the hop graph: readable before the dependency wait
    for (int i = tid; i < p.S * 2 * C; i += kThreads) {
      const int b = blockIdx.x * p.S + i / (2 * C);
      film_s[i] = (p.film && b < p.B) ? __ldg(p.film + static_cast<size_t>(b) * 2 * C + (i % (2 * C))) : 0.0f;
    }
  }
  if (tid == 0) {
    *in_cnt = 0;
    *acc_cnt = 0;
    for (int i = 0; i < n_bars; ++i) {
      const uint32_t b = bar0 + 8 * i;
      const bool is_in = b >= bar_in && b < bar_acc;
      MbarInit(b, is_in ? kQuarters : (b == bar_ups_in ? kEpiWarps : 1));   // the four warps that produce a group / all epilogue warps
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
  // ---- fused upsampler (MrfUpsDesc) ----
  constexpr bool fuse_ups = kUps;
  constexpr int GU = (2 * C) / 16;                     // 16-channel groups of the upsampler's input
  const int r_up = fuse_ups ? p.ups.r : 1, Tin = T / r_up;
  const int rows_in = S * (Tin + 1);                   // one row of history per stream in front (time-major, like X)
  const uint32_t kstep_u = static_cast<uint32_t>(P * r_up * C * 32);     // bytes of an upsampler K step: [plane][2 panels][r C rows][8]
  const uint32_t nku = kChunkBytes / kstep_u;          // K steps per weight-ring chunk (host checks >= 1)
  // its input panels alias the new-row regions of the X slots (2C/8 panels == P * C/8 slots in split mode); lo behind hi
  const uint32_t a_region = static_cast<uint32_t>(HX * S) * 16, a_lo_off = static_cast<uint32_t>(rows_in) * 16;

  if (warp < kEpiWarps) {
    // =========================== epilogue warps ===========================
    const int q4 = warp & 3, whalf = warp >> 2, rtid = tid & 127;   // whalf is 0 when there are only 4 epilogue warps
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
    // Everything up to here (barriers, TMEM, bias, and in the other warps the weight ring and the
    // history loads) touches nothing the preceding kernel -- the upsampler that writes u -- produces.
    if (p.pdl_mode == 0) PdlWait();   // pdl_mode 1: u is already complete, see MrfStageParams
    PdlLaunchDependents();
    if (tid == 0) B200_TR(7, 6);
#pragma unroll 1
    for (int bi = 0; bi < nb; ++bi) {
    const MrfBranchDesc& br = p.br[p.yseq[blockIdx.y][bi]];
    const bool tr = tracing && bi == 0;
    // bar_free[buf] completes three phases per branch (after convs 0/2/4 for X, 1/3/5 for Y).  A further branch
    // reuses X once the previous branch's conv 4 is done with it (its MMAs retired, its tail stored) ...
    if (bi >= 1) MbarWait(bar_free, (3 * bi - 1) & 1);
    if (fuse_ups) {
      constexpr int E = kEpiWarps * 32;
      if (bi >= 1) MbarWait(bar_free + 8, (3 * bi - 1) & 1);   // the staging below aliases Y's new rows
      // ---- gather: xin = lrelu(mean of the previous stage's branches), bf16 hi/lo panels, rows (i, s) time-major.
      //      One item = 4 channels (16 bytes per ring) of one row, the column chunk fastest across threads so that the lanes of a
      //      warp read whole rows (a lane per distinct cache line made the L1 the bottleneck: 8-14k cycles for this phase);
      //      all loads of a thread's kGU items are issued before the first use (volatile asm: under the register cap the
      //      compiler otherwise interleaves them with the arithmetic) ----
      constexpr int kGU = C >= 64 ? 5 : 2;                 // items in flight per thread
      constexpr int kCols = (2 * C) / 4;                   // 16-byte column chunks per row
      const int n_items = rows_in * kCols;
#pragma unroll 1
      for (int it0 = tid; it0 < n_items; it0 += E * kGU) {
        float4 xa[kGU], xb[kGU], xc[kGU];
        uint32_t dst_[kGU];
        bool live_[kGU];
#pragma unroll
        for (int u = 0; u < kGU; ++u) {
          const int it = it0 + u * E;
          const int rho = it / kCols, col = it - rho * kCols;
          const int ti = rho / S - 1, s = rho - (ti + 1) * S;     // ti = -1: the last row of the previous hop
          const int b = group * S + s;
          live_[u] = it < n_items;
          // panel col / 2 (8 channels), row rho, half col & 1
          dst_[u] = x_base + static_cast<uint32_t>(col >> 1) * x_pstride + a_region + static_cast<uint32_t>(rho) * 16 + static_cast<uint32_t>(col & 1) * 8;
          const bool ld = live_[u] && b < p.B;
          const int slot = (ti >= 0 ? frame : frame + p.ups.x_slots - 1) % p.ups.x_slots;
          const size_t off = ld ? (static_cast<size_t>(b) * p.ups.x_slots * Tin + static_cast<size_t>(slot) * Tin + (ti >= 0 ? ti : Tin - 1)) * (2 * C) + 4 * col : 0;
          xa[u] = xb[u] = xc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ld) {
            asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(xa[u].x), "=f"(xa[u].y), "=f"(xa[u].z), "=f"(xa[u].w) : "l"(p.ups.x[0] + off));
            asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(xb[u].x), "=f"(xb[u].y), "=f"(xb[u].z), "=f"(xb[u].w) : "l"(p.ups.x[1] + off));
            asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(xc[u].x), "=f"(xc[u].y), "=f"(xc[u].z), "=f"(xc[u].w) : "l"(p.ups.x[2] + off));
          }
        }
#pragma unroll
        for (int u = 0; u < kGU; ++u) {
          constexpr float kThird = 1.0f / 3.0f;
          float v[4];   // same order as the stand-alone upsampler's gather (b200_tc.cu)
          v[0] = ((xa[u].x + xb[u].x) + xc[u].x) * kThird;
          v[1] = ((xa[u].y + xb[u].y) + xc[u].y) * kThird;
          v[2] = ((xa[u].z + xb[u].z) + xc[u].z) * kThird;
          v[3] = ((xa[u].w + xb[u].w) + xc[u].w) * kThird;
          uint32_t h[2], l[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float f0 = v[2 * i] > 0.0f ? v[2 * i] : 0.1f * v[2 * i], f1 = v[2 * i + 1] > 0.0f ? v[2 * i + 1] : 0.1f * v[2 * i + 1];
            const __nv_bfloat16 a = __float2bfloat16_rn(f0), b2 = __float2bfloat16_rn(f1);
            h[i] = static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b2)) << 16);
            const __nv_bfloat16 ra = __float2bfloat16_rn(f0 - __bfloat162float(a)), rb = __float2bfloat16_rn(f1 - __bfloat162float(b2));
            l[i] = static_cast<uint32_t>(__bfloat16_as_ushort(ra)) | (static_cast<uint32_t>(__bfloat16_as_ushort(rb)) << 16);
          }
          if (live_[u]) {
            asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(dst_[u]), "r"(h[0]), "r"(h[1]) : "memory");
            if (kSplit) asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(dst_[u] + a_lo_off), "r"(l[0]), "r"(l[1]) : "memory");
          }
        }
      }
      FenceProxyAsync();
      TcFenceBefore();   // (a further branch: this warp's TMEM reads of the previous branch's result are done)
      __syncwarp();
      if (lane == 0) MbarArrive(bar_ups_in);
      if (tid == 0 && tr) B200_TR(7, 3);
      // ---- the upsampler's result: + bias, FiLM, fp32 rows re-laid out (t = r i + ph) into the staging ----
      MbarWait(bar_ups_acc, bi & 1);
      TcFenceAfter();
      if (tid == 0 && tr) B200_TR(7, 4);
      if (q4 * 32 < S * Tin) {   // this lane quarter holds result rows (warp-uniform)
        const int i_in = rtid / S, s = rtid - i_in * S;
        const int b = group * S + s;
        const bool live = rtid < S * Tin;
        (void)b;
        const float* frow = (p.film && live) ? film_s + s * 2 * C : nullptr;
#pragma unroll 1
        for (int item = (kEpiWarps == 8 ? whalf : 0); item < r_up * G; item += (kEpiWarps == 8 ? 2 : 1)) {
          const int ph = item / G, g = item - ph * G;
          uint32_t raw[16];
          TmemLd16(t_lane + ph * C + 16 * g, raw);
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(raw[e]) + bias_u[16 * g + e];
          if (frow) {   // FiLM of this vocoder stage (per stream): u * (1 + gamma) + beta
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 ga = *reinterpret_cast<const float4*>(frow + 16 * g + 4 * q);
              const float4 be = *reinterpret_cast<const float4*>(frow + C + 16 * g + 4 * q);
              v[4 * q] = v[4 * q] * (1.0f + ga.x) + be.x;
              v[4 * q + 1] = v[4 * q + 1] * (1.0f + ga.y) + be.y;
              v[4 * q + 2] = v[4 * q + 2] * (1.0f + ga.z) + be.z;
              v[4 * q + 3] = v[4 * q + 3] * (1.0f + ga.w) + be.w;
            }
          }
          if (live) {
            const uint32_t R = static_cast<uint32_t>((r_up * i_in + ph) * S + s);
#pragma unroll
            for (int q = 0; q < 4; ++q) {   // 16-byte chunk j = 4 g + q of row R lives in Y slot j (C/4 chunks == P * C/8 slots)
              const int j = 4 * g + q;
              const uint32_t a = y_base + static_cast<uint32_t>(j) * y_pstride + static_cast<uint32_t>(HY * S) * 16 + R * 16;
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v[4 * q]), "f"(v[4 * q + 1]), "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
            }
          }
        }
      }
      TcFenceBefore();
      asm volatile("bar.sync 1, %0;" ::"n"(E) : "memory");   // staging complete, every warp done reading the result tile
      TcFenceAfter();
      if (tid == 0 && tr) B200_TR(7, 5);
    }
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
          if (fuse_ups) {
            if (valid) {
              const uint32_t a = y_base + static_cast<uint32_t>(4 * g + q) * y_pstride + static_cast<uint32_t>(HY * S + r) * 16;
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(a) : "memory");
            }
          } else if (valid) {
            f = __ldg(reinterpret_cast<const float4*>(urow + 16 * g + 4 * q));
          }
          if (!fuse_ups && valid && frow) {   // FiLM of this vocoder stage (per stream): u * (1 + gamma) + beta
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
    if (tid == 0 && tr) B200_TR(7, 1);
    // ---- the six convs ----
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
      const bool is_c1 = (i & 1) == 0;
      const bool last = i == 5;
      // destination of lrelu(.): the OTHER buffer (c1 -> Y, c2 -> X)
      const uint32_t d_base = is_c1 ? y_base : x_base;
      const uint32_t d_pstride = is_c1 ? y_pstride : x_pstride, d_plane = is_c1 ? y_plane : x_plane;
      const int d_hmax = is_c1 ? HY : HX;
      const float* bias = bias_s + (bi * 6 + i) * C;
      // the buffer lrelu(.) goes to is free: the conv that read it last is done with it (phase 3 bi + that conv's
      // index / 2 of bar_free[buf]); ... and a further branch's conv 0 reuses Y after the previous branch's conv 5
      if (i >= 1 && !last) MbarWait(bar_free + 8 * ((i + 1) & 1), (3 * bi + ((i - 1) >> 1)) & 1);
      if (i == 0 && bi >= 1) MbarWait(bar_free + 8, (3 * bi - 1) & 1);
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int r = m * 128 + rtid;
        const int t = r / S, s = r - t * S;
        const int b = group * S + s;
        const bool exists = r < rows_valid;
        const bool valid = exists && b < p.B;
        MbarWait(bar_acc + 8 * m, i & 1);
        TcFenceAfter();
        if (tid == 0 && m == 0 && tr) B200_TR(i, 0);
        if (tid == 0 && m == MT - 1 && tr) B200_TR(i, 1);
        if (m == MT - 1 && tid == 0) atomicAdd(const_cast<uint32_t*>(acc_cnt), 1u);   // conv i's MMAs have all retired
        const uint32_t tcol = t_lane + (is_c1 ? (MT + m) * DW : m * DW);
        const uint32_t srow = d_base + static_cast<uint32_t>(d_hmax * S + r) * 16;
        float* orow = br.out + (static_cast<size_t>(b) * br.out_slots * T + (frame % br.out_slots) * T + t) * C;
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
          if (kEpiWarps == 8 && ((m * G + g) & 1) != whalf) continue;   // the partner warp's item
          uint32_t raw[16];
          float v[16];
          if (kSplit) {   // accumulator columns [0,C) + the x_hi*W_lo half in [C,2C): both loads in flight, one wait
            uint32_t raw2[16];
            TmemLd16x2(tcol + 16 * g, raw, tcol + C + 16 * g, raw2);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(raw[e]) + __uint_as_float(raw2[e]);
          } else {
            TmemLd16(tcol + 16 * g, raw);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(raw[e]);
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
      if (tid == 0 && tr) B200_TR(i, 3);
      if (!last) {
        __syncwarp();
        if (lane == 0) SmemAddRelease(in_cnt);
      }
    }
    }   // branches of this CTA
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
      for (int bi = 0; bi < nb; ++bi) {
      const int k = p.br[p.yseq[blockIdx.y][bi]].k;
      const bool tr = tracing && bi == 0;
      if (fuse_ups) {
        // D[(i, s)][(ph, co)] in accumulator columns [0, r C): tap 1 reads the rows from S on (xin[i]), tap 0 from 0 (xin[i-1])
        const uint32_t idesc_u = MakeIdesc(r_up * C);
        const uint64_t wu_tmpl = MakeDesc(0, static_cast<uint32_t>(r_up * C) * 16, 128);
        const uint32_t wu_tmpl_lo = static_cast<uint32_t>(wu_tmpl), wu_hi32 = static_cast<uint32_t>(wu_tmpl >> 32);
        const uint64_t au_tmpl = MakeDesc(0, x_pstride, 128);
        const uint32_t au_tmpl_lo = static_cast<uint32_t>(au_tmpl), au_hi32 = static_cast<uint32_t>(au_tmpl >> 32);
        const uint32_t w_plane16 = (static_cast<uint32_t>(r_up * C) * 32) >> 4;    // lo plane of a K step
        MbarWait(bar_ups_in, bi & 1);
        TcFenceAfter();
        uint32_t acc = 0u;
#pragma unroll 1
        for (int tap = 0; tap < 2; ++tap) {
          uint32_t a_lo = au_tmpl_lo + (((x_base + a_region) >> 4) & 0x3FFFu) + static_cast<uint32_t>(tap * S);
#pragma unroll 1
          for (int g = 0; g < GU; ++g) {
            if (within == 0) {
              MbarWait(bar_w_full + 8 * stage, wphase);
              TcFenceAfter();
              w_lo = wu_tmpl_lo + (((w_base + stage * kChunkBytes) >> 4) & 0x3FFFu);
            }
            MmaW2(tmem_base, a_lo, au_hi32, w_lo, wu_hi32, idesc_u, acc);                       // x_hi * W_hi
            if (kSplit) {
              MmaW2(tmem_base, a_lo, au_hi32, w_lo + w_plane16, wu_hi32, idesc_u, 1u);          // x_hi * W_lo
              MmaW2(tmem_base, a_lo + (a_lo_off >> 4), au_hi32, w_lo, wu_hi32, idesc_u, 1u);    // x_lo * W_hi
            }
            acc = 1u;
            a_lo += (2 * x_pstride) >> 4;
            w_lo += kstep_u >> 4;
            ++within;
            if (within == nku || (tap == 1 && g == GU - 1)) {
              MmaCommitW(bar_w_empty + 8 * stage);
              within = 0;
              ++stage;
              if (stage == kNst) {
                stage = 0;
                wphase ^= 1u;
              }
            }
          }
        }
        MmaCommitW(bar_ups_acc);
        w_lo = w_tmpl_lo + (((w_base + stage * kChunkBytes) >> 4) & 0x3FFFu);   // back to the conv weights' descriptor form
      }
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        const int buf = i & 1, dil = ConvDil(i);
        const int hmax = buf ? HY : HX;
        const uint32_t bbase = buf ? y_base : x_base, pstride = buf ? y_pstride : x_pstride, plane = buf ? y_plane : x_plane;
        const uint64_t a_tmpl = MakeDesc(0, pstride, 128);
        const uint32_t a_tmpl_lo = static_cast<uint32_t>(a_tmpl), a_hi32 = static_cast<uint32_t>(a_tmpl >> 32);
        const uint32_t tap_step = static_cast<uint32_t>(dil * S);       // rows between taps, == 16-byte units
        const uint32_t plane16 = plane >> 4, group_step = (2 * pstride) >> 4;
        MbarWait(bar_hist + 8 * buf, (3 * bi + (i >> 1)) & 1);   // three history loads per buffer per branch
        if (lane == 0 && tr) B200_TR(i, 4);
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
            if (m == 0 && g == 0) if (lane == 0 && tr) B200_TR(i, 5);
            if (m == 0 && g == G - 1) if (lane == 0 && tr) B200_TR(i, 6);
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
          if (m == 0) if (lane == 0 && tr) B200_TR(i, 2);
          if (m == MT - 1) if (lane == 0 && tr) B200_TR(i, 7);
        }
      }
      }   // branches of this CTA
    }
    __syncwarp();
  } else if (warp == kWarpW) {
    // =========================== weight producer ===========================
    if (ElectOneSync()) {
      uint32_t cc = 0;
#pragma unroll 1
      for (int bi = 0; bi < nb; ++bi) {
      const MrfBranchDesc& br = p.br[p.yseq[blockIdx.y][bi]];
      const int ksteps = br.k * G;
      const int chunks = (ksteps + NK - 1) / NK;
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(br.w);
      if (fuse_ups) {
        const int n_ku = 2 * GU;
        const uint8_t* wu = reinterpret_cast<const uint8_t*>(p.ups.w);
#pragma unroll 1
        for (int c0 = 0; c0 < n_ku; c0 += static_cast<int>(nku)) {
          const uint32_t stage = cc % kNst, round = cc / kNst;
          if (round > 0) MbarWait(bar_w_empty + 8 * stage, (round - 1) & 1);
          const uint32_t bytes = static_cast<uint32_t>(min(static_cast<int>(nku), n_ku - c0)) * kstep_u;
          MbarExpectTx(bar_w_full + 8 * stage, bytes);
          TmaBulkLoadKeep(w_base + stage * kChunkBytes, wu + static_cast<size_t>(c0) * kstep_u, bytes, bar_w_full + 8 * stage);
          ++cc;
        }
      }
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
      }   // branches of this CTA
    }
    __syncwarp();
  } else if (warp == kWarpH) {
    // =========================== history mover ===========================
    if (ElectOneSync()) {
      auto hist_ptr = [&](const MrfBranchDesc& br, int i, int H) {
        // conv i block: [group][plane][panel][H*S rows][8]; blocks in conv order, (k-1)*dil_i rows (time steps) each
        const size_t hist_unit = static_cast<size_t>(p.n_groups) * P * PAN * S * 8 * (br.k - 1);   // elements per unit dilation
        return br.hist + hist_unit * DilPrefix(i) + static_cast<size_t>(group) * P * PAN * H * S * 8;
      };
      auto load_hist = [&](const MrfBranchDesc& br, int i) {
        const int buf = i & 1, H = (br.k - 1) * ConvDil(i);
        const int hmax = buf ? HY : HX;
        const uint32_t bbase = buf ? y_base : x_base, pstride = buf ? y_pstride : x_pstride, plane = buf ? y_plane : x_plane;
        const uint32_t bytes = static_cast<uint32_t>(H) * S * 16;
        const uint16_t* src = hist_ptr(br, i, H);
        MbarExpectTx(bar_hist + 8 * buf, bytes * P * PAN);
        for (int pl = 0; pl < P; ++pl)
          for (int pn = 0; pn < PAN; ++pn)
            TmaBulkLoad(bbase + pl * plane + pn * pstride + static_cast<uint32_t>((hmax - H) * S) * 16,
                        src + static_cast<size_t>(pl * PAN + pn) * H * S * 8, bytes, bar_hist + 8 * buf);
      };
      load_hist(p.br[p.yseq[blockIdx.y][0]], 0);
      load_hist(p.br[p.yseq[blockIdx.y][0]], 1);
#pragma unroll 1
      for (int bi = 0; bi < nb; ++bi) {
      const MrfBranchDesc& br = p.br[p.yseq[blockIdx.y][bi]];
      const bool tr = tracing && bi == 0;
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        const int buf = i & 1, H = (br.k - 1) * ConvDil(i);
        const int hmax = buf ? HY : HX;
        const uint32_t bbase = buf ? y_base : x_base, pstride = buf ? y_pstride : x_pstride, plane = buf ? y_plane : x_plane;
        // input of conv i complete: history landed + every new row written by the epilogue warps
        MbarWait(bar_hist + 8 * buf, (3 * bi + (i >> 1)) & 1);
        SpinUntil(in_cnt, static_cast<uint32_t>(kEpiWarps * (6 * bi + i + 1)), 100000u + 433u);
        __threadfence_block();
        FenceProxyAsync();
        {
          const uint32_t bytes = static_cast<uint32_t>(H) * S * 16;
          uint16_t* dst = const_cast<uint16_t*>(hist_ptr(br, i, H));
          for (int pl = 0; pl < P; ++pl)
            for (int pn = 0; pn < PAN; ++pn)
              TmaBulkStore(dst + static_cast<size_t>(pl * PAN + pn) * H * S * 8,
                           bbase + pl * plane + pn * pstride + static_cast<uint32_t>((hmax + T - H) * S) * 16, bytes);
          BulkCommit();
          BulkWaitRead0();
        }
        if (tr) B200_TR(i, 8);
        // conv i's MMAs done reading the buffer
        SpinUntil(acc_cnt, static_cast<uint32_t>(6 * bi + i + 1), 100000u + 448u);
        if (i + 2 < 6) load_hist(br, i + 2);
        else if (bi + 1 < nb) load_hist(p.br[p.yseq[blockIdx.y][bi + 1]], i - 4);   // the next branch's first two histories
        MbarArrive(bar_free + 8 * buf);
        if (tr) B200_TR(i, 9);
      }
      }   // branches of this CTA
      BulkWait0();
      if (tracing) B200_TR(7, 2);
    }
    __syncwarp();
  }

  if (p.pdl_mode == 1) PdlWait();   // the stage is complete only when the first launch of the pair is, too
  TcFenceBefore();
  __syncthreads();
  if (tracing && tid == 0) {
    const long long t0 = trace[7 * 16];
    printf("[mrf trace] C=%d k=%d S=%d MT=%d  prologue_done=%lld hist_drained=%lld end=%lld (cycles after init)\n", C, p.br[p.yseq[blockIdx.y][0]].k, S, MT,
           trace[7 * 16 + 1] - t0, trace[7 * 16 + 2] - t0, clock64() - t0);
    if (fuse_ups)
      printf("[mrf trace]  fused upsampler: dep_wait_done %lld gathered %lld mma_done %lld staged %lld\n", trace[7 * 16 + 6] - t0,
             trace[7 * 16 + 3] - t0, trace[7 * 16 + 4] - t0, trace[7 * 16 + 5] - t0);
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

template <int C, bool kSplit, bool kUps>
void LaunchMrfT(const MrfStageParams& p, size_t smem, cudaStream_t s) {
  static bool attr_set[64] = {};
  int dev = 0;
  B200_CHECK(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    B200_CHECK(cudaFuncSetAttribute(mrf_branch_kernel<C, kSplit, kUps>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev & 63] = true;
  }
  LaunchMaybePdl(!(p.late_launch && p.pdl_mode == 0), mrf_branch_kernel<C, kSplit, kUps>, dim3(p.n_groups, p.n_branches, 1),
                 dim3(ThreadsFor(C), 1, 1), smem, s, 1, p);
}


}  // namespace

size_t MrfSmemBytes(int C, int T, int S, bool split, int kmax, int nb, bool fused_ups) {
  const int MT = (S * T + 127) / 128;
  const int P = split ? 2 : 1, PAN = C / 8;
  const int k = kmax;   // the launch is sized for its largest branch
  const size_t rows8 = (static_cast<size_t>(S) * T + 7) & ~static_cast<size_t>(7);
  const size_t RX = static_cast<size_t>((k - 1) * 5) * S + rows8, RY = static_cast<size_t>(k - 1) * S + rows8;
  (void)MT;
  size_t off = (8 * 40 + 16 + ((static_cast<size_t>(nb) * 6 + 1) * C + (fused_ups ? static_cast<size_t>(S) * 2 * C : 0)) * 4 + 127) / 128 * 128;
  off += P * PAN * RX * 16;
  off += P * PAN * RY * 16;
  off = (off + 127) / 128 * 128;
  off += static_cast<size_t>(NstFor(C)) * NkFor(C) * P * C * 32;   // also absorbs the last tile's MMA over-read (< 2 KB)
  off += 1024;   // developer trace area
  return off;
}

void MrfOneBranchPerCta(MrfStageParams* p) {
  p->nb_max = 1;
  for (int y = 0; y < 3; ++y) {
    p->ylen[y] = y < p->n_branches ? 1 : 0;
    p->yseq[y][0] = p->y2br[y];
    p->yseq[y][1] = p->yseq[y][2] = 0;
  }
}

bool MrfFusedSupported(int C, int T, int S, bool split, int kmax, int nb, bool fused_ups) {
  if (C != 16 && C != 32 && C != 64 && C != 128) return false;
  const int MT = (S * T + 127) / 128;
  if (2 * MT * (split ? 2 * C : C) > 512) return false;
  if (2 * NstFor(C) + 4 + MT * (C / 16) + MT + 2 > 40) return false;
  return MrfSmemBytes(C, T, S, split, kmax, nb, fused_ups) <= 227 * 1024;
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

size_t PackMrfUpsWeights(const float* w, int C, int r, uint16_t* out) {
  const int Ci = 2 * C, N = r * C, GU = Ci / 16;
  const size_t kstep = static_cast<size_t>(2) * 2 * N * 8;   // [plane][2 panels][N][8]
  const size_t total = static_cast<size_t>(2) * GU * kstep;
  if (!out) return total;
  for (int tap = 0; tap < 2; ++tap)
    for (int g = 0; g < GU; ++g) {
      uint16_t* blk = out + (static_cast<size_t>(tap) * GU + g) * kstep;
      for (int pp = 0; pp < 2; ++pp)
        for (int n = 0; n < N; ++n)
          for (int e = 0; e < 8; ++e) {
            const int ci = 16 * g + 8 * pp + e;
            const float val = w[(static_cast<size_t>(tap) * Ci + ci) * N + n];
            const uint16_t h = Bf16Rn(val);
            const size_t o = (static_cast<size_t>(pp) * N + n) * 8 + e;
            blk[o] = h;
            blk[static_cast<size_t>(2) * N * 8 + o] = Bf16Rn(val - Bf16ToF(h));
          }
    }
  return total;
}

bool MrfUpsFusable(int C, int T, int S, int r, bool split) {
  if (!split || C > 64 || r < 1 || T % r != 0) return false;
  const int MT = (S * T + 127) / 128, Tin = T / r, N = r * C;
  if (N > 256 || N % 16 != 0 || N > 2 * MT * 2 * C) return false;            // one MMA wide, inside the accumulator columns
  if (S * Tin > 128) return false;                                             // one 128-row tile of result rows
  if (2 * S * (Tin + 1) > ((S * T + 7) & ~7)) return false;                    // hi + lo input panels inside an X slot's new rows
  if (static_cast<uint32_t>(2 * N * 32) > static_cast<uint32_t>(NkFor(C) * 2 * C * 32)) return false;   // a K step fits a ring chunk
  return true;
}

void LaunchMrfStage(const MrfStageParams& p, int C, bool split, cudaStream_t s) {
  int kmax = 3;
  for (int y = 0; y < p.n_branches; ++y) {
    if (p.ylen[y] < 1 || p.ylen[y] > 3 || p.ylen[y] > p.nb_max) Fail(-107, "fused MRF launch without a branch list (MrfOneBranchPerCta)", __FILE__, __LINE__);
    for (int bi = 0; bi < p.ylen[y]; ++bi) kmax = std::max(kmax, p.br[p.yseq[y][bi]].k);
  }
  const size_t smem = std::min<size_t>(std::max<size_t>(MrfSmemBytes(C, p.T, p.S, split, kmax, p.nb_max, p.ups.w != nullptr), static_cast<size_t>(p.smem_min)), 227 * 1024);
  const bool ups = p.ups.w != nullptr;
  if (ups && (!split || C > 64)) Fail(-108, "the fused MRF kernel computes the upsampler only in split precision at C <= 64", __FILE__, __LINE__);
#define B200_MRF_CASE(CC)                                  \
  case CC:                                                 \
    if (split && ups) LaunchMrfT<CC, true, true>(p, smem, s);   \
    else if (split) LaunchMrfT<CC, true, false>(p, smem, s);    \
    else LaunchMrfT<CC, false, false>(p, smem, s);         \
    break
  switch (C) {
    B200_MRF_CASE(16);
    B200_MRF_CASE(32);
    B200_MRF_CASE(64);
    case 128:
      if (split) LaunchMrfT<128, true, false>(p, smem, s);
      else LaunchMrfT<128, false, false>(p, smem, s);
      break;
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
