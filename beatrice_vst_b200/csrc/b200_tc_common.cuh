// Device-side tcgen05 / TMA / mbarrier helpers shared by the tensor-core kernels (sm_100a).
#ifndef BEATRICE_B200_TC_COMMON_CUH_
#define BEATRICE_B200_TC_COMMON_CUH_

#include <cuda_bf16.h>

#include <cstdint>
#include <cstring>

#include "b200_kernels.h"

namespace b200 {
namespace {

// One lane of a converged warp.  Unlike `lane == 0`, elect.sync tells the compiler that exactly one
// thread runs the guarded region, so tcgen05 / TMA instructions (which take warp-uniform operands)
// are issued directly instead of inside a compiler-generated per-thread election loop.
__device__ __forceinline__ bool ElectOneSync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P_el;\n"
      "elect.sync _|P_el, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P_el;\n"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t SmemAddr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void MbarInit(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void MbarExpectTx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void MbarArrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void MbarWait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void TmaBulkLoad(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// Weight tiles: the same bytes are re-read by every CTA and on every hop, while ~160 MB of stream
// state cycles through the 126 MB L2 in between -- ask L2 to keep them (evict-last policy).
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void TmaBulkLoadKeep(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "l"(kL2EvictLast)
      : "memory");
}
__device__ __forceinline__ void FenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void TcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void TcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave"):
// bits [0,14) start>>4, [16,30) LBO>>4 (next K panel), [32,46) SBO>>4 (next 8-row group),
// [46,48) version = 1 on sm_100, [61,64) layout type = 0.
__device__ __forceinline__ uint64_t MakeDesc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
// Instruction descriptor, kind::f16: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15, 16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28.
__device__ __forceinline__ uint32_t MakeIdesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(128 >> 4) << 24);
}
__device__ __forceinline__ void Mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// Warp-converged forms: the whole MMA warp walks the (uniform) issue loop so that descriptors and
// loop state stay in uniform registers; only the instruction itself is issued by one elected lane.
__device__ __forceinline__ void MmaW(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (ElectOneSync()) Mma(d_tmem, adesc, bdesc, idesc, accum);
}
// Same, descriptors passed as (low, high) words: issue loops keep the high words constant and bump the low
// word's start-address field.
__device__ __forceinline__ void MmaW2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                      uint32_t idesc, uint32_t accum) {
  if (ElectOneSync()) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
        "}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
        : "memory");
  }
}
__device__ __forceinline__ void MmaCommit(uint32_t bar);
__device__ __forceinline__ void MmaCommitW(uint32_t bar) {
  if (ElectOneSync()) MmaCommit(bar);
}
__device__ __forceinline__ void MmaCommit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void TmemLd16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// two 16-column loads in flight, one wait
__device__ __forceinline__ void TmemLd16x2(uint32_t taddr0, uint32_t* r0, uint32_t taddr1, uint32_t* r1) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]), "=r"(r0[8]),
        "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11]), "=r"(r0[12]), "=r"(r0[13]), "=r"(r0[14]), "=r"(r0[15])
      : "r"(taddr0));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]), "=r"(r1[8]),
        "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15])
      : "r"(taddr1));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void TmemSt16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void TmemStWait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void TmaBulkStore(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void BulkCommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void BulkWaitRead0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void BulkWait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// counter += 1 with release semantics at CTA scope (publishes this warp's earlier shared-memory writes
// after a __syncwarp) -- cheaper than a __threadfence_block, which also drains in-flight remote stores
__device__ __forceinline__ void SmemAddRelease(volatile uint32_t* cnt) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(SmemAddr(const_cast<uint32_t*>(cnt))) : "memory");
}
// A wait that exceeds ~2^28 polls (seconds) means the kernel is dead-locked: rather than hang the host forever the
// thread leaves a record of WHERE in host-visible memory (SetSpinDebug: a mapped pinned buffer the library prints when
// it latches the resulting launch failure) and traps.  One copy of the pointer per translation unit.
__device__ unsigned long long* g_spin_dbg = nullptr;
inline void SetSpinDebugPtr(unsigned long long* dev_ptr) { cudaMemcpyToSymbol(g_spin_dbg, &dev_ptr, sizeof(dev_ptr)); }
__device__ __noinline__ void SpinRecord(uint32_t site, uint32_t have, uint32_t want) {
  unsigned long long* d = g_spin_dbg;
  if (d != nullptr) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned long long slot = atomicAdd(d, 1ull);
    if (slot < 60) {
      unsigned long long* r = d + 8 + slot * 8;
      r[0] = site;
      r[1] = (static_cast<unsigned long long>(blockIdx.y) << 32) | blockIdx.x;
      r[2] = (static_cast<unsigned long long>(want) << 32) | have;
      r[3] = smid;
      r[4] = threadIdx.x;
    }
    __threadfence_system();
  }
}
// mbarrier waits with a watchdog (globaltimer): after ~1 s without completion the waiter leaves a record (site, CTA,
// the conv index / parity it waits for) and KEEPS waiting, so that every party of a dead-lock reports before the
// first polling thread traps.  try_wait suspends the thread in hardware, the extra compare is off the critical path.
__device__ __forceinline__ void MbarWaitDbg(uint32_t bar, uint32_t parity, uint32_t site, uint32_t info) {
  unsigned long long t0 = 0;
  bool reported = false;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0) t0 = now;
    if (!reported && now - t0 > 1000000000ull) {
      if ((threadIdx.x & 31) == 0) SpinRecord(site, parity, info);   // one record per warp
      reported = true;
    }
  }
}
__device__ __forceinline__ void MbarWaitClusterDbg(uint32_t bar, uint32_t parity, uint32_t site, uint32_t info) {
  unsigned long long t0 = 0;
  bool reported = false;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P2;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P2, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P2;\n"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0) t0 = now;
    if (!reported && now - t0 > 1000000000ull) {
      if ((threadIdx.x & 31) == 0) SpinRecord(site, parity, info);   // one record per warp
      reported = true;
    }
  }
}
__device__ __noinline__ void SpinTimeout(uint32_t site, uint32_t have, uint32_t want) {
  unsigned long long* d = g_spin_dbg;
  if (d != nullptr) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned long long slot = atomicAdd(d, 1ull);
    if (slot < 60) {
      unsigned long long* r = d + 8 + slot * 8;
      r[0] = site;
      r[1] = (static_cast<unsigned long long>(blockIdx.y) << 32) | blockIdx.x;
      r[2] = (static_cast<unsigned long long>(want) << 32) | have;
      r[3] = smid;
      r[4] = threadIdx.x;
    }
    __threadfence_system();
  }
  __trap();
}
__device__ __forceinline__ void SpinUntil(volatile uint32_t* cnt, uint32_t target, uint32_t site = 0) {
  uint32_t spins = 0;
  while (*cnt < target) {
    if (++spins > (1u << 28)) SpinTimeout(site, *cnt, target);
  }
}
__device__ __forceinline__ int ConvDil(int i) { return (i & 1) ? 1 : (i == 0 ? 1 : (i == 2 ? 3 : 5)); }
// sum of the dilations of convs 0..i-1 (1,1,3,1,5,1)
__device__ __forceinline__ int DilPrefix(int i) {
  return static_cast<int>((0xCB65210u >> (4 * i)) & 0xFu);   // {0, 1, 2, 5, 6, 11, 12} as nibbles: no local-memory table
}


// ---- thread-block cluster / distributed shared memory helpers ----
__device__ __forceinline__ uint32_t ClusterCtaRank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t MapToCta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void StCluster16(uint32_t cluster_addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(cluster_addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 16 bytes into a peer CTA's shared memory; the peer's mbarrier counts the bytes (complete_tx), so the
// receiver needs no fence: its wait on that barrier makes the data visible, exactly as for TMA.
__device__ __forceinline__ void StAsync16(uint32_t cluster_addr, float a, float b, float c, float d, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(cluster_addr),
               "f"(a), "f"(b), "f"(c), "f"(d), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ void MbarArriveCluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void MbarWaitCluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P2;\n"
      "LAB_WAIT_CL:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P2, [%0], %1;\n"
      "@P2 bra DONE_CL;\n"
      "bra LAB_WAIT_CL;\n"
      "DONE_CL:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void FenceCluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void ClusterSyncAll() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// GELU(v) = v/2 * (1 + erf(v / sqrt 2)), inline: erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7,
// far below the split-bf16 GEMM noise), exp on the SFU.  An out-of-line erff call per element forces
// the whole 16-column register tile through local memory around every call.
__device__ __forceinline__ float GeluFast(float v) {
  const float x = fabsf(v) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-x * x);   // erf(|v| / sqrt 2)
  return 0.5f * v * (1.0f + copysignf(e, v));
}
__device__ __noinline__ float SlowActTc(float v, int act) {   // tanh: post-conv only, never on this kernel's hot path
  if (act == kActGelu) return GeluFast(v);
  return tanhf(v);
}
__device__ __forceinline__ float ActTc(float v, int act) {
  if (act == kActLrelu) return v > 0.0f ? v : 0.1f * v;
  if (act == kActGelu) return GeluFast(v);
  return SlowActTc(v, act);
}

// 8 fp32 -> 8 bf16 (round to nearest even) packed in a uint4; optionally the bf16 of the residual.
// two fp32 -> packed bf16x2 (round to nearest even), one cvt.rn.bf16x2.f32: `a` in the low half
__device__ __forceinline__ uint32_t CvtBf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
template <bool kSplit>
__device__ __forceinline__ void Pack8(const float* v, uint4* hi, uint4* lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = CvtBf16x2(v[2 * i], v[2 * i + 1]);
    if (kSplit) {   // the bf16 of what the first rounding left: a bf16 is the upper half of its fp32
      const float ra = v[2 * i] - __uint_as_float(h[i] << 16);
      const float rb = v[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u);
      l[i] = CvtBf16x2(ra, rb);
    }
  }
  *hi = make_uint4(h[0], h[1], h[2], h[3]);
  if (kSplit) *lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void CpAsync16(uint32_t dst_smem, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;   // src-size 0 -> the 16 destination bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void CpAsyncWait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

// host-side bf16 helpers (round to nearest even)
inline uint16_t Bf16Rn(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);  // NaN
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return static_cast<uint16_t>(u >> 16);
}
inline float Bf16ToF(uint16_t h) {
  const uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

}  // namespace
}  // namespace b200

#endif  // BEATRICE_B200_TC_COMMON_CUH_
