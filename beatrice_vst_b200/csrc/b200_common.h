// Shared definitions of libbeatrice_b200 (host side).
#ifndef BEATRICE_B200_COMMON_H_
#define BEATRICE_B200_COMMON_H_

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

// The per-frame ABI returns void (reference lib/beatricelib/beatrice.h:243-247, :266-271,
// :301-307), so a CUDA failure has no error channel: fail loudly, never fall back to a CPU.
#define B200_CHECK(expr)                                                                          \
  do {                                                                                            \
    cudaError_t err__ = (expr);                                                                   \
    if (err__ != cudaSuccess) {                                                                   \
      std::fprintf(stderr, "[libbeatrice_b200] FATAL %s:%d: %s -> %s\n", __FILE__, __LINE__, #expr, \
                   cudaGetErrorString(err__));                                                    \
      std::abort();                                                                               \
    }                                                                                             \
  } while (0)

namespace b200 {

// Programmatic dependent launch: the kernel may begin (prologue up to its griddepcontrol.wait)
// while the previous kernel on `s` drains.  BEATRICE_B200_NO_PDL=1 turns the attribute off.
inline bool PdlEnabled() {
  static const bool on = [] {
    const char* e = std::getenv("BEATRICE_B200_NO_PDL");
    return !(e && e[0] == '1');
  }();
  return on;
}
template <typename... KArgs, typename... Args>
inline void LaunchPdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x,
                      Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (PdlEnabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  B200_CHECK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

#ifdef __CUDACC__
// ---- programmatic dependent launch (PDL), device side ----
// A kernel launched with LaunchPdl may start while its predecessor on the stream is still
// running: everything before PdlWait() must touch only data no in-flight predecessor writes
// (weights, biases, barrier / TMEM setup); PdlWait() returns when the predecessor grid has
// completed and its memory is visible.  Without the launch attribute both are no-ops.
__device__ __forceinline__ void PdlLaunchDependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void PdlWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// reference lib/beatricelib/beatrice.h:10-28
constexpr int kInHop = 160;
constexpr int kOutHop = 240;
constexpr int kHidden = 256;
constexpr int kCodebookSize = 512;
constexpr int kKvLength = 384;
constexpr int kKvChannels = 128;
constexpr int kNBlocks = 4;
constexpr int kPitchFeatures = 4;
constexpr int kNFormant = 9;
constexpr int kHostHop48k = 480;

struct FamilyDims {
  int family;          // 0 = 20a2, 1 = 20b1, 2 = 20rc0
  int phone_channels;  // beatrice.h:17,20,23
  int pitch_bins;      // beatrice.h:18,21,24
  bool has_setter;     // rc0: EmbeddingSetter / FiLM / codebook VQ (beatrice.h:207-209)
};
constexpr FamilyDims kFamilies[3] = {{0, 256, 384, false}, {1, 256, 384, false}, {2, 128, 448, true}};

}  // namespace b200

#endif  // BEATRICE_B200_COMMON_H_
