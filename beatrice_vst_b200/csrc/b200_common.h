// Shared definitions of libbeatrice_b200 (host side).
#ifndef BEATRICE_B200_COMMON_H_
#define BEATRICE_B200_COMMON_H_

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <exception>

// The per-frame ABI returns void and "must never fail visibly" (reference lib/beatricelib/beatrice.h:243-247,
// :266-271, :301-307; the call site's own guards zero the output and return a code,
// src/common/processor_core_2.cc:26-43).  A CUDA failure therefore never kills the host process:
// B200_CHECK latches a sticky library-wide error (first failure wins, one line on stderr), and throws
// b200::Failure, which every extern "C" entry point catches -- per-frame calls then write silence, loaders
// return an error code.  While the error is latched every later per-frame call short-circuits to silence;
// BeatriceB200_LastError / _LastErrorString / _ClearError (include/beatrice_b200.h) expose it.  There is still no
// CPU fallback.  BEATRICE_B200_ABORT_ON_ERROR=1 restores abort() for debugging.
namespace b200 {
struct Failure {
  int code;  // cudaError_t, or a negative BEATRICE_B200_ERR_* value for non-CUDA conditions
};
[[noreturn]] void Fail(int code, const char* what, const char* file, int line);
bool Failed();                 // a failure is latched
int LastErrorCode();           // 0 when none
const char* LastErrorText();   // "" when none; stable storage
void ClearError();
void NoteException(const char* what) noexcept;   // latches a host-side exception (bad_alloc ...) without throwing
}  // namespace b200

#define B200_CHECK(expr)                                                                          \
  do {                                                                                            \
    cudaError_t err__ = (expr);                                                                   \
    if (err__ != cudaSuccess) b200::Fail(static_cast<int>(err__), #expr, __FILE__, __LINE__);     \
  } while (0)

// Body of an extern "C" entry point: runs `...` unless a failure is latched; on failure runs `on_fail`.
#define B200_GUARDED(on_fail, ...)                 \
  do {                                             \
    if (b200::Failed()) {                          \
      on_fail;                                     \
    } else {                                       \
      try {                                        \
        __VA_ARGS__                                \
      } catch (const b200::Failure&) {             \
        on_fail;                                   \
      } catch (const std::exception& ex__) {       \
        b200::NoteException(ex__.what());          \
        on_fail;                                   \
      }                                            \
    }                                              \
  } while (0)

namespace b200 {

// Load-time host -> device upload.  A cudaMemcpy from pageable memory may return once the bytes are staged, before
// the DMA into `dst` has finished, and the kernels that read `dst` run on cudaStreamNonBlocking streams, which do not
// order against the legacy stream such a copy drains on.  So the copy goes through a private non-blocking stream
// that is synchronised before anyone can consume the data.  (Not cudaDeviceSynchronize: other host threads may be
// capturing their hop graphs at this moment -- independent plug-in instances -- and a device-wide sync is illegal
// while any capture is open.)
inline void UploadSync(void* dst, const void* src, size_t bytes) {
  cudaStream_t up = nullptr;
  B200_CHECK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
  cudaError_t err = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, up);
  if (err == cudaSuccess) err = cudaStreamSynchronize(up);
  cudaStreamDestroy(up);
  B200_CHECK(err);
}

// Same for the zero-fill of a fresh allocation (cudaMemset on device memory is asynchronous to the host too).
inline void ZeroSync(void* dst, size_t bytes) {
  cudaStream_t up = nullptr;
  B200_CHECK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
  cudaError_t err = cudaMemsetAsync(dst, 0, bytes, up);
  if (err == cudaSuccess) err = cudaStreamSynchronize(up);
  cudaStreamDestroy(up);
  B200_CHECK(err);
}

// Programmatic dependent launch: the kernel may begin (prologue up to its griddepcontrol.wait)
// while the previous kernel on `s` drains.  BEATRICE_B200_NO_PDL=1 turns the attribute off.
inline bool PdlEnabled() {
  static const bool on = [] {
    const char* e = std::getenv("BEATRICE_B200_NO_PDL");
    return !(e && e[0] == '1');
  }();
  return on;
}
// early: allow the kernel to start before its predecessor on the stream has completed (programmatic dependent launch)
template <typename... KArgs, typename... Args>
inline void LaunchMaybePdl(bool early, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x,
                           Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (early && PdlEnabled()) {
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
template <typename... KArgs, typename... Args>
inline void LaunchPdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x,
                      Args... args) {
  LaunchMaybePdl(true, kernel, grid, block, smem, s, cluster_x, args...);
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
