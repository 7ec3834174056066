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
