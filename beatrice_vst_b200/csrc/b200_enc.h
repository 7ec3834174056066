// Fused residual stack of the encoders (b200_enc.cu): descriptors and launchers.
#ifndef BEATRICE_B200_ENC_H_
#define BEATRICE_B200_ENC_H_

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <vector>

#include "b200_mrf.h"

namespace b200 {

struct ResStackParams {
  const float* x_in;    // [B][C] fp32: output of the last front-end layer (current hop's row of every stream)
  float* x_out;         // [B][C] fp32: the stack's output
  uint16_t* xh_out;     // [B][C] bf16 hi (+ lo) copy of x_out for the head conv, or nullptr
  uint16_t* xl_out;
  const uint16_t* w;    // PackResStackWeights image
  const float* bias;    // [n_res][C]
  const float* gamma;   // [n_res][C]
  const float* beta;    // [n_res][C]
  uint16_t* hist;       // conv-input histories of all blocks, see ResStackHistElems
  int n_res;
  int dil[6];
  int B;
  int n_tiles;          // ceil(B / 32): stream tiles == clusters
  int trace;            // developer aid (BEATRICE_B200_ENC_TRACE=1): CTA 0 prints a per-block timeline
};

bool ResStackSupported(int C, int n_res, const int* dil);
int ResStackTiles(int B);
size_t ResStackHistElems(int C, int n_res, const int* dil, int B);
void ResStackHistBlocks(int C, int n_res, const int* dil, int B, uint16_t* base, std::vector<MrfHistBlock>* out);
// w[r] = fp32 [3][C][C] (tap, in, out) of block r; returns bf16 elements written (out may be null)
size_t PackResStackWeights(const float* const* w, int n_res, int C, uint16_t* out);
void LaunchResStack(const ResStackParams& p, int C, cudaStream_t s);

}  // namespace b200

#endif  // BEATRICE_B200_ENC_H_
