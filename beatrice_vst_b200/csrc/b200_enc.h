// Fused residual stack of the encoders (b200_enc.cu): descriptors and launchers.
#ifndef BEATRICE_B200_ENC_H_
#define BEATRICE_B200_ENC_H_

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <vector>

#include "b200_mrf.h"

namespace b200 {

// One launch runs a CHAIN of blocks on every tile of 32 streams:
//   kind 1 (optional, first):  x = GELU(b + Conv_{k=2, stride 2}(fin))        -- the last front-end layer
//   kind 0 (n times):          x = x + b + Conv_{k=3, dil}(GELU(ChanNorm(x) * gamma + beta))
//   kind 2 (optional, last):   head_out = b + W x                            -- the 1x1 head
//   kind 3:                    x = b + Conv_{k=7}(x)                         -- the vocoder's pre conv
//   kind 4 (first):            x = b + W x_in + pitch_emb[q] + Wf feat + spk (+ formant)
//                                                                            -- the vocoder's conditioning
//                                 (1x1 phone embedding; x_in has x_in_C <= C channels, W zero-padded to C)
struct ResStackParams {
  const float* x_in;    // [B][x_in_C] fp32 input of the first block when there is no kind-1 block
  int x_in_C;           // channels of x_in (0: C)
  // kind 4 only (reference semantics: oracle Generate(), conditioning): per-stream pitch bin, pitch features,
  // additive speaker embedding and (rc0) formant embedding
  const int* emb_q;         // [B]
  int emb_bins;
  const float* emb_pitch;   // [bins][C]
  const float* emb_feat;    // [B][4]
  const float* emb_wf;      // [4][C]
  const float* emb_spk;     // [B][C] or nullptr
  const float* emb_formant; // [B][C] or nullptr
  float* x_out;         // [B][out_slots][C] fp32: x after the last block (nullptr when the head is fused)
  uint16_t* xh_out;     // [B][out_slots][C] bf16 hi (+ lo) copy of out_act(x) for the consumer conv, or nullptr
  uint16_t* xl_out;
  int out_slots;        // ring slots of the outputs (row frame % out_slots is written); 0 or 1: flat
  int out_act;          // activation applied to the bf16 copy only: 0 none, 1 LeakyReLU(0.1)
  const int* frame;     // device hop counter (needed when out_slots > 1)
  const uint16_t* fin_h;  // kind 1: input rows, bf16 hi / lo planes [B][2][C]
  const uint16_t* fin_l;
  float* head_out;      // kind 2: [B][head_n] fp32
  float* head_out2;     // optional second copy of the head output (the consumer's input buffer), or nullptr
  int head_n;           // 128 or 256
  const uint16_t* w;    // PackChainWeights image
  const float* bias;    // [n_blk][C]
  const float* gamma;   // [n_blk][C] (kind 0 rows used)
  const float* beta;    // [n_blk][C]
  uint16_t* hist;       // conv-input histories of the kind-0 blocks, see ResStackHistElems
  int n_blk;
  int kind[8];
  int dil[8];           // kind 0 only
  int B;
  int n_tiles;          // ceil(B / 32): stream tiles == clusters
  int trace;            // developer aid (BEATRICE_B200_ENC_TRACE=1): CTA 0 prints a per-block timeline
};

struct ChainLayer {     // host description of one block's conv for PackChainWeights
  const float* w;       // fp32 [taps][C][n_out]
  int taps;
  int n_out;
};

bool ResStackSupported(int C, int n_res, const int* dil);
int ResStackTiles(int B);
size_t ResStackHistElems(int C, int n_res, const int* dil, int B);
void ResStackHistBlocks(int C, int n_res, const int* dil, int B, uint16_t* base, std::vector<MrfHistBlock>* out);
// returns bf16 elements written (out may be null)
size_t PackChainWeights(const ChainLayer* layers, int n_layers, int C, uint16_t* out);
void LaunchResStack(const ResStackParams& p, int C, cudaStream_t s);

}  // namespace b200

#endif  // BEATRICE_B200_ENC_H_
