// Fused MRF branch kernel of the vocoder (b200_mrf.cu): descriptors and launchers.
#ifndef BEATRICE_B200_MRF_H_
#define BEATRICE_B200_MRF_H_

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace b200 {

// One MRF branch (kernel size k) of one vocoder stage.
struct MrfBranchDesc {
  const uint16_t* w;   // PackMrfWeights image of the six convs (c1_d1, c2_d1, c1_d3, c2_d3, c1_d5, c2_d5)
  const float* bias;   // [6][C]
  uint16_t* hist;      // conv-input histories of all stream groups, see MrfHistElems
  float* out;          // fp32 ring [B][out_slots * T][C]; the hop's rows are written
  int out_slots;
  int k;
};

struct MrfStageParams {
  MrfBranchDesc br[3];
  const float* u;      // stage input (upsampler output), fp32 ring [B][u_slots * T][C]
  const float* film;   // [B][2C] = gamma | beta applied to u as u*(1+gamma)+beta in the prologue, or nullptr
  int u_slots;
  int T;               // rows per stream per hop
  int S;               // streams per CTA (rows inside a CTA are time-major: row = t * S + s)
  int MT;              // 128-row tiles per CTA = ceil(S * T / 128)
  int B;               // streams
  int n_groups;        // ceil(B / S)
  const int* frame;    // device hop counter
  int n_branches;      // grid.y: branches this launch runs; blockIdx.y = y runs br[y2br[y]] (CTAs are scheduled in y order)
  int y2br[3];
  // 0: ordinary PDL kernel (dependency wait in front of the first read of u).
  // 1: second launch of a pair that together make up one stage: launched (programmatically) only after the
  //    first launch's CTAs passed THEIR wait, so u is already complete -- no wait at the start; instead it
  //    waits for the first launch at its very end, so that "this kernel complete" implies "stage complete".
  int pdl_mode;
  int trace;           // developer aid: 1 + blockIdx.x of the CTA whose timeline is printed (0 = off)
};

// One history block, for per-stream reset: [group][planes * panels][H][S][8] bf16
struct MrfHistBlock {
  uint16_t* base;
  int planes_panels, H, S, pad_;
};

// kmax: the largest branch kernel size this launch runs (shared memory is sized for it)
bool MrfFusedSupported(int C, int T, int S, bool split, int kmax = 11);
size_t MrfSmemBytes(int C, int T, int S, bool split, int kmax = 11);
// bf16 elements of one branch's history state: for conv i (dilation dil_i) a block
// [group][plane][C/8 panels][(k-1)*dil_i * S rows][8], blocks in conv order
size_t MrfHistElems(int C, int k, int S, int n_groups, bool split);
// w[i] = fp32 [k][C][C] (tap, in, out) of conv i; returns bf16 elements written (out may be null)
// concat (split mode, single-CTA kernel): weight rows of a K step packed [W_hi ; W_lo] for the 2-MMA scheme
size_t PackMrfWeights(const float* const w[6], int k, int C, bool split, bool concat, uint16_t* out);
void LaunchMrfStage(const MrfStageParams& p, int C, bool split, cudaStream_t s);
// K-split cluster form (b200_mrfc.cu): NC CTAs per (group, branch), same weight / history images
bool MrfClusterSupported(int C, int NC, int T, int S, bool split);
size_t MrfClusterSmemBytes(int C, int NC, int T, int S, bool split);
void LaunchMrfStageCluster(const MrfStageParams& p, int C, int NC, bool split, cudaStream_t s);
void LaunchMrfZeroStream(const MrfHistBlock* d_blocks, int n_blocks, int b, cudaStream_t s);

}  // namespace b200

#endif  // BEATRICE_B200_MRF_H_
