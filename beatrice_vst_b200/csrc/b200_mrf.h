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

// The stage's ConvTranspose1d upsampler, computed in the prologue of every CTA of the fused kernel instead of
// by a launch of its own (single-CTA kernel, split precision): u[r i + ph][co] = b[co] + sum_ci
// xin[i][ci] W[1][ci][ph][co] + xin[i-1][ci] W[0][ci][ph][co] with xin = lrelu(mean of the previous stage's three
// branch outputs) -- a GEMM of S * T/r rows, N = r C, K = 2 * 2C whose result is re-laid out through shared memory.
struct MrfUpsDesc {
  const uint16_t* w;   // PackMrfUpsWeights image, nullptr = not fused (u is read from MrfStageParams::u)
  const float* bias;   // [C]
  const float* x[3];   // previous stage's branch outputs, fp32 rings [B][x_slots * T/r][2C]
  int x_slots;
  int r;               // upsampling rate
};

struct MrfStageParams {
  MrfBranchDesc br[3];
  MrfUpsDesc ups;
  const float* u;      // stage input (upsampler output), fp32 ring [B][u_slots * T][C]
  const float* film;   // [B][2C] = gamma | beta applied to u as u*(1+gamma)+beta in the prologue, or nullptr
  int u_slots;
  int T;               // rows per stream per hop
  int S;               // streams per CTA (rows inside a CTA are time-major: row = t * S + s)
  int MT;              // 128-row tiles per CTA = ceil(S * T / 128)
  int B;               // streams
  int n_groups;        // ceil(B / S)
  const int* frame;    // device hop counter
  int n_branches;      // grid.y: CTA classes of this launch (CTAs are scheduled in y order)
  // cluster kernel: blockIdx.y = y runs br[y2br[y]].  Single-CTA kernel: blockIdx.y = y runs the branches
  // br[yseq[y][0 .. ylen[y])] ONE AFTER THE OTHER in the same CTA (e.g. k = 7 then k = 3: 10 taps beside a k = 11
  // CTA's 11 -- two equal CTA classes instead of three unequal ones, and 2/3 of the CTAs)
  int y2br[3];
  int yseq[3][3];
  int ylen[3];
  int nb_max;          // max over y of ylen[y] (bias staging is sized for it)
  int smem_min;        // host only: lower bound on the launch's dynamic shared memory (bytes) -- a launch can ask
                       // for more than it needs so that fewer of its CTAs (or none of another launch) share an SM
  // 0: ordinary PDL kernel (dependency wait in front of the first read of u).
  // 1: second launch of a pair that together make up one stage: launched (programmatically) only after the
  //    first launch's CTAs passed THEIR wait, so u is already complete -- no wait at the start; instead it
  //    waits for the first launch at its very end, so that "this kernel complete" implies "stage complete".
  int pdl_mode;
  int trace;           // developer aid: 1 + blockIdx.x of the CTA whose timeline is printed (0 = off)
  int late_launch;     // host only: the stage's first launch starts when its predecessor has COMPLETED (no programmatic
                       // early launch): its CTAs then do not hold SMs while they wait for the stage input
};

// One history block, for per-stream reset: [group][planes * panels][H][S][8] bf16
struct MrfHistBlock {
  uint16_t* base;
  int planes_panels, H, S, pad_;
};

// kmax: the largest branch kernel size this launch runs (shared memory is sized for it)
bool MrfFusedSupported(int C, int T, int S, bool split, int kmax = 11, int nb = 1, bool fused_ups = false);
size_t MrfSmemBytes(int C, int T, int S, bool split, int kmax = 11, int nb = 1, bool fused_ups = false);
// fills yseq / ylen / nb_max for the one-branch-per-CTA form described by n_branches / y2br
void MrfOneBranchPerCta(MrfStageParams* p);
// bf16 elements of one branch's history state: for conv i (dilation dil_i) a block
// [group][plane][C/8 panels][(k-1)*dil_i * S rows][8], blocks in conv order
size_t MrfHistElems(int C, int k, int S, int n_groups, bool split);
// w[i] = fp32 [k][C][C] (tap, in, out) of conv i; returns bf16 elements written (out may be null)
// concat (split mode, single-CTA kernel): weight rows of a K step packed [W_hi ; W_lo] for the 2-MMA scheme
size_t PackMrfWeights(const float* const w[6], int k, int C, bool split, bool concat, uint16_t* out);
// upsampler image for MrfUpsDesc: w = fp32 [2 taps][2C][r*C] (tap 0 multiplies xin[i-1]); K step ks = tap * (2C/16) + g:
// [plane hi | lo][2 panels][r*C rows][8]; returns bf16 elements (out may be null)
size_t PackMrfUpsWeights(const float* w, int C, int r, uint16_t* out);
bool MrfUpsFusable(int C, int T, int S, int r, bool split);
void LaunchMrfStage(const MrfStageParams& p, int C, bool split, cudaStream_t s);
// K-split cluster form (b200_mrfc.cu): NC CTAs per (group, branch), same weight / history images
bool MrfClusterSupported(int C, int NC, int T, int S, bool split);
size_t MrfClusterSmemBytes(int C, int NC, int T, int S, bool split);
void LaunchMrfStageCluster(const MrfStageParams& p, int C, int NC, bool split, cudaStream_t s);
void LaunchMrfZeroStream(const MrfHistBlock* d_blocks, int n_blocks, int b, cudaStream_t s);

}  // namespace b200

#endif  // BEATRICE_B200_MRF_H_
