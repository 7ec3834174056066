// micro-experiment: does PDL shorten a chain of small dependent kernels (stream and graph)?
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void k(float* p, int pre_spin, int use_pdl) {
  // "prologue": independent work
  long long t0 = clock64();
  while (clock64() - t0 < pre_spin) {}
  if (use_pdl) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  float v = p[threadIdx.x];
  t0 = clock64();
  while (clock64() - t0 < 4000) {}
  p[threadIdx.x] = v + 1.0f;
}
static void launch(float* p, int pre, bool pdl, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(8); cfg.blockDim = dim3(128); cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k, p, pre, pdl ? 1 : 0);
}
int main() {
  float* p; CK(cudaMalloc(&p, 4096)); CK(cudaMemset(p, 0, 4096));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int N = 50;
  for (int pre : {0, 4000}) for (int pdl = 0; pdl < 2; ++pdl) {
    for (int mode = 0; mode < 2; ++mode) {
      cudaGraphExec_t ex = nullptr;
      if (mode == 1) {
        cudaGraph_t g;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < N; ++i) launch(p, pre, pdl, s);
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ex, g, 0));
        size_t ne = 0; cudaGraphGetEdges(g, nullptr, nullptr, &ne);
        cudaGraphDestroy(g);
      }
      float best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, s);
        if (mode == 1) cudaGraphLaunch(ex, s); else for (int i = 0; i < N; ++i) launch(p, pre, pdl, s);
        cudaEventRecord(e1, s);
        CK(cudaStreamSynchronize(s));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("pre_spin %d pdl %d %s: %.2f us per kernel\n", pre, pdl, mode ? "graph " : "stream", best * 1e3 / N);
      if (ex) cudaGraphExecDestroy(ex);
    }
  }
  float h[128]; CK(cudaMemcpy(h, p, 512, cudaMemcpyDeviceToHost)); printf("check %f\n", h[0]);
  return 0;
}
