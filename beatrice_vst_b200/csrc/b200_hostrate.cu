#include "b200_hostrate.h"

#include <cmath>
#include <cstring>

#include "b200_kernels.h"

namespace b200 {
namespace {

constexpr double kPi = 3.14159265358979323846;  // == std::numbers::pi
constexpr int kTaps = 33;      // filter_size(32) * ratio_high(1) + 1, resample.h:210
constexpr int kHop = kHostHop48k;
constexpr int kGHist = 30;     // oldest sample the decimating FIR reaches: n - 30
constexpr int kZHist = 16;     // 24 kHz samples of hop c-2 the interpolating FIR reaches

// Gain::Process amplitude recurrence (gain.h:41-71) for sample index 0..479; thread 0 only.
__device__ void GainScan(const GainSeg& s, double* amp) {
  double a = s.amp0;
  for (int n = 0; n < kHop; ++n) {
    if (s.mode == 1) {
      if (a < s.target) a = fmin(__dmul_rn(a, s.ratio), s.target);
    } else if (s.mode == 2) {
      if (a > s.target) a = fmax(__dmul_rn(a, s.ratio), s.target);
    }
    amp[n] = a;
  }
}

// gain_in -> decimating FIR evaluated only at the kept samples 3i+2.
__global__ void __launch_bounds__(160) hostrate_in_kernel(const float* __restrict__ in48, float* __restrict__ g_ring,
                                                          const GainSeg* __restrict__ seg,
                                                          const float* __restrict__ cd, float* __restrict__ x16,
                                                          const int* __restrict__ frame_ptr) {
  __shared__ float g[kGHist + kHop];
  __shared__ double amp[kHop];
  __shared__ float c[kTaps];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int frame = *frame_ptr;
  const int cur = frame & 1, prev = cur ^ 1;
  const GainSeg s = seg[b];
  if (tid < kTaps) c[tid] = cd[tid];
  if (s.mode != 0) {
    if (tid == 0) GainScan(s, amp);
    __syncthreads();
  }
  float* ring_b = g_ring + static_cast<long long>(b) * 2 * kHop;
  if (tid < kGHist) g[tid] = ring_b[prev * kHop + (kHop - kGHist) + tid];
#pragma unroll
  for (int j = 0; j < kHop / 160; ++j) {
    const int n = tid + 160 * j;
    const double a = s.mode != 0 ? amp[n] : s.amp0;
    // output[i] = static_cast<float>(input[i] * current_amplitude)   gain.h:54,62,68
    const float v = __double2float_rn(__dmul_rn(static_cast<double>(in48[static_cast<long long>(b) * kHop + n]), a));
    g[kGHist + n] = v;
    ring_b[cur * kHop + n] = v;
  }
  __syncthreads();
  // resample.h:149-157 with ratio 1/1: out = sum_{m=1..31} buf[-m] * coef[m], buf[-1] = newest
  const int n = 3 * tid + 2;
  float acc = 0.0f;
#pragma unroll
  for (int m = 1; m < kTaps - 1; ++m) acc = __fadd_rn(acc, __fmul_rn(g[kGHist + n - m + 1], c[m]));
  x16[b * kInHop + tid] = acc;  // "* gain" with gain == 1.0f is exact
}

// zero-stuffed interpolating FIR over the PREVIOUS hop's model output -> gain_out.
// The block handed back by hop c is built from the model outputs of hops c-1 and c-2 only (the block FIFO of
// resample.h:343-363), so the two halves of this kernel can run apart: `compute` (out48 from the ring; may run
// before the hop's model call, with the hop index passed by value in frame_value >= 0) and `store` (this hop's
// model output into the ring for hop c+1; the last block to finish then advances the hop counter).
__global__ void __launch_bounds__(kHop) hostrate_out_kernel(const float* __restrict__ o24, float* __restrict__ o_ring,
                                                            const GainSeg* __restrict__ seg,
                                                            const float* __restrict__ cu, float* __restrict__ out48,
                                                            int* __restrict__ frame_ptr, int frame_value, int compute,
                                                            int store, int* __restrict__ done) {
  __shared__ float z[kZHist + kOutHop];
  __shared__ double amp[kHop];
  __shared__ float c[kTaps];
  __shared__ int frame_s;
  const int b = blockIdx.x, tid = threadIdx.x;
  // ONE read of the hop counter per block, published through shared memory: the last block to arrive below
  // advances the counter, and no thread of any block may look at it after that
  if (tid == 0) frame_s = frame_value >= 0 ? frame_value : *frame_ptr;
  __syncthreads();
  const int frame = frame_s;
  const int cur = frame % 3, p1 = (frame + 2) % 3, p2 = (frame + 1) % 3;
  float* ring_b = o_ring + static_cast<long long>(b) * 3 * kOutHop;
  if (store) {
    PdlWait();   // o24 is the previous kernel's output (no-op unless launched with the PDL attribute)
    if (tid < kOutHop) ring_b[cur * kOutHop + tid] = o24[b * kOutHop + tid];  // becomes hop c+1's input
    if (tid == 0 && atomicAdd(done, 1) == static_cast<int>(gridDim.x) - 1) {
      // every block's thread 0 read the counter before arriving here (and its other threads use frame_s); the last
      // one to arrive advances it (wrap: see advance_kernel)
      *done = 0;
      *frame_ptr = (frame + 1 >= 738017280) ? 0 : frame + 1;
    }
  }
  if (!compute) return;
  const GainSeg s = seg[b];
  if (tid < kTaps) c[tid] = cu[tid];
  if (tid < kOutHop) {
    z[kZHist + tid] = ring_b[p1 * kOutHop + tid];
  } else if (tid < kOutHop + kZHist) {
    const int j = tid - kOutHop;
    z[j] = ring_b[p2 * kOutHop + (kOutHop - kZHist) + j];
  }
  if (s.mode != 0 && tid == kHop - 1) GainScan(s, amp);
  __syncthreads();
  // resample.h:193-200 with ratio 1/1: out = sum_{i=0..31} buf[-1-i] * coef[i]; the stuffed
  // zeros (odd positions, resample.h:390-393) contribute +-0 and are skipped
  float acc = 0.0f;
#pragma unroll
  for (int i = 0; i < kTaps - 1; ++i) {
    const int idx = tid - i;
    if ((idx & 1) == 0) acc = __fadd_rn(acc, __fmul_rn(z[kZHist + (idx >> 1)], c[i]));
  }
  const double a = s.mode != 0 ? amp[tid] : s.amp0;
  out48[static_cast<long long>(b) * kHop + tid] = __double2float_rn(__dmul_rn(static_cast<double>(acc), a));
}

double DbToAmp(double db) { return std::pow(10.0, db * 0.05); }      // gain.h:12-14
double AmpToDb(double amp) { return 20.0 * std::log10(amp); }        // gain.h:15-17
double NormalizedSinc(double x) {                                     // resample.h:17-23
  const double pi = kPi;
  if (std::abs(x) < 1e-8) return 1.0;
  return std::sin(x * pi) / (x * pi);
}

}  // namespace

void HostRateState::Init(int device, int B) {
  device_ = device;
  B_ = B;
  in48_.Alloc(device, sizeof(float) * B * kHop, true);
  out48_.Alloc(device, sizeof(float) * B * kHop, true);
  g_ring_.Alloc(device, sizeof(float) * B * 2 * kHop, true);
  o_ring_.Alloc(device, sizeof(float) * B * 3 * kOutHop, true);
  frame_.Alloc(device, sizeof(int), true);
  done_.Alloc(device, sizeof(int), true);
  host_frame_ = 0;
  seg_in_.Alloc(device, sizeof(GainSeg) * B, true);
  seg_out_.Alloc(device, sizeof(GainSeg) * B, true);
  // filter tables: DownUpSamplerImpl::Reset (resample.h:209-230) for outer = inner = 48 kHz,
  // cut-offs from AnyFreqInOut (resample.h:412-417)
  const double sample_rate = 48000.0;
  const double cutoff_down = 0.99 * 16000.0 / std::min(std::max(sample_rate, 16000.0), 48000.0);
  const double cutoff_up = 0.99 * 24000.0 / std::min(std::max(sample_rate, 24000.0), 48000.0);
  const int ratio_high = 1;
  const int coef_length = 32 * ratio_high + 1;
  const int center_idx = coef_length / 2;
  float table[2 * kTaps];
  for (int i = 0; i < coef_length; ++i) {
    const double sinc_down =
        NormalizedSinc(static_cast<double>(i - center_idx) / static_cast<double>(ratio_high) * cutoff_down);
    const double sinc_up =
        NormalizedSinc(static_cast<double>(i - center_idx) / static_cast<double>(ratio_high) * cutoff_up);
    const double window =
        0.5 - 0.5 * std::cos(kPi * 2.0 / static_cast<double>(coef_length - 1) * static_cast<double>(i));
    table[i] = static_cast<float>(cutoff_down * sinc_down * window);
    table[kTaps + i] = static_cast<float>(cutoff_up * sinc_up * window);
  }
  coef_.Alloc(device, sizeof(table), false);
  UploadSync(coef_.p, table, sizeof(table));
  gin_.assign(B, HostGain());
  gout_.assign(B, HostGain());
  hseg_in_.assign(B, GainSeg{1.0, 1.0, 1.0, 0, 0});
  hseg_out_ = hseg_in_;
  up_in_.clear();
  up_out_.clear();
  lag_out_.clear();
  uploaded_ = false;
}

void HostRateState::SetTargetGain(int b, bool input, double db) {
  HostGain& g = input ? gin_[b] : gout_[b];
  g.target_db = db;  // Gain::Context::SetTargetGain, gain.h:28
  g.settled = false;
}

void HostRateState::ResetStream(int b, cudaStream_t s) {
  B200_CHECK(cudaMemsetAsync(g_ring_.as<float>() + static_cast<size_t>(b) * 2 * kHop, 0, sizeof(float) * 2 * kHop, s));
  B200_CHECK(cudaMemsetAsync(o_ring_.as<float>() + static_cast<size_t>(b) * 3 * kOutHop, 0, sizeof(float) * 3 * kOutHop, s));
}

void HostRateState::PrepareHop(cudaStream_t s, bool out_lag) {
  // Gain::Process (gain.h:41-71) on the host for the scalar state; the device replays the
  // same recurrence per sample.
  auto step = [](HostGain* g, GainSeg* seg) {
    if (g->settled) return;
    const double sample_rate = 48000.0;
    const double target = DbToAmp(g->target_db);
    double cur = DbToAmp(g->current_db);
    seg->amp0 = cur;
    seg->target = target;
    seg->ratio = 1.0;
    seg->mode = 0;
    int i = 0;
    if (cur < target) {
      seg->mode = 1;
      seg->ratio = DbToAmp(2.0 / (sample_rate * 0.001));
      while (i < kHop && cur < target) {
        cur = std::min(cur * seg->ratio, target);
        ++i;
      }
    } else if (cur > target) {
      seg->mode = 2;
      seg->ratio = DbToAmp(-2.0 / (sample_rate * 0.001));
      while (i < kHop && cur > target) {
        cur = std::max(cur * seg->ratio, target);
        ++i;
      }
    }
    const double new_db = AmpToDb(cur);
    g->settled = (std::memcmp(&new_db, &g->current_db, sizeof(double)) == 0);
    g->current_db = new_db;
  };
  std::vector<GainSeg> prev_out;
  if (out_lag) prev_out = lag_out_.empty() ? std::vector<GainSeg>(B_, GainSeg{1.0, 1.0, 1.0, 0, 0}) : lag_out_;
  for (int b = 0; b < B_; ++b) {
    step(&gin_[b], &hseg_in_[b]);
    step(&gout_[b], &hseg_out_[b]);
  }
  if (out_lag) {
    lag_out_ = hseg_out_;       // this call's segment: applied by the next call
    hseg_out_.swap(prev_out);   // uploaded now: the previous call's
  }
  const size_t bytes = sizeof(GainSeg) * B_;
  if (!uploaded_ || std::memcmp(up_in_.data(), hseg_in_.data(), bytes) != 0) {
    B200_CHECK(cudaMemcpyAsync(seg_in_.p, hseg_in_.data(), bytes, cudaMemcpyHostToDevice, s));
    up_in_ = hseg_in_;
  }
  if (!uploaded_ || std::memcmp(up_out_.data(), hseg_out_.data(), bytes) != 0) {
    B200_CHECK(cudaMemcpyAsync(seg_out_.p, hseg_out_.data(), bytes, cudaMemcpyHostToDevice, s));
    up_out_ = hseg_out_;
  }
  uploaded_ = true;
  if (out_lag) hseg_out_ = lag_out_;   // the host-side recurrence continues from the un-lagged state
}

void HostRateState::EnqueueIn(float* x16, cudaStream_t s) {
  hostrate_in_kernel<<<B_, 160, 0, s>>>(in48_.as<float>(), g_ring_.as<float>(), seg_in_.as<GainSeg>(),
                                        coef_.as<float>(), x16, frame_.as<int>());
  B200_CHECK(cudaGetLastError());
}

void HostRateState::EnqueueOut(const float* o24, cudaStream_t s) {
  hostrate_out_kernel<<<B_, kHop, 0, s>>>(o24, o_ring_.as<float>(), seg_out_.as<GainSeg>(), coef_.as<float>() + kTaps,
                                          out48_.as<float>(), frame_.as<int>(), -1, 1, 1, done_.as<int>());
  B200_CHECK(cudaGetLastError());
}

void HostRateState::EnqueueOutEarly(cudaStream_t s) {
  hostrate_out_kernel<<<B_, kHop, 0, s>>>(nullptr, o_ring_.as<float>(), seg_out_.as<GainSeg>(), coef_.as<float>() + kTaps,
                                          out48_.as<float>(), frame_.as<int>(), host_frame_, 1, 0, done_.as<int>());
  B200_CHECK(cudaGetLastError());
}

void HostRateState::EnqueueStore(const float* o24, cudaStream_t s) {
  LaunchPdl(hostrate_out_kernel, dim3(B_), dim3(kHop), 0, s, 1, o24, o_ring_.as<float>(), seg_out_.as<GainSeg>(),
            coef_.as<float>() + kTaps, out48_.as<float>(), frame_.as<int>(), -1, 0, 1, done_.as<int>());
  B200_CHECK(cudaGetLastError());
}

void HostRateState::HopDone() { host_frame_ = (host_frame_ + 1 >= 738017280) ? 0 : host_frame_ + 1; }

}  // namespace b200
