#include "b200_anyrate.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "b200_hostrate.h"
#include "b200_kernels.h"

namespace b200 {
namespace {

constexpr double kPi = 3.14159265358979323846;  // == std::numbers::pi
constexpr int kFilterSize = 32;                  // AnyFreqInOut, resample.h:416
constexpr int kFifo = kHostHop48k;               // ConvertStreamFunctionBlockSize<80 * 6>, resample.h:405-406
constexpr int kThreads = 256;

struct GainSegDev {
  double amp0, ratio, target;
  int mode, n_slew;
};

struct ResampleArgs {
  const float* in;      // [B][in_pitch], q samples used
  float* out;           // [B][out_pitch], N samples written
  float* hist;          // [B][H]: the last H samples of the input sequence (after the gain, if it is applied in front)
  const float* coef;    // windowed-sinc table of this direction, L coefficients
  const GainSegDev* seg;
  int in_pitch, out_pitch, q, N, H;
  int down;             // 1: Downsample (resample.h:130-164), 0: Upsample (:168-206)
  int fc0, rh, rl, L;   // fraction clock at the start of the call, ratio high / low, table length
  float gainmul;        // Downsample: static_cast<float>(ratio_low) / static_cast<float>(ratio_high)
  int gain_where;       // 0: none, 1: Gain::Process on the input, 2: on the output
};

// output[i] = static_cast<float>(input[i] * current_amplitude)   gain.h:54,62,68
__device__ __forceinline__ float ApplyGain(float x, double a) { return __double2float_rn(__dmul_rn(static_cast<double>(x), a)); }

// Gain::Process over buf[0..n) in place: the slewing head sequentially (one thread), the settled tail in parallel.
__device__ void GainBlock(float* buf, int n, const GainSegDev& s) {
  const int tid = threadIdx.x;
  const int head = s.mode != 0 ? min(s.n_slew, n) : 0;
  const double steady = s.mode != 0 ? s.target : s.amp0;   // once the recurrence stops, the amplitude sits on the target
  for (int i = head + tid; i < n; i += blockDim.x) buf[i] = ApplyGain(buf[i], steady);
  if (tid == 0 && head > 0) {
    double a = s.amp0;
    for (int i = 0; i < head; ++i) {
      a = s.mode == 1 ? fmin(__dmul_rn(a, s.ratio), s.target) : fmax(__dmul_rn(a, s.ratio), s.target);
      buf[i] = ApplyGain(buf[i], a);
    }
  }
}

__global__ void __launch_bounds__(kThreads) anyrate_resample_kernel(const ResampleArgs a) {
  extern __shared__ float sm[];
  float* seq = sm;                  // [H + q]: history, then the call's input samples
  float* res = sm + a.H + a.q;      // [N]
  const int b = blockIdx.x, tid = threadIdx.x;
  float* hist = a.hist + static_cast<size_t>(b) * a.H;
  for (int i = tid; i < a.H; i += kThreads) seq[i] = hist[i];
  for (int i = tid; i < a.q; i += kThreads) seq[a.H + i] = a.in[static_cast<size_t>(b) * a.in_pitch + i];
  __syncthreads();
  if (a.gain_where == 1) {
    GainBlock(seq + a.H, a.q, a.seg[b]);
    __syncthreads();
  }
  const float* cur = seq + a.H;     // cur[i]: i-th sample of this call, cur[-1]: the newest sample of the history
  for (int j = tid; j < a.N; j += kThreads) {
    float acc = 0.0f;
    if (a.down) {
      // the j-th output of the call is produced right after input i, the first with fc0 + (i+1) rl >= (j+1) rh;
      // its clock is what is left over (resample.h:143-156)
      const int need = (j + 1) * a.rh - a.fc0;
      const int i = (need + a.rl - 1) / a.rl - 1;
      const int frac = a.fc0 + (i + 1) * a.rl - (j + 1) * a.rh;
      int idx = i;
      for (int f = a.rl - frac; f < a.L - 1; f += a.rl) acc = __fadd_rn(acc, __fmul_rn(cur[idx--], a.coef[f]));
      acc = __fmul_rn(acc, a.gainmul);
    } else {
      // output o: the clock has advanced o+1 times by rl; every wrap pushed one more input (resample.h:186-200)
      const int c = a.fc0 + (j + 1) * a.rl;
      const int cnt = c / a.rh, frac = c - cnt * a.rh;
      int idx = cnt - 1;
      for (int f = frac; f < a.L - 1; f += a.rh) acc = __fadd_rn(acc, __fmul_rn(cur[idx--], a.coef[f]));
    }
    res[j] = acc;
  }
  __syncthreads();
  if (a.gain_where == 2) {
    GainBlock(res, a.N, a.seg[b]);
    __syncthreads();
  }
  for (int j = tid; j < a.N; j += kThreads) a.out[static_cast<size_t>(b) * a.out_pitch + j] = res[j];
  for (int i = tid; i < a.H; i += kThreads) hist[i] = seq[a.q + i];   // the last H samples of history + input
}

// ConvertStreamFunctionBlockSize (resample.h:343-363): the call's samples [pos, pos+len) swap places with FIFO slots
// [idx, idx+len) -- what leaves is the previously processed block
__global__ void anyrate_fifo_swap_kernel(const float* __restrict__ x48, float* __restrict__ y48, int pitch, int pos, int len,
                                         float* __restrict__ fifo, int idx) {
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < len; t += blockDim.x) {
    float* f = fifo + static_cast<size_t>(b) * kFifo + idx + t;
    const float old = *f;
    *f = x48[static_cast<size_t>(b) * pitch + pos + t];
    y48[static_cast<size_t>(b) * pitch + pos + t] = old;
  }
}
// ConvertStreamFunctionFrom2In3OutTo6InOut (resample.h:384-394)
__global__ void anyrate_pick_kernel(const float* __restrict__ fifo, float* __restrict__ x16) {
  const int b = blockIdx.x, i = threadIdx.x;
  if (i < kInHop) x16[b * kInHop + i] = fifo[static_cast<size_t>(b) * kFifo + (i + 1) * 3 - 1];
}
__global__ void anyrate_stuff_kernel(const float* __restrict__ o24, float* __restrict__ fifo) {
  const int b = blockIdx.x, i = threadIdx.x;
  if (i < kOutHop) {
    fifo[static_cast<size_t>(b) * kFifo + 2 * i] = o24[b * kOutHop + i];
    fifo[static_cast<size_t>(b) * kFifo + 2 * i + 1] = 0.0f;
  }
}
__global__ void echo_model_kernel(const float* __restrict__ x16, float* __restrict__ o24) {
  const int b = blockIdx.x, i = threadIdx.x;
  if (i < kOutHop) o24[b * kOutHop + i] = i < kInHop ? x16[b * kInHop + i] : 0.0f;
}

double DbToAmp(double db) { return std::pow(10.0, db * 0.05); }      // gain.h:12-14
double AmpToDb(double amp) { return 20.0 * std::log10(amp); }        // gain.h:15-17
double NormalizedSinc(double x) {                                     // resample.h:17-23
  if (std::abs(x) < 1e-8) return 1.0;
  return std::sin(x * kPi) / (x * kPi);
}
// resample.h:25-46: Stern-Brocot search for numer / denom < 1000
void SimpleFraction(double ratio, int* numer, int* denom) {
  int ln = 0, ld = 1, rn = 1, rd = 0;
  for (;;) {
    const int mn = ln + rn, md = ld + rd;
    if (ratio * md < mn) {
      if (mn >= 1000 || md >= 1000) {
        *numer = ln;
        *denom = ld;
        return;
      }
      rn = mn;
      rd = md;
    } else {
      if (mn >= 1000 || md >= 1000) {
        *numer = rn;
        *denom = rd;
        return;
      }
      ln = mn;
      ld = md;
    }
  }
}

size_t ResampleSmem(int H, int q, int N) { return sizeof(float) * (static_cast<size_t>(H) + q + N); }

}  // namespace

bool AnyRateState::Init(int device, int B, double sample_rate, const HostRateState* seed) {
  device_ = device;
  B_ = B;
  rate_ = 0.0;
  if (!(sample_rate > 0.0)) return false;                                 // resample.h:243-246
  const double inner = 48000.0;
  down_first_ = sample_rate >= inner;                                      // resample.h:247
  const double high = down_first_ ? sample_rate : inner, low = down_first_ ? inner : sample_rate;
  // cut-offs from AnyFreqInOut (resample.h:412-417); which table gets which: resample.h:248-258
  const double cut_in = 0.99 * 16000.0 / std::min(std::max(sample_rate, 16000.0), 48000.0);
  const double cut_out = 0.99 * 24000.0 / std::min(std::max(sample_rate, 24000.0), 48000.0);
  const double cut_down = down_first_ ? cut_in : cut_out, cut_up = down_first_ ? cut_out : cut_in;
  int numer = 0, denom = 0;
  SimpleFraction(high / low, &numer, &denom);
  if (numer == 0 || denom == 0) return false;                              // resample.h:261-264
  rh_ = numer;
  rl_ = denom;
  // DownUpSamplerImpl::Reset, resample.h:209-236
  L_ = kFilterSize * rh_ + 1;
  const int center = L_ / 2;
  std::vector<float> table(static_cast<size_t>(2) * L_);
  for (int i = 0; i < L_; ++i) {
    const double sinc_down = NormalizedSinc(static_cast<double>(i - center) / static_cast<double>(rh_) * cut_down);
    const double sinc_up = NormalizedSinc(static_cast<double>(i - center) / static_cast<double>(rh_) * cut_up);
    const double window = 0.5 - 0.5 * std::cos(kPi * 2.0 / static_cast<double>(L_ - 1) * static_cast<double>(i));
    table[i] = static_cast<float>(cut_down * sinc_down * window);
    table[L_ + i] = static_cast<float>(cut_up * sinc_up * window);
  }
  fc_down_ = rh_ - 1;
  fc_up_ = rh_ - 1;
  hist_high_ = kFilterSize * rh_ / rl_ + 1;
  hist_low_ = kFilterSize + 1;
  fifo_idx_ = 0;
  // 48 kHz samples a call of kMaxBlock host samples can produce
  cap48_ = down_first_ ? kMaxBlock + 8 : static_cast<int>((static_cast<long long>(kMaxBlock) + 1) * rh_ / rl_) + 8;
  const int h_in = down_first_ ? hist_high_ : hist_low_, h_out = down_first_ ? hist_low_ : hist_high_;
  if (ResampleSmem(h_in, kMaxBlock, cap48_) > 200 * 1024 || ResampleSmem(h_out, cap48_, kMaxBlock) > 200 * 1024) return false;
  B200_CHECK(cudaSetDevice(device));
  B200_CHECK(cudaFuncSetAttribute(anyrate_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  coef_.Alloc(device, sizeof(float) * table.size(), false);
  UploadSync(coef_.p, table.data(), sizeof(float) * table.size());
  in_.Alloc(device, sizeof(float) * B * kMaxBlock, true);
  out_.Alloc(device, sizeof(float) * B * kMaxBlock, true);
  x48_.Alloc(device, sizeof(float) * B * cap48_, true);
  y48_.Alloc(device, sizeof(float) * B * cap48_, true);
  hist_in_.Alloc(device, sizeof(float) * B * h_in, true);     // Buffer starts as zeros (resample.h:55-59)
  hist_out_.Alloc(device, sizeof(float) * B * h_out, true);
  fifo_.Alloc(device, sizeof(float) * B * kFifo, true);       // buffer_() value-initialised (resample.h:339)
  seg_in_.Alloc(device, sizeof(GainSegDev) * B, true);
  seg_out_.Alloc(device, sizeof(GainSegDev) * B, true);
  if (static_cast<int>(gin_.size()) != B) {   // first set-up: start from the gains the engine has been given so far
    gin_.assign(B, HostGain());
    gout_.assign(B, HostGain());
    for (int b = 0; seed && b < B; ++b) {
      seed->GetGain(b, true, &gin_[b].target_db, &gin_[b].current_db);
      seed->GetGain(b, false, &gout_[b].target_db, &gout_[b].current_db);
    }
  }
  hseg_in_.assign(B, Seg{1.0, 1.0, 1.0, 0, 0});
  hseg_out_ = hseg_in_;
  rate_ = sample_rate;
  return true;
}

void AnyRateState::SetTargetGain(int b, bool input, double db) {
  if (b < 0 || b >= static_cast<int>(gin_.size())) return;
  (input ? gin_[b] : gout_[b]).target_db = db;   // Gain::Context::SetTargetGain, gain.h:28
}

// Gain::Process (gain.h:41-71) for the scalar state of one call of m samples; the device replays the recurrence
void AnyRateState::StepGain(HostGain* g, Seg* seg, int m) const {
  const double target = DbToAmp(g->target_db);
  double cur = DbToAmp(g->current_db);
  seg->amp0 = cur;
  seg->target = target;
  seg->ratio = 1.0;
  seg->mode = 0;
  seg->n_slew = 0;
  int i = 0;
  if (cur < target) {
    seg->mode = 1;
    seg->ratio = DbToAmp(2.0 / (rate_ * 0.001));
    while (i < m && cur < target) {
      cur = std::min(cur * seg->ratio, target);
      ++i;
    }
  } else if (cur > target) {
    seg->mode = 2;
    seg->ratio = DbToAmp(-2.0 / (rate_ * 0.001));
    while (i < m && cur > target) {
      cur = std::max(cur * seg->ratio, target);
      ++i;
    }
  }
  seg->n_slew = i;
  g->current_db = AmpToDb(cur);
}

void AnyRateState::Process(const float* in_host, float* out_host, int m, float* x16_dev, const float* o24_dev,
                           const std::function<void()>& run_hop, cudaStream_t s, uint64_t* launches) {
  const size_t row = sizeof(float) * m;
  B200_CHECK(cudaMemcpy2DAsync(in_.p, sizeof(float) * kMaxBlock, in_host, row, row, B_, cudaMemcpyHostToDevice, s));
  std::vector<GainSegDev> up_in(B_), up_out(B_);
  for (int b = 0; b < B_; ++b) {
    StepGain(&gin_[b], &hseg_in_[b], m);
    StepGain(&gout_[b], &hseg_out_[b], m);
    up_in[b] = GainSegDev{hseg_in_[b].amp0, hseg_in_[b].ratio, hseg_in_[b].target, hseg_in_[b].mode, hseg_in_[b].n_slew};
    up_out[b] = GainSegDev{hseg_out_[b].amp0, hseg_out_[b].ratio, hseg_out_[b].target, hseg_out_[b].mode, hseg_out_[b].n_slew};
  }
  // pageable sources: the copies below have read them when they return
  B200_CHECK(cudaMemcpyAsync(seg_in_.p, up_in.data(), sizeof(GainSegDev) * B_, cudaMemcpyHostToDevice, s));
  B200_CHECK(cudaMemcpyAsync(seg_out_.p, up_out.data(), sizeof(GainSegDev) * B_, cudaMemcpyHostToDevice, s));

  // ---- ResampleIn (resample.h:101-112): host rate -> 48 kHz ----
  int n = 0;
  ResampleArgs a;
  std::memset(&a, 0, sizeof(a));
  a.in = in_.as<float>();
  a.in_pitch = kMaxBlock;
  a.out = x48_.as<float>();
  a.out_pitch = cap48_;
  a.q = m;
  a.hist = hist_in_.as<float>();
  a.seg = seg_in_.as<GainSegDev>();
  a.gain_where = 1;
  a.rh = rh_;
  a.rl = rl_;
  a.L = L_;
  if (down_first_) {   // Downsample, resample.h:141-143
    n = (m * rl_ + fc_down_) / rh_;
    a.down = 1;
    a.fc0 = fc_down_;
    a.H = hist_high_;
    a.coef = coef_.as<float>();
    a.gainmul = static_cast<float>(rl_) / static_cast<float>(rh_);
    fc_down_ += m * rl_ - n * rh_;
  } else {             // Upsample, resample.h:181-184
    n = ((m + 1) * rh_ - fc_up_ - 1) / rl_;
    a.down = 0;
    a.fc0 = fc_up_;
    a.H = hist_low_;
    a.coef = coef_.as<float>() + L_;
    fc_up_ += n * rl_ - m * rh_;
  }
  a.N = n;
  anyrate_resample_kernel<<<B_, kThreads, ResampleSmem(a.H, a.q, a.N), s>>>(a);
  B200_CHECK(cudaGetLastError());
  ++*launches;

  // ---- the 480-sample block FIFO around the model hop (resample.h:343-363, :384-394) ----
  for (int pos = 0; pos < n;) {
    const int len = std::min(kFifo - fifo_idx_, n - pos);
    anyrate_fifo_swap_kernel<<<B_, 128, 0, s>>>(x48_.as<float>(), y48_.as<float>(), cap48_, pos, len, fifo_.as<float>(), fifo_idx_);
    B200_CHECK(cudaGetLastError());
    ++*launches;
    fifo_idx_ += len;
    pos += len;
    if (fifo_idx_ == kFifo) {
      fifo_idx_ = 0;
      anyrate_pick_kernel<<<B_, kInHop, 0, s>>>(fifo_.as<float>(), x16_dev);
      B200_CHECK(cudaGetLastError());
      run_hop();
      anyrate_stuff_kernel<<<B_, kOutHop, 0, s>>>(o24_dev, fifo_.as<float>());
      B200_CHECK(cudaGetLastError());
      *launches += 2;
    }
  }

  // ---- ResampleOut (resample.h:113-125): 48 kHz -> host rate, then gain_out ----
  ResampleArgs o;
  std::memset(&o, 0, sizeof(o));
  o.in = y48_.as<float>();
  o.in_pitch = cap48_;
  o.out = out_.as<float>();
  o.out_pitch = kMaxBlock;
  o.q = n;
  o.hist = hist_out_.as<float>();
  o.seg = seg_out_.as<GainSegDev>();
  o.gain_where = 2;
  o.rh = rh_;
  o.rl = rl_;
  o.L = L_;
  int N = 0;
  if (down_first_) {   // Upsample, resample.h:173-180 (the down clock has already advanced)
    N = (n * rh_ + fc_down_ - fc_up_) / rl_;
    o.down = 0;
    o.fc0 = fc_up_;
    o.H = hist_low_;
    o.coef = coef_.as<float>() + L_;
    fc_up_ += N * rl_ - n * rh_;
  } else {             // Downsample
    N = (n * rl_ + fc_down_) / rh_;
    o.down = 1;
    o.fc0 = fc_down_;
    o.H = hist_high_;
    o.coef = coef_.as<float>();
    o.gainmul = static_cast<float>(rl_) / static_cast<float>(rh_);
    fc_down_ += n * rl_ - N * rh_;
  }
  if (N != m) Fail(-120, "any-rate adapter: the resampler clocks lost step (output length != block size)", __FILE__, __LINE__);
  o.N = N;
  anyrate_resample_kernel<<<B_, kThreads, ResampleSmem(o.H, o.q, o.N), s>>>(o);
  B200_CHECK(cudaGetLastError());
  ++*launches;
  B200_CHECK(cudaMemcpy2DAsync(out_host, row, out_.p, sizeof(float) * kMaxBlock, row, B_, cudaMemcpyDeviceToHost, s));
  B200_CHECK(cudaStreamSynchronize(s));
}

void LaunchEchoModel(const float* x16, float* o24, int B, cudaStream_t s) {
  echo_model_kernel<<<B, kOutHop, 0, s>>>(x16, o24);
  B200_CHECK(cudaGetLastError());
}

}  // namespace b200
