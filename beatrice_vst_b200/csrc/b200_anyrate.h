// Host-rate adapter on the device for ANY host sample rate and block size: what ProcessorCore2::Process wraps around
// the per-frame model call (reference src/common/processor_core_2.cc:44-46, resample.h:401-438):
//
//   gain_in (gain.h:41-71, slew rate from the host sample rate)
//   -> ResampleIn : host rate -> 48 kHz, rational polyphase FIR with a fraction clock (resample.h:130-206; Downsample
//                   when the host runs at >= 48 kHz, Upsample below), windowed-sinc table of 32 * ratio_high + 1
//                   coefficients (resample.h:209-230), ratio from the Stern-Brocot search of resample.h:25-46
//   -> 480-sample block FIFO: a call's 48 kHz samples swap places with the previously processed block
//                   (resample.h:343-363); every time the FIFO fills, ONE model hop runs on it:
//                   keep samples 3i+2 (48k -> 16k, resample.h:384-386) -> MODEL -> zero-stuff x2 (:390-393)
//   -> ResampleOut: 48 kHz -> host rate (the other direction of the same sampler, cut-off 0.99 * 24 kHz)
//   -> gain_out
//
// All streams of the engine share the host rate and the block size of a call, so the fraction clocks and the FIFO index
// are scalars kept on the host; per stream the device keeps the two sample histories, the FIFO block and nothing else.
// Every floating-point operation is performed in the reference's order with explicit round-to-nearest mul / add (no FMA
// contraction), the tables are computed on the host with the reference's expressions: bit-exact against the reference
// compiled without -march flags (tests: the reference call site over an echo stub library, array_equal).
// The 48 kHz / 480-sample case of b200_hostrate.h is the same chain with ratio 1/1, fused into two kernels per hop and
// graph-captured; this general form is a handful of small launches per call -- it serves hosts at 44.1 / 88.2 / 96 kHz
// and arbitrary block sizes, not the throughput benchmark.
#ifndef BEATRICE_B200_ANYRATE_H_
#define BEATRICE_B200_ANYRATE_H_

#include <functional>
#include <vector>

#include "b200_common.h"
#include "b200_engine.h"

namespace b200 {

class AnyRateState {
 public:
  static constexpr int kMaxBlock = 4096;   // host samples per call
  // false: the reference's resampler would not be ready for this rate (resample.h:243-258)
  // The gain state survives a change of rate (Gain::Context is kept across SetSampleRate, processor_core_2.cc:425-428);
  // the first Init takes it from `seed` (the engine's 48 kHz adapter, which has seen every gain setter so far).
  bool Init(int device, int B, double sample_rate, const class HostRateState* seed);
  double sample_rate() const { return rate_; }
  void SetTargetGain(int b, bool input, double db);
  // One Process call: in_host / out_host [B][m].  run_hop(x16_dev -> o24_dev is implied by the engine): called once per
  // filled FIFO, between PickFrames and StuffFrames.
  void Process(const float* in_host, float* out_host, int m, float* x16_dev, const float* o24_dev,
               const std::function<void()>& run_hop, cudaStream_t s, uint64_t* launches);

 private:
  struct HostGain {
    double target_db = 0.0, current_db = 0.0;
  };
  struct Seg {     // one stream, one call: amplitude recurrence of Gain::Process
    double amp0, ratio, target;
    int mode;      // 0 steady, 1 rising, 2 falling
    int n_slew;    // samples for which the recurrence runs before the amplitude sits on the target
  };
  void StepGain(HostGain* g, Seg* seg, int m) const;
  int device_ = -1, B_ = 0;
  double rate_ = 0.0;
  bool down_first_ = true;
  int rh_ = 1, rl_ = 1, L_ = 0, hist_high_ = 0, hist_low_ = 0;
  int fc_down_ = 0, fc_up_ = 0, fifo_idx_ = 0;
  int cap48_ = 0;   // row pitch of the 48 kHz work buffers
  DeviceBuffer coef_, in_, out_, x48_, y48_, hist_in_, hist_out_, fifo_, seg_in_, seg_out_;
  std::vector<HostGain> gin_, gout_;
  std::vector<Seg> hseg_in_, hseg_out_;
};

// test hook: o24[b][i] = x16[b][i] for i < 160, 0 above (the "model" of oracle/stub_beatricelib.cc in echo mode)
void LaunchEchoModel(const float* x16, float* o24, int B, cudaStream_t s);

}  // namespace b200

#endif  // BEATRICE_B200_ANYRATE_H_
