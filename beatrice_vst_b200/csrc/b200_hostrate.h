// 48 kHz host-rate adapter on the device: what ProcessorCore2::Process wraps around the
// per-frame model call when the host runs at 48 kHz with 480-sample blocks (reference
// src/common/processor_core_2.cc:44-46):
//
//   gain_in (gain.h:41-71)
//   -> FIR low-pass, 33-coefficient Hann-windowed sinc, cutoff 0.33  (resample.h:130-164, :209-230)
//   -> 480-sample block FIFO: the block handed back is the PREVIOUS processed block
//      (resample.h:343-363)
//   -> keep samples 3i+2 (48k -> 16k, resample.h:384-386) -> MODEL -> zero-stuff x2 (:390-393)
//   -> FIR low-pass, cutoff 0.495 (resample.h:168-206)
//   -> gain_out
//
// Every floating-point operation is performed in the reference's order with explicit
// round-to-nearest mul/add (no FMA contraction), so the adapter is bit-exact against the
// reference code compiled without -march flags; the filter tables and the dB<->amplitude
// conversions are computed on the host with the same libm expressions.
#ifndef BEATRICE_B200_HOSTRATE_H_
#define BEATRICE_B200_HOSTRATE_H_

#include <vector>

#include "b200_common.h"
#include "b200_engine.h"

namespace b200 {

struct GainSeg {  // one stream, one hop: amplitude recurrence of Gain::Process
  double amp0;    // amplitude entering the hop (DbToAmp(current_gain_db))
  double ratio;   // per-sample factor while slewing
  double target;  // DbToAmp(target_gain_db)
  int mode;       // 0 steady, 1 rising (min(a*ratio,target)), 2 falling (max(a*ratio,target))
  int pad;
};

class HostRateState {
 public:
  void Init(int device, int B);
  void SetTargetGain(int b, bool input, double db);
  // Gain::Context state of stream b (target / current gain in dB): the any-rate adapter starts from it
  void GetGain(int b, bool input, double* target_db, double* current_db) const {
    const HostGain& g = input ? gin_[b] : gout_[b];
    *target_db = g.target_db;
    *current_db = g.current_db;
  }
  void ResetStream(int b, cudaStream_t s);
  // host side of one hop: advances the per-stream gain state like Gain::Process does and
  // uploads the segments if they changed.  Call before EnqueueIn, outside graph capture.
  // out_lag (pipeline depth 2 of the engine): the block handed back by this call is the one depth 1 hands back one
  // call earlier, so the OUTPUT gain segment uploaded is the one computed for the previous call.
  void PrepareHop(cudaStream_t s, bool out_lag = false);
  // in48 (device, [B][480]) -> x16 (device, [B][160]); graph-capturable
  void EnqueueIn(float* x16, cudaStream_t s);
  // model output o24 (device, [B][240]) -> out48 (device, [B][480]); graph-capturable
  void EnqueueOut(const float* o24, cudaStream_t s);
  // The same in two halves.  The block a hop hands back depends on the model outputs of the two PREVIOUS hops
  // only (the block FIFO), so EnqueueOutEarly may run -- on another stream, outside the hop graph: the hop index
  // is a launch argument -- before this hop's model call, and EnqueueStore after it.
  void EnqueueOutEarly(cudaStream_t s);
  void EnqueueStore(const float* o24, cudaStream_t s);   // graph-capturable
  void HopDone();   // host mirror of the device hop counter: call once per enqueued hop
  float* in48() const { return in48_.as<float>(); }
  float* out48() const { return out48_.as<float>(); }
  static constexpr int kKernelsPerHop = 2;

 private:
  struct HostGain {
    double target_db = 0.0, current_db = 0.0;
    bool settled = false;
  };
  int device_ = -1, B_ = 0;
  DeviceBuffer in48_, out48_, g_ring_, o_ring_, frame_, done_, coef_, seg_in_, seg_out_;
  int host_frame_ = 0;
  std::vector<HostGain> gin_, gout_;
  std::vector<GainSeg> hseg_in_, hseg_out_, up_in_, up_out_, lag_out_;
  bool uploaded_ = false;
};

}  // namespace b200

#endif  // BEATRICE_B200_HOSTRATE_H_
