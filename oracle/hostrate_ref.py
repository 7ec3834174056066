"""ORACLE (test infrastructure only) -- numpy restatement of what the reference wraps around
the per-frame model call at a 48 kHz host rate with 480-sample blocks:

  ``Gain::Process``                         reference src/common/gain.h:41-71
  ``DownUpSamplerImpl::Reset/Downsample``   src/common/resample.h:209-230, :130-164
  ``ConvertStreamFunctionBlockSize<480>``   src/common/resample.h:331-364
  ``ConvertStreamFunctionFrom2In3OutTo6InOut<80>``  src/common/resample.h:370-395
  ``DownUpSamplerImpl::Upsample``           src/common/resample.h:168-206

Pinned against the real reference code: ``tests/test_callsite.py`` checks it bit-for-bit
against ``oracle/_ref/callsite_runner_oracle`` (the reference's own src/common compiled in
place).  float32 element-wise numpy ops round exactly like the reference's scalar float code
built without FMA contraction, and the tap loop below keeps the reference's summation order.
"""
from __future__ import annotations

import math

import numpy as np

HOP = 480
TAPS = 33


def _sinc(x: float) -> float:
    if abs(x) < 1e-8:
        return 1.0
    return math.sin(x * math.pi) / (x * math.pi)


def fir_tables(sample_rate: float = 48000.0):
    cutoff_down = 0.99 * 16000.0 / min(max(sample_rate, 16000.0), 48000.0)
    cutoff_up = 0.99 * 24000.0 / min(max(sample_rate, 24000.0), 48000.0)
    down = np.empty(TAPS, np.float32)
    up = np.empty(TAPS, np.float32)
    for i in range(TAPS):
        window = 0.5 - 0.5 * math.cos(math.pi * 2.0 / float(TAPS - 1) * float(i))
        down[i] = np.float32(cutoff_down * _sinc(float(i - 16) / 1.0 * cutoff_down) * window)
        up[i] = np.float32(cutoff_up * _sinc(float(i - 16) / 1.0 * cutoff_up) * window)
    return down, up


class GainRef:
    """gain.h:19-72, sample-exact."""

    def __init__(self, sample_rate=48000.0):
        self.sr = sample_rate
        self.target_db = 0.0
        self.current_db = 0.0

    def process(self, x: np.ndarray) -> np.ndarray:
        target = math.pow(10.0, self.target_db * 0.05)
        cur = math.pow(10.0, self.current_db * 0.05)
        amp = np.empty(len(x), np.float64)
        i = 0
        if cur < target:
            ratio = math.pow(10.0, (2.0 / (self.sr * 0.001)) * 0.05)
            while i < len(x) and cur < target:
                cur = min(cur * ratio, target)
                amp[i] = cur
                i += 1
        elif cur > target:
            ratio = math.pow(10.0, (-2.0 / (self.sr * 0.001)) * 0.05)
            while i < len(x) and cur > target:
                cur = max(cur * ratio, target)
                amp[i] = cur
                i += 1
        amp[i:] = cur
        self.current_db = 20.0 * math.log10(cur)
        return (x.astype(np.float64) * amp).astype(np.float32)


class HostRateRef:
    """One stream; ``model(x160) -> y240`` is the per-frame call (Process1)."""

    def __init__(self, model):
        self.model = model
        self.cd, self.cu = fir_tables()
        self.gain_in, self.gain_out = GainRef(), GainRef()
        self.g_hist = np.zeros(32, np.float32)     # gained input history
        self.z_hist = np.zeros(32, np.float32)     # zero-stuffed model output history
        self.fifo = np.zeros(HOP, np.float32)      # ConvertStreamFunctionBlockSize::buffer_

    def process(self, x480: np.ndarray) -> np.ndarray:
        assert len(x480) == HOP
        g = self.gain_in.process(np.asarray(x480, np.float32))
        ext = np.concatenate([self.g_hist, g])
        self.g_hist = ext[-32:].copy()
        n = np.arange(HOP) + 32
        y = np.zeros(HOP, np.float32)
        for m in range(1, TAPS - 1):                       # resample.h:149-157, ratio 1/1
            y = (y + ext[n - m + 1] * self.cd[m]).astype(np.float32)
        previous = self.fifo                               # resample.h:343-363
        x16 = y[2::3].copy()                               # input[(i+1)*3-1], resample.h:384-386
        o24 = np.asarray(self.model(x16), np.float32)
        z = np.zeros(HOP, np.float32)
        z[0::2] = o24                                      # resample.h:390-393
        self.fifo = z
        ext = np.concatenate([self.z_hist, previous])
        self.z_hist = ext[-32:].copy()
        w = np.zeros(HOP, np.float32)
        for i in range(0, TAPS - 1):                       # resample.h:193-200, ratio 1/1
            w = (w + ext[n - i] * self.cu[i]).astype(np.float32)
        return self.gain_out.process(w)
