"""Synthetic voice-like test signals.

Harmonic source ``0.3 * sum_{h=1..8} sin(2 pi h int f0) / h`` with f0 gliding
110 -> 330 Hz (log sweep over the signal), plus ``0.01 * N(0,1)``, hard clipped
to [-1, 1].  Never all-zero: the VST path short-cuts silent blocks before they
reach the library (reference ``src/vst/processor.cc:195-214``).
"""
from __future__ import annotations

import numpy as np


def voice_like(n_samples: int, sample_rate: float = 48000.0, seed: int = 0,
               f0_scale: float = 1.0, sweep_seconds: float = 10.0) -> np.ndarray:
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    f0 = 110.0 * f0_scale * np.power(3.0, np.minimum(t / sweep_seconds, 1.0))
    phase = 2.0 * np.pi * np.cumsum(f0) / sample_rate
    x = np.zeros(n_samples, np.float64)
    for h in range(1, 9):
        x += np.sin(h * phase) / h
    rng = np.random.default_rng(seed)
    x = 0.3 * x + 0.01 * rng.standard_normal(n_samples)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def stream_f0_scale(stream_id: int) -> float:
    """Config 2: f0 range scaled per stream by 2^((id%25-12)/12)."""
    return float(2.0 ** (((stream_id % 25) - 12) / 12.0))


def batch_16k(n_streams: int, n_frames: int, seed0: int = 0) -> np.ndarray:
    """[n_frames, n_streams, 160] model-rate frames for the batched API."""
    out = np.empty((n_frames, n_streams, 160), np.float32)
    for s in range(n_streams):
        x = voice_like(n_frames * 160, 16000.0, seed0 + s, stream_f0_scale(s))
        out[:, s, :] = x.reshape(n_frames, 160)
    return out


def batch_48k(n_streams: int, n_frames: int, seed0: int = 0) -> np.ndarray:
    """[n_frames, n_streams, 480] host-rate hops (48 kHz, 10 ms)."""
    out = np.empty((n_frames, n_streams, 480), np.float32)
    for s in range(n_streams):
        x = voice_like(n_frames * 480, 48000.0, seed0 + s, stream_f0_scale(s))
        out[:, s, :] = x.reshape(n_frames, 480)
    return out
