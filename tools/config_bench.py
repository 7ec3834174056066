"""The two BASELINE.json configurations that are not the headline bench line, measured on one GPU:

  latency  (config 4) batch-1 latency mode: per-frame wall time host-call -> samples-back, p50 / p99,
           (a) through the single-stream beatrice.h ABI exactly as ProcessorCore2::Process1 drives it
           (ExtractPhone1 + EstimatePitch1 + GenerateWaveform1, three synchronous calls per 10 ms frame)
           and (b) through the batched engine with one stream (BeatriceB200_Process48k, host buffers).
  sweep    (config 5) per-stream speaker / pitch-shift / formant sweep, 128 streams per GPU (= 1024 over
           8 GPUs): stream s has speaker s % n, pitch shift -12..+12 st, formant ((s % 9) - 4) / 2, and every
           100 frames each stream advances its speaker (set-speaker + the 4-hop key-value schedule).

   python tools/config_bench.py latency [frames=3000]
   python tools/config_bench.py sweep [frames=1000] [streams=128]
Prints one JSON line per measurement."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beatrice_vst_b200 import batch as bbatch  # noqa: E402
from beatrice_vst_b200 import lib as blib  # noqa: E402
from beatrice_vst_b200 import model_spec, signals  # noqa: E402


def pct(a, p):
    return float(np.percentile(np.asarray(a, np.float64), p))


def latency(frames):
    product = blib.load_product()
    with tempfile.TemporaryDirectory() as d:
        model_spec.write_model_dir(d, 8, 2, 0)
        # (a) the reference call site's three calls per frame
        s = blib.SingleStream(product, d)
        assert s.ok, s.errors
        s.set_pitch_range(1, 383)
        x = signals.voice_like((frames + 50) * 160, 16000.0, seed=1)
        t = []
        for i in range(frames + 50):
            f0 = time.perf_counter_ns()
            s.frame(x[i * 160:(i + 1) * 160])
            t.append((time.perf_counter_ns() - f0) * 1e-3)
        s.close()
        t = t[50:]
        print(json.dumps({"config": "batch-1 latency, beatrice.h ABI (3 calls per frame, host buffers)", "frames": frames,
                          "p50_us": pct(t, 50), "p99_us": pct(t, 99), "max_us": max(t), "mean_us": float(np.mean(t)),
                          "realtime_budget_us": 10000}))
        # (b) one stream through the batched 48 kHz entry
        eng = bbatch.Engine(product, 1, precision=2)
        assert eng.load(d) == 0
        x48 = signals.batch_48k(1, frames + 50, seed0=2)     # [hops][1][480]
        hin, hout = eng.pinned("in", (1, 480)), eng.pinned("out", (1, 480))
        t = []
        for i in range(frames + 50):
            hin[:] = x48[i]
            f0 = time.perf_counter_ns()
            eng.process_48k(hin, hout)
            t.append((time.perf_counter_ns() - f0) * 1e-3)
        eng.close()
        t = t[50:]
        print(json.dumps({"config": "batch-1 latency, BeatriceB200_Process48k (1 call per hop, pinned host buffers, bf16x3)",
                          "frames": frames, "p50_us": pct(t, 50), "p99_us": pct(t, 99), "max_us": max(t),
                          "mean_us": float(np.mean(t)), "realtime_budget_us": 10000}))


def sweep(frames, n):
    product = blib.load_product()
    with tempfile.TemporaryDirectory() as d:
        model_spec.write_model_dir(d, 8, 2, 0)
        eng = bbatch.Engine(product, n, precision=2)
        assert eng.load(d) == 0
        ns = eng.n_speakers
        spk = [s % ns for s in range(n)]
        for s in range(n):
            eng.set("TargetSpeaker", spk[s], s)
            eng.set("PitchShift", float((s % 25) - 12), s)
            eng.set("FormantShift", ((s % 9) - 4) / 2.0, s)
        eng.reset_stream(-1)
        x48 = signals.batch_48k(min(n, 32), 64, seed0=5)
        x48 = np.tile(x48, (1, (n + 31) // 32, 1))[:, :n, :]
        hin, hout = eng.pinned("in", (n, 480)), eng.pinned("out", (n, 480))
        for i in range(20):
            hin[:] = x48[i % 64]
            eng.process_48k(hin, hout)
        t0 = time.perf_counter()
        changes = 0
        for i in range(frames):
            if i and i % 100 == 0:                      # every stream moves to its next speaker
                for s in range(n):
                    spk[s] = (spk[s] + 1) % ns
                    eng.set("TargetSpeaker", spk[s], s)
                changes += n
            hin[:] = x48[i % 64]
            eng.process_48k(hin, hout)
        wall = time.perf_counter() - t0
        finite = bool(np.isfinite(hout).all())
        eng.close()
        print(json.dumps({"config": f"{n} streams, per-stream speaker + pitch-shift + formant sweep, speaker change every 100 frames "
                          "(host buffers, set-speaker calls inside the timed region, bf16x3)", "frames": frames,
                          "frames_per_s": n * frames / wall, "ms_per_hop": 1e3 * wall / frames, "speaker_changes": changes,
                          "output_finite": finite}))


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "latency"
    if mode == "latency":
        latency(int(sys.argv[2]) if len(sys.argv) > 2 else 3000)
    else:
        sweep(int(sys.argv[2]) if len(sys.argv) > 2 else 1000, int(sys.argv[3]) if len(sys.argv) > 3 else 128)
