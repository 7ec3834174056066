"""Randomised differential soak: the same stream of hops and parameter events through two engines -- pipeline depth 1 and
depth 2 (same upsampler form) -- must give bit-identical samples, one call apart, for thousands of hops.
Events: speaker changes (with their four-hop key-value schedules), pitch / formant shifts, gains, kNN-VQ on / off, pitch
correction, single-stream resets, morphing slots with new weights.
   python tools/soak_diff.py [hops=2000] [streams=24] [seed=1] [entry=host48|frames]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beatrice_vst_b200 import batch as bbatch  # noqa: E402
from beatrice_vst_b200 import lib as blib  # noqa: E402
from beatrice_vst_b200 import model_spec, signals  # noqa: E402


def soak(hops=2000, n=24, seed=1, entry="host48", product=None):
    """Returns the number of events applied; raises SystemExit(1) at the first sample that differs."""
    rng = np.random.default_rng(seed)
    product = product or blib.load_product()
    with tempfile.TemporaryDirectory() as d:
        model_spec.write_model_dir(d, 8, 2, 0)
        a, b = bbatch.Engine(product, n), bbatch.Engine(product, n)
        for e, depth in ((a, 1), (b, 2)):
            assert e.load(d) == 0
            assert e.set_upsampler_form(0) == 0
            assert e.seed_morph_lottery(7) == 0
            assert e.set_pipeline_depth(depth) == 0
        base = signals.batch_48k(n, 64, seed0=11) if entry == "host48" else signals.batch_16k(n, 64, seed0=11)
        prev_a = None
        n_events = 0
        log = [[] for _ in range(n)]
        for h in range(hops):
            for _ in range(rng.poisson(0.6)):
                s = int(rng.integers(0, n))
                kind = int(rng.integers(0, 9))
                if kind == 0:
                    ev = ("TargetSpeaker", int(rng.integers(0, 8)))
                elif kind == 1:
                    ev = ("PitchShift", float(rng.integers(-12, 13)))
                elif kind == 2:
                    ev = ("FormantShift", float(rng.integers(-4, 5)) / 2.0)
                elif kind == 3:
                    ev = ("InputGain", float(rng.integers(-12, 7))) if entry == "host48" else ("PitchShift", 0.0)
                elif kind == 4:
                    ev = ("OutputGain", float(rng.integers(-12, 7))) if entry == "host48" else ("FormantShift", 0.0)
                elif kind == 5:
                    ev = ("VQNumNeighbors", int(rng.choice([0, 0, 2, 4, 8])))
                elif kind == 6:
                    ev = ("PitchCorrection", float(rng.choice([0.0, 0.3, 1.0])))
                elif kind == 7:
                    ev = ("reset", 0)
                else:
                    w = rng.random(8).astype(np.float32)
                    w[rng.integers(0, 8, 3)] = 0.0
                    for e in (a, b):
                        assert e.set_morph_weights(w, s) == 0
                    ev = ("TargetSpeaker", 8)     # the morphing slot
                for e in (a, b):
                    assert (e.reset_stream(s) if ev[0] == "reset" else e.set(ev[0], ev[1], s)) == 0
                n_events += 1
                log[s].append((h, ev[0], ev[1]))
            x = base[h % 64] * np.float32(0.8 + 0.4 * ((h * 7) % 10) / 10.0)
            ya = (a.process_48k(x) if entry == "host48" else a.process_frames(x)).copy()
            yb = (b.process_48k(x) if entry == "host48" else b.process_frames(x)).copy()
            if prev_a is not None and not np.array_equal(yb, prev_a):
                bad = np.argwhere(yb != prev_a)
                print(f"MISMATCH at hop {h}: {len(bad)} samples, first stream {bad[0][0]} sample {bad[0][1]}; events so far {n_events}")
                for st in sorted(set(int(r[0]) for r in bad)):
                    print(f"  stream {st}: rms {float(np.sqrt(np.mean((yb[st] - prev_a[st]) ** 2))):.3e}  last events {log[st][-8:]}")
                raise SystemExit(1)
            if h == 0:
                assert not yb.any()
            prev_a = ya
        last = b.drain(model_rate=(entry != "host48"))
        assert np.array_equal(last, prev_a), "drain"
        print(f"soak ok: {hops} hops x {n} streams, {n_events} events, entry {entry}, signal std {float(prev_a.std()):.3f}; "
              f"errors {bbatch.last_error(product)}")
        a.close()
        b.close()
    return n_events


if __name__ == "__main__":
    soak(int(sys.argv[1]) if len(sys.argv) > 1 else 2000, int(sys.argv[2]) if len(sys.argv) > 2 else 24,
         int(sys.argv[3]) if len(sys.argv) > 3 else 1, sys.argv[4] if len(sys.argv) > 4 else "host48")
