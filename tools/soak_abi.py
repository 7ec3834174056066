"""Randomised soak of the DROP-IN: the reference's unmodified call site (ProcessorProxy -> ProcessorCore2::Process, oracle/_ref)
linked against the CUDA library behind beatrice.h vs the same call site over the CPU oracle, one stream per run, a random stream
of parameter events at random blocks (the call site's own morphing, key-value schedule, pitch transform and resampler run in both;
what differs is only the library behind the 77 symbols).  Any host rate / block size.
   python tools/soak_abi.py [hops=300] [runs=6] [seed=1] [rate=48000] [block=480]"""
import os
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import callsite  # noqa: E402
from beatrice_vst_b200 import model_spec, signals  # noqa: E402


def plan(rng, blocks):
    ev, morphing, vq = [], False, 0
    for b in range(blocks):
        for _ in range(rng.poisson(0.08)):
            kind = int(rng.integers(0, 11))
            if kind == 0:
                ev.append((b, "voice", int(rng.integers(0, 8))))
                morphing = False
            elif kind == 1:
                ev.append((b, "pitch_shift", float(rng.integers(-24, 25)) / 2.0))
            elif kind == 2:
                ev.append((b, "formant_shift", float(rng.integers(-4, 5)) / 2.0))
            elif kind == 3:
                ev.append((b, "input_gain", float(rng.integers(-12, 7))))
            elif kind == 4:
                ev.append((b, "output_gain", float(rng.integers(-12, 7))))
            elif kind == 5 and not morphing:
                vq = int(rng.choice([0, 2, 4, 8]))
                ev.append((b, "vq_num_neighbors", float(vq)))
            elif kind == 6:
                ev += [(b, "pitch_correction", float(rng.choice([0.0, 0.3, 0.7, 1.0]))), (b, "pitch_correction_type", int(rng.integers(0, 2)))]
            elif kind == 7:
                ev.append((b, "reset", 1))
            elif kind == 8:
                ev += [(b, "intonation_intensity", float(rng.choice([0.5, 1.0, 1.5]))), (b, "average_source_pitch", float(rng.integers(48, 66)))]
            elif kind == 9:
                lo = float(rng.integers(30, 45))
                ev += [(b, "min_source_pitch", lo), (b, "max_source_pitch", lo + float(rng.integers(20, 45)))]
            elif kind == 10:
                w = rng.random(8)
                w[rng.integers(0, 8, 3)] = 0.0
                if vq:
                    ev.append((b, "vq_num_neighbors", 0.0))     # the codebook lottery is seeded from std::random_device
                    vq = 0
                ev += [(b, f"morphw{k}", float(w[k])) for k in range(8)] + [(b, "morph_apply", 1), (b, "voice", 8)]
                morphing = True
    return ev


def main():
    hops = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    rate = float(sys.argv[4]) if len(sys.argv) > 4 else 48000.0
    block = int(sys.argv[5]) if len(sys.argv) > 5 else 480
    rng = np.random.default_rng(seed)
    with tempfile.TemporaryDirectory() as d:
        model_spec.write_model_dir(d, 8, 2, 0)
        toml = os.path.join(d, "model.toml")
        samples = int(hops * rate / 100.0) // block * block
        xs = signals.batch_48k(runs, (samples + 479) // 480, seed0=900 + seed).transpose(1, 0, 2).reshape(runs, -1)[:, :samples]
        plans = [plan(rng, samples // block) for _ in range(runs)]

        def one(r):
            ya, ia = callsite.run("b200", toml, xs[r], rate, block, events=plans[r])
            yb, ib = callsite.run("oracle", toml, xs[r], rate, block, events=plans[r])
            assert ia == ib and ia["load"] == 0, (ia, ib)
            return float(np.sqrt(np.mean((ya.astype(np.float64) - yb) ** 2))), float(yb.std()), len(plans[r])

        with ThreadPoolExecutor(max_workers=min(runs, 4)) as pool:
            res = list(pool.map(one, range(runs)))
    worst = max(r[0] for r in res)
    print(f"drop-in soak: {runs} runs x {hops} hops at {rate:.0f} Hz / {block}, events {[r[2] for r in res]}: worst RMS {worst:.2e}; "
          f"per run {[f'{r[0]:.1e}' for r in res]} (signal {min(r[1] for r in res):.3f})")
    raise SystemExit(0 if worst <= 1e-4 else 1)


if __name__ == "__main__":
    main()
