"""Randomised soak of the any-rate device adapter (BeatriceB200_SetHostSampleRate / _ProcessAnyRate, echo model) against the
REFERENCE call site over the echo stub (oracle/_ref/callsite_runner_stub): random host rates, block sizes and gain events,
bit-exact over the whole run.
   python tools/soak_anyrate.py [runs=24] [seed=1]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import callsite  # noqa: E402
from beatrice_vst_b200 import batch as bbatch  # noqa: E402
from beatrice_vst_b200 import lib as blib  # noqa: E402
from beatrice_vst_b200 import model_spec, signals  # noqa: E402

RATES = [8000.0, 11025.0, 12000.0, 16000.0, 22050.0, 24000.0, 32000.0, 37800.0, 44100.0, 47999.0, 48000.0, 48001.0, 50000.0, 64000.0,
         88200.0, 96000.0, 176400.0, 192000.0]


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    product = blib.load_product()
    n = 2
    with tempfile.TemporaryDirectory() as d:
        model_spec.write_model_dir(d, 8, 2, 0)
        toml = os.path.join(d, "model.toml")
        for r in range(runs):
            rate = float(rng.choice(RATES))
            block = int(rng.choice([1, 7, 31, 64, 100, 128, 256, 441, 480, 512, 999, 1024, 2048, 4096]))
            samples = max(int(rate * 0.25) // block, 3) * block
            hops48 = (samples + 479) // 480
            x = np.ascontiguousarray(signals.batch_48k(n, hops48, seed0=7000 + 10 * seed + r).transpose(1, 0, 2).reshape(n, -1)[:, :samples])
            nb = samples // block
            events = [(int(rng.integers(0, nb)), str(rng.choice(["input_gain", "output_gain"])), float(rng.integers(-18, 9)))
                      for _ in range(int(rng.integers(0, 6)))]
            events.sort()
            eng = bbatch.Engine(product, n)
            assert eng.load(d) == 0
            rc = eng.set_host_sample_rate(rate)
            if rc != 0:
                print(f"run {r}: rate {rate:.0f} not supported by the device adapter (rc {rc})")
                eng.close()
                continue
            assert eng.set_echo_model(True) == 0
            got = np.zeros_like(x)
            for bi in range(nb):
                for (at, name, db) in events:
                    if at == bi:
                        assert eng.set("InputGain" if name == "input_gain" else "OutputGain", db, -1) == 0
                got[:, bi * block:(bi + 1) * block] = eng.process_any_rate(x[:, bi * block:(bi + 1) * block])
            eng.close()
            for s in range(n):
                want, info = callsite.run("stub", toml, x[s], rate, block, events=events, echo=True)
                assert info["load"] == 0 and info["last"] == 0, info
                if not np.array_equal(got[s], want):
                    k = int(np.argmax(got[s] != want))
                    print(f"MISMATCH run {r}: rate {rate:.0f} block {block} stream {s} at sample {k}: {got[s][k]!r} vs {want[k]!r}; events {events}")
                    raise SystemExit(1)
            print(f"run {r}: rate {rate:.0f} block {block} samples {samples} events {len(events)}: bit-exact (std {float(got.std()):.3f})", flush=True)
    print("any-rate soak ok")


if __name__ == "__main__":
    main()
