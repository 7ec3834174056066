"""Prints (no asserts) the per-hop RMS of the batched engine against the CPU oracle for a small
batch, under whatever BEATRICE_B200_* developer overrides are set -- used to bisect the fused
MRF kernel on a GPU box.   python tools/mrf_probe.py <precision> <streams> <hops>"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beatrice_vst_b200 import batch as bbatch  # noqa: E402
from beatrice_vst_b200 import lib as blib  # noqa: E402
from beatrice_vst_b200 import model_spec, signals  # noqa: E402


def rms(a, b):
    return float(np.sqrt(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)))


def oracle_stream(oracle, d, x):
    s = blib.SingleStream(oracle, d)
    s.set_pitch_range(1, 383)
    _, _, _, w = s.run(x)
    s.close()
    return w


def main():
    prec = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    hops = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import loader as oracle_loader
    product, oracle = blib.load_product(), oracle_loader.load_oracle()
    with tempfile.TemporaryDirectory() as d:
        model_spec.write_model_dir(d, 8, 2, 0)
        xs = signals.batch_16k(n, hops, seed0=300)
        eng = bbatch.Engine(product, n, precision=prec)
        assert eng.load(d) == 0
        got = np.stack([eng.process_frames(xs[h]).copy() for h in range(hops)], axis=1)
        eng.close()
        tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("BEATRICE_B200_"))
        print(f"[probe] precision {prec} streams {n} hops {hops} {tag}")
        for s in sorted(set([0, 1, n // 2, n - 1])):
            ref = oracle_stream(oracle, d, xs[:, s, :].reshape(-1))
            per_hop = " ".join(f"{rms(got[s][h], ref[h]):.1e}" for h in range(hops))
            print(f"[probe]   stream {s}: total {rms(got[s], ref):.2e}  per hop {per_hop}")


if __name__ == "__main__":
    main()
