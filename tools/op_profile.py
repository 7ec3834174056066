"""Per-op device time of one hop of the batched engine (BeatriceB200_ProfileHop: CUDA events around
every launch, ops run serially), under whatever BEATRICE_B200_* developer overrides are set.
   python tools/op_profile.py [precision=2] [streams=256] [repeats=8] [pipeline depth=1]
(the depth selects the upsampler form: in the fused MRF kernels' prologue at depth 1, launches of their own at depth 2)"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beatrice_vst_b200 import batch as bbatch  # noqa: E402
from beatrice_vst_b200 import lib as blib  # noqa: E402
from beatrice_vst_b200 import model_spec, signals  # noqa: E402


def _peaks():
    import json
    try:
        m = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(m["bf16_tflops_sustained"]), float(m["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        return 1357.1, 6539.2


PEAK_TF, PEAK_GBS = _peaks()


def main():
    prec = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    depth = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    product = blib.load_product()
    with tempfile.TemporaryDirectory() as d:
        model_spec.write_model_dir(d, 8, 2, 0)
        eng = bbatch.Engine(product, n, precision=prec)
        assert eng.load(d) == 0
        d_in = eng.dev_alloc("in16", n * 160)
        d_out = eng.dev_alloc("out24", n * 240)
        eng.to_device(d_in, signals.batch_16k(min(n, 32), 1, seed0=3)[0].repeat((n + 31) // 32, axis=0)[:n])
        for _ in range(3):
            eng.process_frames_device(d_in, d_out)
        eng.synchronize()
        assert eng.set_pipeline_depth(depth) == 0     # after the warm-up: ProfileHop wants an empty pipeline
        allr = [eng.profile_hop(d_in, d_out) for _ in range(reps)]
        eng.close()
    tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("BEATRICE_B200_"))
    print(f"[ops] precision {prec} streams {n} pipeline depth {depth} {tag}")
    tot = 0.0
    for i, r in enumerate(allr[-1]):
        us = 1e3 * float(np.median([a[i]["ms"] for a in allr[1:]]))
        tot += us
        tf = r["flops"] / (us * 1e-6) / 1e12 if us > 0 else 0.0
        gbs = r["bytes"] / (us * 1e-6) / 1e9 if us > 0 else 0.0
        # algorithmic work of the op (weights + activations in and out, once) against the two measured ceilings
        print(f"[ops] {r['name']:<28s} {us:8.1f} us  {r['flops'] / 1e9:7.3f} GFLOP  {tf:7.1f} TF/s ({100 * tf / PEAK_TF:4.1f} %)  "
              f"{r['bytes'] / 1e6:7.2f} MB  {gbs:7.1f} GB/s ({100 * gbs / PEAK_GBS:4.1f} %)")
    print(f"[ops] serial sum {tot:.1f} us")


if __name__ == "__main__":
    main()
