"""Runs a few hops of a 256-stream engine at the requested precision (for ncu captures and in-kernel timelines).
   python tools/tc_probe.py [precision=1] [streams=256] [hops=3] [upsampler form=0]
form 0: the upsamplers of stages 1-3 as launches of their own (what the depth-2 headline runs), 1: in the prologue of
the fused MRF kernels (the depth-1 latency path)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beatrice_vst_b200 import batch as bbatch  # noqa: E402
from beatrice_vst_b200 import lib as blib  # noqa: E402
from beatrice_vst_b200 import model_spec, signals  # noqa: E402

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
hops = int(sys.argv[3]) if len(sys.argv) > 3 else 3
form = int(sys.argv[4]) if len(sys.argv) > 4 else 0
product = blib.load_product()
with tempfile.TemporaryDirectory() as d:
    model_spec.write_model_dir(d, 8, 2, 0)
    os.environ["BEATRICE_B200_NO_GRAPH"] = "1"
    eng = bbatch.Engine(product, n, precision=prec)
    assert eng.load(d) == 0
    assert eng.set_upsampler_form(form) == 0
    xs = np.tile(signals.batch_16k(8, hops, seed0=5), (1, n // 8, 1))
    for h in range(hops):
        out = eng.process_frames(xs[h])
    print("ok", float(out.std()))
    eng.close()
