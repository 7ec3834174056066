"""Runs a few 48 kHz hops of a 256-stream engine through the hop GRAPH at a given pipeline depth -- with
BEATRICE_B200_MRF_TRACE=-1 every fused-MRF CTA prints its residency window (summarise with tools/cta_windows.py).
   BEATRICE_B200_MRF_TRACE=-1 python tools/timeline_probe.py <depth=2> <hops=6> [streams=256]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beatrice_vst_b200 import batch as bbatch  # noqa: E402
from beatrice_vst_b200 import lib as blib  # noqa: E402
from beatrice_vst_b200 import model_spec, signals  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
hops = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n = int(sys.argv[3]) if len(sys.argv) > 3 else 256
product = blib.load_product()
with tempfile.TemporaryDirectory() as d:
    model_spec.write_model_dir(d, 8, 2, 0)
    eng = bbatch.Engine(product, n, precision=2)
    assert eng.load(d) == 0
    assert eng.set_pipeline_depth(depth) == 0
    xs = np.tile(signals.batch_48k(8, hops, seed0=5), (1, n // 8, 1))
    for h in range(hops):
        out = eng.process_48k(xs[h])
    print("ok depth", depth, float(out.std()))
    eng.close()
