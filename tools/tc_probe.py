"""Runs a few hops of a 256-stream engine at the requested precision (for ncu captures)."""
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
product = blib.load_product()
with tempfile.TemporaryDirectory() as d:
    model_spec.write_model_dir(d, 8, 2, 0)
    os.environ["BEATRICE_B200_NO_GRAPH"] = "1"
    eng = bbatch.Engine(product, n, precision=prec)
    assert eng.load(d) == 0
    xs = np.tile(signals.batch_16k(8, hops, seed0=5), (1, n // 8, 1))
    for h in range(hops):
        out = eng.process_frames(xs[h])
    print("ok", float(out.std()))
    eng.close()
