import os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from beatrice_vst_b200 import batch as bbatch, lib as blib, model_spec, signals
mode = sys.argv[1]
product = blib.load_product()
d = tempfile.mkdtemp(); model_spec.write_model_dir(d, 8, 2, 0)
N = 256
depth = int(os.environ.get("DEPTH", "2"))
eng = bbatch.Engine(product, N, precision=2); assert eng.load(d) == 0; assert eng.set_pipeline_depth(depth) == 0
bank = 64
x = np.tile(signals.batch_48k(32, bank, seed0=1), (1, N // 32, 1))
d_bank = eng.dev_alloc("bank", bank * N * 480); d_out = eng.dev_alloc("out", N * 480)
for h in range(bank):
    eng.dll.BeatriceB200_CopyToDevice(eng.h, d_bank + h * N * 480 * 4, x[h].ctypes.data, x[h].nbytes)
def run(hops):
    for i in range(hops):
        eng.process_48k_device(d_bank + (i % bank) * N * 480 * 4, d_out)
    eng.synchronize()
t0 = time.time()
if mode == "recapture":
    for k in range(int(os.environ.get("RECAP", "150"))):
        if eng.set_pipeline_plan("") != 0:
            print("FAILED at recapture", k, bbatch.last_error(product), time.time() - t0); break
        run(int(os.environ.get("HOPS", "300")))
    print("recapture default plan x150 ok", time.time() - t0, bbatch.last_error(product))
elif mode == "long":
    plan = sys.argv[2] if len(sys.argv) > 2 else ""
    assert eng.set_pipeline_plan(plan) == 0
    for k in range(100):
        run(500)
        if bbatch.last_error(product)[0]:
            print("error at", k, bbatch.last_error(product)); break
    print("long run ok", repr(plan), time.time() - t0, bbatch.last_error(product))
