"""Prints (no asserts) how far every output of the CUDA library is from the CPU oracle --
first thing to run on a GPU box after a kernel change."""
import ctypes as C
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


def rms(a, b):
    return float(np.sqrt(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)))


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import loader as oracle_loader
    product, oracle = blib.load_product(), oracle_loader.load_oracle()
    print("devices:", bbatch.device_count(product))
    with tempfile.TemporaryDirectory() as d:
        model_spec.write_model_dir(d, 8, 2, 0)
        x = signals.voice_like(160 * 8, 16000.0, seed=21)
        a = blib.SingleStream(product, d, speaker=3, formant_index=6)
        b = blib.SingleStream(oracle, d, speaker=3, formant_index=6)
        print("load errors", a.errors, b.errors)
        product.dll.BeatriceB200_WaveformTap.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int]
        oracle.dll.BeatriceOracle_WaveformTap.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int]
        for i in range(8):
            xi = x[i * 160:(i + 1) * 160]
            pa, qa, fa, wa = a.frame(xi)
            pb, qb, fb, wb = b.frame(xi, q_override=None)
            line = f"hop {i}: phone {rms(pa, pb):.2e} q {qa}/{qb} feat {rms(fa, fb):.2e} wave {rms(wa, wb):.2e}"
            for which, n in enumerate([256, 256, 640, 1280, 2560, 3840]):
                ta, tb = np.zeros(n, np.float32), np.zeros(n, np.float32)
                fp = lambda v: v.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
                product.dll.BeatriceB200_WaveformTap(a.wc, which, fp(ta), n)
                oracle.dll.BeatriceOracle_WaveformTap(b.wc, which, fp(tb), n)
                line += f" tap{which} {rms(ta, tb):.1e}"
            print(line)
        t = time.time()
        for i in range(50):
            a.frame(x[:160])
        print("single-stream ABI: %.1f us/hop" % ((time.time() - t) / 50 * 1e6))
        import itertools
        for n, prec in itertools.product((4, 256), (0, 1, 2)):
            print(f"---- batch {n} precision {prec} (0 f32, 1 bf16 tcgen05, 2 split-bf16 tcgen05) ----", flush=True)
            eng = bbatch.Engine(product, n, precision=prec)
            print("batch load", eng.load(d))
            xs = signals.batch_16k(min(n, 8), 3, seed0=5)
            xs = np.tile(xs, (1, n // min(n, 8), 1))
            for h in range(3):
                out = eng.process_frames(xs[h])
            o = blib.SingleStream(oracle, d)
            o.set_pitch_range(1, 383)
            _, _, _, w = o.run(xs[:, 1, :].reshape(-1))
            print(f"batch {n}: stream1 last-hop rms vs oracle {rms(out[1], w[-1]):.2e}  (signal rms {float(np.sqrt(np.mean(w[-1] ** 2))):.3f})", flush=True)
            t = time.time()
            for i in range(20):
                eng.process_frames(xs[0])
            dt = (time.time() - t) / 20
            print(f"batch {n}: {dt * 1e6:.0f} us/hop -> {n / dt:.0f} frames/s (host buffers)")
            din, dout = eng.dev_alloc("in", n * 160), eng.dev_alloc("out", n * 240)
            eng.to_device(din, xs[0])
            recs = eng.profile_hop(din, dout)
            tot = sum(r["ms"] for r in recs)
            print(f"profile: {len(recs)} kernels, {tot * 1e3:.0f} us total")
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"diag_b{n}_p{prec}.txt"), "w") as fh:
                for r in recs:
                    fh.write(f"{r['name']:28s} {r['ms'] * 1e3:8.1f} us {r['flops'] / 1e6:10.1f} MFLOP\n")
            for r in sorted(recs, key=lambda r: -r["ms"])[:12]:
                print(f"   {r['name']:28s} {r['ms'] * 1e3:8.1f} us  {r['flops'] / max(r['ms'], 1e-9) / 1e9:8.2f} TFLOP/s")
            eng.close()


if __name__ == "__main__":
    main()
