"""Does splitting the 256 streams into G independent engines (own CUDA stream + hop graph each) that run
their hops concurrently raise frames/s?  Wall-clock over K hops, device-resident input.
   python tools/split_probe.py [streams=256] [hops=300]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beatrice_vst_b200 import batch as bbatch  # noqa: E402
from beatrice_vst_b200 import lib as blib  # noqa: E402
from beatrice_vst_b200 import model_spec, signals  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    hops = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    product = blib.load_product()
    with tempfile.TemporaryDirectory() as d:
        model_spec.write_model_dir(d, 8, 2, 0)
        for G in (1, 2, 3, 4, 8):
            per = n // G
            engs = []
            for g in range(G):
                eng = bbatch.Engine(product, per, precision=2)
                assert eng.load(d) == 0
                d_in = eng.dev_alloc("in16", per * 160)
                d_out = eng.dev_alloc("out24", per * 240)
                eng.to_device(d_in, signals.batch_16k(min(per, 32), 1, seed0=3 + g)[0].repeat((per + 31) // 32, axis=0)[:per])
                engs.append((eng, d_in, d_out))
            for _ in range(20):
                for eng, a, b in engs:
                    eng.process_frames_device(a, b)
            for eng, _, _ in engs:
                eng.synchronize()
            t0 = time.perf_counter()
            for _ in range(hops):
                for eng, a, b in engs:
                    eng.process_frames_device(a, b)
            t_issue = time.perf_counter() - t0
            for eng, _, _ in engs:
                eng.synchronize()
            dt = time.perf_counter() - t0
            print(f"[split] G={G} x {per} streams: {1e6 * dt / hops:8.1f} us/hop  {per * G * hops / dt:10.0f} frames/s  (issue {1e6 * t_issue / hops:.1f} us/hop)")
            for eng, _, _ in engs:
                eng.close()


if __name__ == "__main__":
    main()
