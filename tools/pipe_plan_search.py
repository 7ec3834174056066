"""Coordinate search over the depth-2 gating plan (BeatriceB200_SetPipelinePlan): for every encoder kernel, the
vocoder kernel behind which it may start.  One engine, 256 streams, device-resident 48 kHz hops, CUDA-event timing.
   python tools/pipe_plan_search.py [hops=300] [rounds=2]"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beatrice_vst_b200 import batch as bbatch  # noqa: E402
from beatrice_vst_b200 import lib as blib  # noqa: E402
from beatrice_vst_b200 import model_spec, signals  # noqa: E402

HOPS = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ROUNDS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
N = 256
LANE_OPS = ["phone.fe1", "phone.fe2", "phone.fe3", "phone.fe4", "phone.chain",
            "pitch.fe1", "pitch.fe2", "pitch.fe3", "pitch.fe4", "pitch.chain", "pitch.head", "pitch.argmax"]
# vocoder ops a lane kernel can be gated behind (names that do not exist in the program -- the upsamplers of stages the fused
# MRF kernel computes in its prologue -- are ignored by the engine)
GATES = ["wave.cond", "wave.ups0", "wave.mrf0", "wave.ups1", "wave.mrf1", "wave.ups2", "wave.mrf2", "wave.ups3", "wave.mrf3"]
if os.environ.get("BEATRICE_B200_FUSE_UPS", "1") != "0":
    GATES = ["wave.cond", "wave.ups0", "wave.mrf0", "wave.mrf1", "wave.mrf2", "wave.mrf3"]


def main():
    product = blib.load_product()
    with tempfile.TemporaryDirectory() as d:
        model_spec.write_model_dir(d, 8, 2, 0)
        eng = bbatch.Engine(product, N, precision=2)
        assert eng.load(d) == 0
        assert eng.set_pipeline_depth(2) == 0
        bank = 64
        x = np.tile(signals.batch_48k(32, bank, seed0=1), (1, N // 32, 1))
        d_bank = eng.dev_alloc("bank", bank * N * 480)
        d_out = eng.dev_alloc("out", N * 480)
        for h in range(bank):
            eng.dll.BeatriceB200_CopyToDevice(eng.h, d_bank + h * N * 480 * 4, x[h].ctypes.data, x[h].nbytes)
        stream = torch.cuda.ExternalStream(eng.cuda_stream)

        def measure(plan):
            assert eng.set_pipeline_plan(",".join(f"{k}={v}" for k, v in plan.items())) == 0
            for i in range(12):
                eng.process_48k_device(d_bank + (i % bank) * N * 480 * 4, d_out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(HOPS):
                eng.process_48k_device(d_bank + (i % bank) * N * 480 * 4, d_out)
            e1.record(stream)
            eng.synchronize()
            return 1e3 * e0.elapsed_time(e1) / HOPS

        plan = {op: (("wave.ups1" if "wave.ups1" in GATES else "wave.mrf0") if op.endswith("chain") else "wave.mrf0") for op in LANE_OPS}
        best = measure(plan)
        print(f"start {best:.1f} us  {plan}", flush=True)
        for r in range(ROUNDS):
            for op in LANE_OPS:
                for g in GATES:
                    if plan[op] == g:
                        continue
                    trial = dict(plan)
                    trial[op] = g
                    print(f"  try {op} -> {g}", flush=True)
                    t = measure(trial)
                    if t < best - 0.3:
                        best, plan = t, trial
                        print(f"round {r} {op} -> {g}: {best:.1f} us", flush=True)
        print(f"best {best:.1f} us")
        print("PLAN " + ",".join(f"{k}={v}" for k, v in plan.items()))
        # confirm: default vs best, three times each
        for name, p in (("built-in", {}), ("best", plan)):
            print(name, [round(measure(p), 1) for _ in range(3)])
        eng.close()


if __name__ == "__main__":
    main()
