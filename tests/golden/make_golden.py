"""Regenerates the committed golden vectors from the CPU oracle (run in the build container).

  python tests/golden/make_golden.py

* ``m0_family{0,2}.npz`` -- 12 hops of the config-1 test signal through the per-frame ABI
  (speaker 1, formant index 5; rc0 additionally with VQ off/on) computed by
  ``oracle/libbeatrice_oracle.so`` on the seed-0 synthetic model, after the oracle was checked
  against the independent PyTorch model (tests/test_oracle_cpu.py does that check again).
* ``callsite_48k.npz``  -- 20 hops of 48 kHz audio through the REFERENCE's own call site
  (oracle/_ref/callsite_runner_oracle = /root/reference/src/common compiled in place, linked
  to the oracle) with a gain slew, a pitch shift and a speaker change; this one needs
  /root/reference at generation time, which is why the output is committed.
* ``beatrice_h_symbols.txt`` -- the names of every function the reference header
  lib/beatricelib/beatrice.h declares (names only), used by the ABI-export test.
"""
import json
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
HERE = os.path.dirname(os.path.abspath(__file__))

from beatrice_vst_b200 import lib, model_spec, signals  # noqa: E402


def main():
    import loader as oracle_loader
    oracle = oracle_loader.load_oracle()
    with tempfile.TemporaryDirectory() as d:
        for fam in (0, 2):
            md = os.path.join(d, f"f{fam}")
            model_spec.write_model_dir(md, n_speakers=8, family=fam, seed=0)
            x = signals.voice_like(12 * 160, 16000.0, seed=0)
            s = lib.SingleStream(oracle, md, family=fam, speaker=1, formant_index=5)
            assert s.ok
            s.set_pitch_range(1, 383)
            phone, q, feat, wave = s.run(x)
            out = dict(x=x, phone=phone, q=q, feat=feat, wave=wave)
            s.close()
            if fam == 2:
                s = lib.SingleStream(oracle, md, family=2, speaker=1, formant_index=5)
                s.set_pitch_range(1, 383)
                s.set_vq(4)
                p2, q2, f2, w2 = s.run(x)
                out.update(phone_vq4=p2, wave_vq4=w2)
                s.close()
            np.savez_compressed(os.path.join(HERE, f"m0_family{fam}.npz"), **out)
        import callsite
        if callsite.available("oracle"):
            md = os.path.join(d, "f2")
            x48 = signals.voice_like(20 * 480, 48000.0, seed=7)
            events = [(-1, "input_gain", -3.0), (-1, "pitch_shift", 5.0), (6, "voice", 3), (9, "output_gain", 2.0),
                      (12, "formant_shift", -1.0), (14, "pitch_correction", 0.5)]
            y, info = callsite.run("oracle", os.path.join(md, "model.toml"), x48, events=events)
            assert info["load"] == 0 and info["last"] == 0
            np.savez_compressed(os.path.join(HERE, "callsite_48k.npz"), x=x48, y=y,
                                events=np.array(json.dumps(events)))
    hdr = "/root/reference/lib/beatricelib/beatrice.h"
    if os.path.exists(hdr):
        names = sorted(set(re.findall(r"\b(Beatrice20(?:a2|b1|rc0)_[A-Za-z0-9]+)\s*\(", open(hdr).read())))
        open(os.path.join(HERE, "beatrice_h_symbols.txt"), "w").write("\n".join(names) + "\n")
        print(len(names), "symbols")


if __name__ == "__main__":
    main()
