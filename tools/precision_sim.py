"""CPU simulation of cheaper tensor-core arithmetic for the MRF convs, per vocoder stage.

Runs the independent PyTorch statement of spec M0 (oracle/torch_model.py) on a few seconds of synthetic input with
the MRF convolutions of chosen stages replaced by an operand-rounding model of a candidate MMA scheme, fp32
accumulation, everything else exact fp32, and prints the output RMS against the exact model.  Schemes:
  x3    bf16 hi+lo activations x bf16 hi+lo weights, hi*hi + hi*lo + lo*hi     (3 products; what the library ships)
  f16a  fp16 activations (ONE plane) x fp16 hi+lo weights                        (2 products, half the A-panel reads)
  f16   fp16 activations x fp16 weights                                          (1 product)
  bf16  bf16 x bf16                                                              (1 product)
Usage: python tools/precision_sim.py [hops=60] [streams=3]
"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch_model as tm  # noqa: E402
from beatrice_vst_b200 import model_spec, signals  # noqa: E402

SCHEME = ["x3", "x3", "x3", "x3"]   # per stage, read by the patched conv
STAGE = [0]


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def h16(x):
    return x.to(torch.float16).to(torch.float32)


def conv_q(x, w, b, dil, scheme):
    k = w.shape[2]
    xp = F.pad(x, ((k - 1) * dil, 0))
    c = lambda a, ww: F.conv1d(a, ww, None, dilation=dil)  # noqa: E731
    if scheme == "exact":
        y = c(xp, w)
    elif scheme == "x3":
        xh, wh = bf(xp), bf(w)
        xl, wl = bf(xp - xh), bf(w - wh)
        y = c(xh, wh) + c(xh, wl) + c(xl, wh)
    elif scheme == "f16a":
        xh, wh = h16(xp), h16(w)
        wl = h16(w - wh)
        y = c(xh, wh) + c(xh, wl)
    elif scheme == "f16":
        y = c(h16(xp), h16(w))
    elif scheme == "bf16":
        y = c(bf(xp), bf(w))
    else:
        raise ValueError(scheme)
    return y + b[None, :, None]


def wavegen_q(self, phone, q, feat, spk_add, formant_add, film, schemes):
    h = phone @ self.embed[0][:, :, 0].T + self.embed[1]
    h = h + self.pitch_emb[q] + feat @ self.feat_proj + spk_add + formant_add
    x = tm._causal_conv(h.T[None], *self.pre)
    for s in range(4):
        wt, b, rate = self.ups[s]
        n_in = x.shape[2]
        x = F.conv_transpose1d(F.leaky_relu(x, 0.1), wt, b, stride=rate)[:, :, :n_in * rate]
        c = tm.STAGE_CH[s + 1]
        x = x * (1.0 + film[s][:c])[None, :, None] + film[s][c:][None, :, None]
        total = None
        for layers in self.mrf[s]:
            y = x
            for (w1, b1), (w2, b2), d in layers:
                a = conv_q(F.leaky_relu(y, 0.1), w1, b1, d, schemes[s])
                y = y + conv_q(F.leaky_relu(a, 0.1), w2, b2, 1, schemes[s])
            total = y if total is None else total + y
        x = total * np.float32(1.0 / 3.0)
    return torch.tanh(tm._causal_conv(F.leaky_relu(x, 0.1), *self.post)).reshape(-1)


@torch.no_grad()
def main():
    hops = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    torch.set_num_threads(8)
    with tempfile.TemporaryDirectory() as d:
        model_spec.write_model_dir(d, 8, 2, 0)
        m = tm.Model(d, 2)
        xs = signals.batch_16k(n, hops, seed0=300)          # [hops][n][160]
        conds = []
        for s in range(n):
            x = torch.from_numpy(np.ascontiguousarray(xs[:, s, :].reshape(-1), np.float32))
            phone = m.phone(x)
            head = m.pitch(x)
            q = head[:, 1:m.bins].argmax(dim=1) + 1
            feat = head[:, m.bins:]
            spk = m.setter.additive(m.speakers.additive[s % 8])
            fm = m.setter.formant(m.speakers.formant[4])
            film = m.setter.film(m.speakers.kv[s % 8])
            conds.append((phone, q, feat, spk, fm, film))

        def run(schemes):
            return [wavegen_q(m.wavegen, *c, schemes) for c in conds]

        ref = run(["exact"] * 4)
        sig = float(torch.sqrt(torch.mean(torch.cat(ref) ** 2)))
        print(f"signal RMS {sig:.3f}, {hops} hops x {n} streams")

        def rms(schemes):
            got = run(schemes)
            return float(torch.sqrt(torch.mean((torch.cat(got) - torch.cat(ref)) ** 2)))

        print(f"all x3                         {rms(['x3'] * 4):.2e}")
        for alt in ("f16a", "f16", "bf16"):
            print(f"all {alt:<5s}                      {rms([alt] * 4):.2e}")
            for s in range(4):
                sch = ["x3"] * 4
                sch[s] = alt
                print(f"  stage {s} {alt:<5s}, others x3       {rms(sch):.2e}")
        for combo in (["x3", "x3", "f16a", "f16a"], ["x3", "f16a", "f16a", "f16a"]):
            print(f"{combo}   {rms(combo):.2e}")


if __name__ == "__main__":
    main()
