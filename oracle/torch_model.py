"""ORACLE CROSS-CHECK -- test infrastructure only.

An independent, NON-streaming PyTorch (CPU, fp32) statement of network spec
"M0": whole-utterance causal convolutions with explicit left zero padding, which
must agree with the frame-by-frame streaming C++ oracle
(``oracle/beatrice_oracle.cc``) to fp32 rounding.  Because the reference's real
inference library is closed source and absent ("parity
unpinned"), this second implementation written against the spec document --
not against the C++ code -- is what pins the oracle.

It parses the ``*.bin`` files itself (format: ``beatrice_vst_b200/model_spec.py``
module docstring) rather than sharing a loader with either library.
"""
from __future__ import annotations

import math
import os
import struct

import numpy as np
import torch
import torch.nn.functional as F

HIDDEN = 256
RATES = [5, 4, 4, 3]
STAGE_CH = [256, 128, 64, 32, 16]
MRF_K = [3, 7, 11]
MRF_D = [1, 3, 5]
FAMILY = {0: (256, 384, False), 1: (256, 384, False), 2: (128, 448, True)}
PHONE_FRONT = [(10, 1, 32, 5), (3, 32, 64, 2), (3, 64, 128, 2), (3, 128, 256, 2),
               (3, 256, 256, 2), (2, 256, 256, 2)]
PITCH_FRONT = [(10, 1, 16, 5), (3, 16, 32, 2), (3, 32, 64, 2), (3, 64, 128, 2),
               (3, 128, 128, 2), (2, 128, 128, 2)]


class _Reader:
    def __init__(self, path):
        raw = open(path, "rb").read()
        self.magic, self.family, self.kind, self.count = struct.unpack("<4I", raw[:16])
        assert self.magic == 0x42323042
        self.data = np.frombuffer(raw, dtype="<f4", offset=16)
        self.pos = 0

    def take(self, *shape):
        n = int(np.prod(shape))
        out = torch.from_numpy(self.data[self.pos:self.pos + n].copy()).reshape(*shape)
        self.pos += n
        return out

    def conv(self, k, cin, cout):
        w = self.take(k, cin, cout)           # [k][cin][cout] on disk
        b = self.take(cout)
        return w.permute(2, 1, 0).contiguous(), b   # torch: [cout][cin][k]

    def done(self):
        assert self.pos == self.data.size, (self.pos, self.data.size)


def _causal_conv(x, w, b, stride=1, dil=1):
    """x [1,C,T]; output step t ends at input step t*stride + stride - 1."""
    k = w.shape[2]
    pad = (k - 1) * dil - (stride - 1)
    return F.conv1d(F.pad(x, (pad, 0)), w, b, stride=stride, dilation=dil)


def _chan_norm(x, gamma, beta):
    mean = x.mean(dim=1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=1, keepdim=True)
    return (x - mean) / torch.sqrt(var + 1e-5) * gamma[None, :, None] + beta[None, :, None]


class Encoder:
    def __init__(self, path, front, dils, width, head_out):
        r = _Reader(path)
        self.front = [(r.conv(k, ci, co), s) for (k, ci, co, s) in front]
        self.res = []
        for d in dils:
            g, bt = r.take(width), r.take(width)
            self.res.append((g, bt, r.conv(3, width, width), d))
        self.head = r.conv(1, width, head_out)
        r.done()

    def __call__(self, x16k: torch.Tensor) -> torch.Tensor:
        x = x16k.reshape(1, 1, -1)
        for (w, b), s in self.front:
            x = F.gelu(_causal_conv(x, w, b, stride=s))
        for g, bt, (w, b), d in self.res:
            x = x + _causal_conv(F.gelu(_chan_norm(x, g, bt)), w, b, dil=d)
        w, b = self.head
        return F.conv1d(x, w, b)[0].T   # [frames][head_out]


class WaveGen:
    def __init__(self, path, family):
        pc, bins, self.has_setter = FAMILY[family]
        r = _Reader(path)
        self.embed = r.conv(1, pc, HIDDEN)
        self.pitch_emb = r.take(bins, HIDDEN)
        self.feat_proj = r.take(4, HIDDEN)
        self.pre = r.conv(7, HIDDEN, HIDDEN)
        self.ups, self.mrf = [], []
        for s, rate in enumerate(RATES):
            cin, cout = STAGE_CH[s], STAGE_CH[s + 1]
            w2 = r.take(2, cin, rate, cout)      # [tap][ci][phase][co]
            b = r.take(cout)
            # ConvTranspose1d weight [cin][cout][2r]: kernel index p <- tap 1, p+r <- tap 0
            wt = torch.cat([w2[1], w2[0]], dim=1).permute(0, 2, 1).contiguous()
            self.ups.append((wt, b, rate))
            branches = []
            for k in MRF_K:
                layers = []
                for d in MRF_D:
                    layers.append((r.conv(k, cout, cout), r.conv(k, cout, cout), d))
                branches.append(layers)
            self.mrf.append(branches)
        self.post = r.conv(7, 16, 1)
        r.done()

    def __call__(self, phone, q, feat, spk_add, formant_add, film):
        """phone [n,P]; q [n] long; feat [n,4]; spk_add/formant_add [256] (or speaker
        vector for a2/b1 in spk_add with formant_add None); film: list of 4 [2C] or None."""
        h = phone @ self.embed[0][:, :, 0].T + self.embed[1]
        h = h + self.pitch_emb[q]
        h = h + feat @ self.feat_proj
        if spk_add is not None:
            h = h + spk_add
        if formant_add is not None:
            h = h + formant_add
        x = h.T[None]                                         # [1,256,n]
        hidden = x
        x = _causal_conv(x, *self.pre)
        taps = {"hidden": hidden[0].T, "pre": x[0].T, "stage": []}
        for s in range(4):
            wt, b, rate = self.ups[s]
            n_in = x.shape[2]
            x = F.conv_transpose1d(F.leaky_relu(x, 0.1), wt, b, stride=rate)[:, :, :n_in * rate]
            if film is not None:
                c = STAGE_CH[s + 1]
                x = x * (1.0 + film[s][:c])[None, :, None] + film[s][c:][None, :, None]
            total = None
            for layers in self.mrf[s]:
                y = x
                for (w1, b1), (w2, b2), d in layers:
                    a = _causal_conv(F.leaky_relu(y, 0.1), w1, b1, dil=d)
                    y = y + _causal_conv(F.leaky_relu(a, 0.1), w2, b2)
                total = y if total is None else total + y
            x = total * np.float32(1.0 / 3.0)
            taps["stage"].append(x[0].T)
        out = torch.tanh(_causal_conv(F.leaky_relu(x, 0.1), *self.post))
        return out.reshape(-1), taps


class EmbSetter:
    def __init__(self, path):
        r = _Reader(path)
        self.add_w, self.add_b = r.take(HIDDEN, HIDDEN), r.take(HIDDEN)
        self.for_w, self.for_b = r.take(HIDDEN, HIDDEN), r.take(HIDDEN)
        self.blocks = []
        for blk in range(4):
            c = STAGE_CH[blk + 1]
            self.blocks.append((r.take(128), r.take(128, 2 * c), r.take(2 * c)))
        r.done()

    def additive(self, e):
        return e @ self.add_w + self.add_b

    def formant(self, e):
        return e @ self.for_w + self.for_b

    def film(self, kv):
        out = []
        for q, w, b in self.blocks:
            p = torch.softmax((kv @ q) / math.sqrt(128.0), dim=0)
            out.append((p @ kv) @ w + b)
        return out


class Speakers:
    def __init__(self, path, family):
        r = _Reader(path)
        n = r.count
        pc = FAMILY[family][0]
        if family == 2:
            self.formant = r.take(9, HIDDEN)
            self.codebook, self.additive, self.kv = [], [], []
            for _ in range(n):
                self.codebook.append(r.take(512, pc))
                self.additive.append(r.take(HIDDEN))
                self.kv.append(r.take(384, 128))
        else:
            self.additive = list(r.take(n, HIDDEN))
        r.done()


def vq_knn(phone, codebook, n):
    """Mean of the n nearest (squared L2) codebook rows, per frame."""
    d = (codebook ** 2).sum(1)[None, :] - 2.0 * phone @ codebook.T
    idx = torch.topk(-d, n, dim=1).indices
    return codebook[idx].mean(dim=1)


class Model:
    """Whole-utterance forward of one stream (default parameters of the call site)."""

    def __init__(self, model_dir, family=2):
        self.family = family
        pc, bins, has_setter = FAMILY[family]
        j = lambda s: os.path.join(model_dir, s)  # noqa: E731
        self.phone = Encoder(j("phone_extractor.bin"), PHONE_FRONT, [1, 2, 4, 1, 2, 4], 256, pc)
        self.pitch = Encoder(j("pitch_estimator.bin"), PITCH_FRONT, [1, 2, 4], 128, bins + 4)
        self.wavegen = WaveGen(j("waveform_generator.bin"), family)
        self.speakers = Speakers(j("speaker_embeddings.bin"), family)
        self.bins = bins
        if has_setter:
            self.setter = EmbSetter(j("embedding_setter.bin"))
        else:
            self.formant = _Reader(j("formant_shift_embeddings.bin")).take(9, HIDDEN)

    @torch.no_grad()
    def forward(self, x16k: np.ndarray, speaker=0, formant_index=4, min_q=1, max_q=None, vq=0,
                q_override=None):
        x = torch.from_numpy(np.ascontiguousarray(x16k, np.float32))
        phone = self.phone(x)
        if vq > 0 and self.family == 2:
            phone = vq_knn(phone, self.speakers.codebook[speaker], vq)
        head = self.pitch(x)
        max_q = self.bins - 1 if max_q is None else max_q
        q = head[:, min_q:max_q + 1].argmax(dim=1) + min_q
        feat = head[:, self.bins:]
        qq = q if q_override is None else torch.as_tensor(q_override, dtype=torch.long)
        if self.family == 2:
            spk = self.setter.additive(self.speakers.additive[speaker])
            fm = self.setter.formant(self.speakers.formant[formant_index])
            film = self.setter.film(self.speakers.kv[speaker])
            wave, taps = self.wavegen(phone, qq, feat, spk, fm, film)
        else:
            spk = self.speakers.additive[speaker] + self.formant[formant_index]
            wave, taps = self.wavegen(phone, qq, feat, spk, None, None)
        return phone.numpy(), q.numpy().astype(np.int32), feat.numpy(), wave.numpy(), taps
