"""Network spec "M0" and the seeded synthetic-weight generator.

The reference ships only the C header of its inference library
(``lib/beatricelib/beatrice.h``); the network behind it is closed source
(SURVEY.md section 0, fact 1).  This module therefore *defines* the network the
B200 engine and the CPU oracle both implement, inside the shapes the header
fixes (``beatrice.h:10-28``: 160 samples in, 240 out, hidden 256, phone
128/256, pitch bins 448/384, codebook 512x128, KV 384x128, 4 blocks), and it
writes seeded random weights in the on-disk layout described below.

On-disk format (all little endian)
----------------------------------
Every ``*.bin`` starts with a 16-byte header ``<4I``:
``magic, family, kind, count`` followed by ``float32`` payload.

* ``magic``  = 0x42323042  ("B02B")
* ``family`` = 0 (2.0.0-alpha.2), 1 (2.0.0-beta.1), 2 (2.0.0-rc.0)
* ``kind``   = 1 phone_extractor, 2 pitch_estimator, 3 waveform_generator,
  4 embedding_setter, 5 speaker_embeddings, 6 formant_shift_embeddings
* ``count``  = number of payload floats (kinds 1-4), number of speakers
  (kind 5), number of embeddings (kind 6)

The expected payload size of kinds 1-4 is fixed by the spec, which is what
gives ``Beatrice_kFileTooSmall / kFileTooLarge / kInvalidFileSize``
(``beatrice.h:30-36``) a meaning.

Conv weights are stored ``[k][C_in][C_out]`` (tap-major, output channel
fastest) followed by the bias ``[C_out]``; tap ``k-1`` multiplies the newest
sample (all convolutions are causal).  The ConvTranspose1d upsamplers
(kernel ``2r``, stride ``r``) are stored as a 2-tap conv with ``r*C_out``
output columns: ``w[tap][ci][p*C_out+co]`` with tap 0 = previous input row
(ConvTranspose kernel index ``p+r``), tap 1 = current row (kernel index
``p``), then bias ``[C_out]``.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

MAGIC = 0x42323042
HEADER_BYTES = 16

KIND_PHONE = 1
KIND_PITCH = 2
KIND_WAVEGEN = 3
KIND_EMBSETTER = 4
KIND_SPEAKERS = 5
KIND_FORMANT = 6

IN_HOP = 160
OUT_HOP = 240
HIDDEN = 256
CODEBOOK_SIZE = 512
KV_LENGTH = 384
KV_CHANNELS = 128
N_BLOCKS = 4
N_FORMANT = 9

VERSION_STRINGS = {0: "2.0.0-alpha.2", 1: "2.0.0-beta.1", 2: "2.0.0-rc.0"}


@dataclass(frozen=True)
class Family:
    """Per-API-family widths (``beatrice.h:17-28``)."""

    family: int
    phone_channels: int
    pitch_bins: int
    has_setter: bool  # rc0: EmbeddingSetter / FiLM / codebook VQ


FAMILIES = {
    0: Family(0, 256, 384, False),
    1: Family(1, 256, 384, False),
    2: Family(2, 128, 448, True),
}

# (k, C_in, C_out, stride) of the strided causal front end; product of strides = 160
PHONE_FRONT = [(10, 1, 32, 5), (3, 32, 64, 2), (3, 64, 128, 2), (3, 128, 256, 2),
               (3, 256, 256, 2), (2, 256, 256, 2)]
PHONE_RES_DIL = [1, 2, 4, 1, 2, 4]
PITCH_FRONT = [(10, 1, 16, 5), (3, 16, 32, 2), (3, 32, 64, 2), (3, 64, 128, 2),
               (3, 128, 128, 2), (2, 128, 128, 2)]
PITCH_RES_DIL = [1, 2, 4]
PITCH_FEATS = 4

UPS_RATES = [5, 4, 4, 3]          # product = 240
STAGE_CH = [256, 128, 64, 32, 16]  # C before stage s / after stage s
MRF_KERNELS = [3, 7, 11]
MRF_DILATIONS = [1, 3, 5]
PRE_K = 7
POST_K = 7


def _conv(name, k, cin, cout, std, bias_std=0.02):
    return [(name + ".w", (k, cin, cout), std), (name + ".b", (cout,), bias_std)]


def phone_layout(fam: Family):
    """[(name, shape, std)] in file order for phone_extractor.bin."""
    out = []
    for i, (k, cin, cout, _s) in enumerate(PHONE_FRONT):
        gain = 3.0 if i == 0 else 1.6
        out += _conv(f"fe{i}", k, cin, cout, gain / np.sqrt(k * cin))
    for i, _d in enumerate(PHONE_RES_DIL):
        out += [(f"res{i}.gamma", (256,), -0.1), (f"res{i}.beta", (256,), 0.1)]
        out += _conv(f"res{i}.conv", 3, 256, 256, 0.7 / np.sqrt(3 * 256))
    out += _conv("head", 1, 256, fam.phone_channels, 0.6 / np.sqrt(256))
    return out


def pitch_layout(fam: Family):
    out = []
    for i, (k, cin, cout, _s) in enumerate(PITCH_FRONT):
        gain = 3.0 if i == 0 else 1.6
        out += _conv(f"fe{i}", k, cin, cout, gain / np.sqrt(k * cin))
    for i, _d in enumerate(PITCH_RES_DIL):
        out += [(f"res{i}.gamma", (128,), -0.1), (f"res{i}.beta", (128,), 0.1)]
        out += _conv(f"res{i}.conv", 3, 128, 128, 0.7 / np.sqrt(3 * 128))
    out += _conv("head", 1, 128, fam.pitch_bins + PITCH_FEATS, 2.0 / np.sqrt(128))
    return out


def wavegen_layout(fam: Family):
    out = []
    out += _conv("embed_phone", 1, fam.phone_channels, HIDDEN,
                 1.0 / np.sqrt(fam.phone_channels))
    out += [("pitch_emb", (fam.pitch_bins, HIDDEN), 0.5)]
    out += [("feat_proj", (PITCH_FEATS, HIDDEN), 0.3)]
    out += _conv("pre", PRE_K, HIDDEN, HIDDEN, 1.0 / np.sqrt(PRE_K * HIDDEN))
    for s, r in enumerate(UPS_RATES):
        cin, cout = STAGE_CH[s], STAGE_CH[s + 1]
        out += [(f"ups{s}.w", (2, cin, r * cout), 1.0 / np.sqrt(cin)),
                (f"ups{s}.b", (cout,), 0.02)]
        for k in MRF_KERNELS:
            for d in MRF_DILATIONS:
                out += _conv(f"mrf{s}.k{k}.d{d}.c1", k, cout, cout,
                             1.4 / np.sqrt(k * cout))
                out += _conv(f"mrf{s}.k{k}.d{d}.c2", k, cout, cout,
                             0.45 / np.sqrt(k * cout))
    out += _conv("post", POST_K, STAGE_CH[-1], 1,
                 0.35 / np.sqrt(POST_K * STAGE_CH[-1]))
    return out


def embsetter_layout(fam: Family):
    assert fam.has_setter
    out = []
    out += [("add_proj.w", (HIDDEN, HIDDEN), 0.5 / np.sqrt(HIDDEN)),
            ("add_proj.b", (HIDDEN,), 0.02)]
    out += [("formant_proj.w", (HIDDEN, HIDDEN), 0.5 / np.sqrt(HIDDEN)),
            ("formant_proj.b", (HIDDEN,), 0.02)]
    for blk in range(N_BLOCKS):
        c = STAGE_CH[blk + 1]
        out += [(f"film{blk}.query", (KV_CHANNELS,), 1.0),
                (f"film{blk}.w", (KV_CHANNELS, 2 * c), 0.3 / np.sqrt(KV_CHANNELS)),
                (f"film{blk}.b", (2 * c,), 0.02)]
    return out


LAYOUTS = {KIND_PHONE: phone_layout, KIND_PITCH: pitch_layout,
           KIND_WAVEGEN: wavegen_layout, KIND_EMBSETTER: embsetter_layout}
FILE_NAMES = {KIND_PHONE: "phone_extractor.bin", KIND_PITCH: "pitch_estimator.bin",
              KIND_WAVEGEN: "waveform_generator.bin",
              KIND_EMBSETTER: "embedding_setter.bin",
              KIND_SPEAKERS: "speaker_embeddings.bin",
              KIND_FORMANT: "formant_shift_embeddings.bin"}


def n_params(kind: int, family: int = 2) -> int:
    return int(sum(int(np.prod(s)) for _n, s, _std in LAYOUTS[kind](FAMILIES[family])))


def speakers_payload_floats(n_speakers: int, family: int = 2) -> int:
    if family == 2:
        per = CODEBOOK_SIZE * FAMILIES[2].phone_channels + HIDDEN + KV_LENGTH * KV_CHANNELS
        return N_FORMANT * HIDDEN + n_speakers * per
    return n_speakers * HIDDEN


def _draw(rng, shape, std):
    """std < 0 means "1 + |std| * N(0,1)" (norm gains)."""
    n = int(np.prod(shape))
    z = rng.standard_normal(n, dtype=np.float32)
    if std < 0:
        return (1.0 + (-std) * z).astype(np.float32)
    return (np.float32(std) * z).astype(np.float32)


def _write_bin(path, family, kind, count, payload: np.ndarray):
    with open(path, "wb") as f:
        f.write(struct.pack("<4I", MAGIC, family, kind, count))
        f.write(np.ascontiguousarray(payload, dtype="<f4").tobytes())


def generate_tensors(kind: int, family: int, seed: int):
    """Deterministic {name: array} for one parameter file."""
    rng = np.random.default_rng([seed, family, kind])
    return {name: _draw(rng, shape, std).reshape(shape)
            for name, shape, std in LAYOUTS[kind](FAMILIES[family])}


def generate_speakers(n_speakers: int, family: int, seed: int):
    rng = np.random.default_rng([seed, family, KIND_SPEAKERS])
    fam = FAMILIES[family]
    out = {}
    if family == 2:
        out["formant"] = _draw(rng, (N_FORMANT, HIDDEN), 1.0).reshape(N_FORMANT, HIDDEN)
        out["codebook"] = np.empty((n_speakers, CODEBOOK_SIZE, fam.phone_channels), np.float32)
        out["additive"] = np.empty((n_speakers, HIDDEN), np.float32)
        out["kv"] = np.empty((n_speakers, KV_LENGTH, KV_CHANNELS), np.float32)
        for i in range(n_speakers):
            out["codebook"][i] = _draw(rng, (CODEBOOK_SIZE, fam.phone_channels), 1.0).reshape(
                CODEBOOK_SIZE, fam.phone_channels)
            out["additive"][i] = _draw(rng, (HIDDEN,), 1.0)
            out["kv"][i] = _draw(rng, (KV_LENGTH, KV_CHANNELS), 1.0).reshape(KV_LENGTH, KV_CHANNELS)
    else:
        out["additive"] = _draw(rng, (n_speakers, HIDDEN), 1.0).reshape(n_speakers, HIDDEN)
        rng2 = np.random.default_rng([seed, family, KIND_FORMANT])
        out["formant"] = _draw(rng2, (N_FORMANT, HIDDEN), 1.0).reshape(N_FORMANT, HIDDEN)
    return out


def write_model_dir(path: str, n_speakers: int = 8, family: int = 2, seed: int = 0) -> str:
    """Write a model directory; returns the path of the ``.toml`` to load.

    Layout follows ``ProcessorCore2::LoadModel`` (reference
    ``src/common/processor_core_2.cc:301-351``) and the TOML schema of
    ``src/common/model_config.h:77-138``.
    """
    os.makedirs(path, exist_ok=True)
    fam = FAMILIES[family]
    kinds = [KIND_PHONE, KIND_PITCH, KIND_WAVEGEN] + ([KIND_EMBSETTER] if fam.has_setter else [])
    for kind in kinds:
        tensors = generate_tensors(kind, family, seed)
        payload = np.concatenate([t.ravel() for t in tensors.values()])
        _write_bin(os.path.join(path, FILE_NAMES[kind]), family, kind, payload.size, payload)
    spk = generate_speakers(n_speakers, family, seed)
    if family == 2:
        parts = [spk["formant"].ravel()]
        for i in range(n_speakers):
            parts += [spk["codebook"][i].ravel(), spk["additive"][i].ravel(), spk["kv"][i].ravel()]
        _write_bin(os.path.join(path, FILE_NAMES[KIND_SPEAKERS]), family, KIND_SPEAKERS,
                   n_speakers, np.concatenate(parts))
    else:
        _write_bin(os.path.join(path, FILE_NAMES[KIND_SPEAKERS]), family, KIND_SPEAKERS,
                   n_speakers, spk["additive"].ravel())
        _write_bin(os.path.join(path, FILE_NAMES[KIND_FORMANT]), family, KIND_FORMANT,
                   N_FORMANT, spk["formant"].ravel())
    toml_path = os.path.join(path, "model.toml")
    with open(toml_path, "w", encoding="utf-8") as f:
        f.write("[model]\n")
        f.write(f'version = "{VERSION_STRINGS[family]}"\n')
        f.write(f'name = "M0 synthetic (seed {seed})"\n')
        f.write('description = "builder-defined spec M0, seeded random weights"\n\n')
        for i in range(n_speakers):
            f.write(f"[voice.{i}]\n")
            f.write(f'name = "voice {i}"\n')
            f.write(f'description = "synthetic speaker {i}"\n')
            f.write(f"average_pitch = {52.0 + i:.1f}\n")
            f.write(f"[voice.{i}.portrait]\n")
            f.write(f'path = "portrait_{i}.png"\n')
            f.write('description = "none"\n\n')
    return toml_path


def flops_per_frame(family: int = 2):
    """Algorithmic FLOPs per 10 ms frame per stream (2*C_in*k*C_out*T_out)."""
    fam = FAMILIES[family]

    def front(layers, res_c, res_n, head_out):
        t, fl = IN_HOP, 0
        for k, cin, cout, s in layers:
            t //= s
            fl += 2 * cin * k * cout * t
        fl += res_n * 2 * res_c * 3 * res_c
        fl += 2 * res_c * head_out
        return fl

    phone = front(PHONE_FRONT, 256, len(PHONE_RES_DIL), fam.phone_channels)
    pitch = front(PITCH_FRONT, 128, len(PITCH_RES_DIL), fam.pitch_bins + PITCH_FEATS)
    wg = 2 * fam.phone_channels * HIDDEN + 2 * PITCH_FEATS * HIDDEN
    wg += 2 * HIDDEN * PRE_K * HIDDEN
    t = 1
    mrf_total = 0
    for s, r in enumerate(UPS_RATES):
        cin, cout = STAGE_CH[s], STAGE_CH[s + 1]
        wg += 2 * cin * 2 * r * cout * t
        t *= r
        mrf = sum(2 * 2 * k * cout * cout * t * len(MRF_DILATIONS) for k in MRF_KERNELS)
        mrf_total += mrf
        wg += mrf
    wg += 2 * STAGE_CH[-1] * POST_K * t
    return {"phone": phone, "pitch": pitch, "wavegen": wg, "mrf": mrf_total,
            "total": phone + pitch + wg}


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description="write a synthetic M0 model directory")
    ap.add_argument("out")
    ap.add_argument("--speakers", type=int, default=8)
    ap.add_argument("--family", type=int, default=2)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    print(write_model_dir(a.out, a.speakers, a.family, a.seed))
    print({k: n_params(k, a.family) for k in LAYOUTS if k != KIND_EMBSETTER or a.family == 2})
    print(flops_per_frame(a.family))
