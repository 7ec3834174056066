"""ctypes bindings for the beatricelib C ABI (reference ``lib/beatricelib/beatrice.h``).

Binds the CUDA product library (``csrc/libbeatrice_b200.so``).  The class is ABI-generic, which is why the
test infrastructure (``oracle/loader.py``) can reuse it for the CPU oracle -- nothing in this package loads
or knows about the oracle.

There is no CPU fallback: :func:`load_product` raises if the CUDA library is missing; with no usable GPU the
library latches an error (``BeatriceB200_LastError``), reports it on stderr and writes silence.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
PRODUCT_SO = os.path.join(_HERE, "csrc", "libbeatrice_b200.so")

IN_HOP = 160
OUT_HOP = 240
HIDDEN = 256
FAMILY_PREFIX = {0: "Beatrice20a2", 1: "Beatrice20b1", 2: "Beatrice20rc0"}
PHONE_CHANNELS = {0: 256, 1: 256, 2: 128}
PITCH_BINS = {0: 384, 1: 384, 2: 448}

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)
_vp = C.c_void_p


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


# name suffix -> (restype, argtypes); every symbol of beatrice.h, per family
_COMMON = {
    "CreatePhoneExtractor": (_vp, []),
    "DestroyPhoneExtractor": (None, [_vp]),
    "CreatePhoneContext1": (_vp, []),
    "DestroyPhoneContext1": (None, [_vp]),
    "ReadPhoneExtractorParameters": (C.c_int, [_vp, C.c_char_p]),
    "ExtractPhone1": (None, [_vp, _f32p, _f32p, _vp]),
    "CreatePitchEstimator": (_vp, []),
    "DestroyPitchEstimator": (None, [_vp]),
    "CreatePitchContext1": (_vp, []),
    "DestroyPitchContext1": (None, [_vp]),
    "ReadPitchEstimatorParameters": (C.c_int, [_vp, C.c_char_p]),
    "SetMinQuantizedPitch": (None, [_vp, C.c_int]),
    "SetMaxQuantizedPitch": (None, [_vp, C.c_int]),
    "EstimatePitch1": (None, [_vp, _f32p, _i32p, _f32p, _vp]),
    "ReadNSpeakers": (C.c_int, [C.c_char_p, _i32p]),
    "CreateWaveformGenerator": (_vp, []),
    "DestroyWaveformGenerator": (None, [_vp]),
    "CreateWaveformContext1": (_vp, []),
    "DestroyWaveformContext1": (None, [_vp]),
    "ReadWaveformGeneratorParameters": (C.c_int, [_vp, C.c_char_p]),
}
_LEGACY = {
    "ReadSpeakerEmbeddings": (C.c_int, [C.c_char_p, _f32p]),
    "GenerateWaveform1": (None, [_vp, _f32p, _i32p, _f32p, _f32p, _f32p, _vp]),
}
_RC0 = {
    "SetVQNumNeighbors": (None, [_vp, C.c_int]),
    "ReadSpeakerEmbeddings": (C.c_int, [C.c_char_p, _f32p, _f32p, _f32p, _f32p]),
    "GenerateWaveform1": (None, [_vp, _f32p, _i32p, _f32p, _f32p, _vp]),
    "CreateEmbeddingSetter": (_vp, []),
    "DestroyEmbeddingSetter": (None, [_vp]),
    "CreateEmbeddingContext": (_vp, []),
    "DestroyEmbeddingContext": (None, [_vp]),
    "ReadEmbeddingSetterParameters": (C.c_int, [_vp, C.c_char_p]),
    "SetCodebook": (None, [_vp, _f32p]),
    "SetAdditiveSpeakerEmbedding": (None, [_vp, _f32p, _vp, _vp]),
    "SetFormantShiftEmbedding": (None, [_vp, _f32p, _vp, _vp]),
    "RegisterKeyValueSpeakerEmbedding": (None, [_vp, _f32p, _vp]),
    "SetKeyValueSpeakerEmbedding": (None, [_vp, C.c_int, _vp, _vp]),
}


def abi_symbols(family: int):
    """All C symbols ``beatrice.h`` declares for one family."""
    table = dict(_COMMON)
    table.update(_RC0 if family == 2 else _LEGACY)
    return {f"{FAMILY_PREFIX[family]}_{k}": v for k, v in table.items()}


def all_abi_symbols():
    out = {}
    for fam in (0, 1, 2):
        out.update(abi_symbols(fam))
    return out


class BeatriceLib:
    """A loaded library exporting the beatrice.h ABI."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found -- run `python -c 'import __graft_entry__ as g; g.build()'`")
        self.path = path
        self.dll = C.CDLL(path, mode=C.RTLD_LOCAL)
        for name, (res, args) in all_abi_symbols().items():
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args

    def fn(self, family: int, suffix: str):
        return getattr(self.dll, f"{FAMILY_PREFIX[family]}_{suffix}")


def load_product() -> BeatriceLib:
    return BeatriceLib(PRODUCT_SO)


class SingleStream:
    """One voice stream through the per-frame ABI, the way
    ``ProcessorCore2::Process1`` drives it (reference
    ``src/common/processor_core_2.cc:181-255``) -- minus the call-site pitch
    transform, which :mod:`tests` cover through the real call site.

    Used by tests to feed identical frames to the oracle and to the CUDA
    library.  ``family`` 0/1 use the a2/b1 entry points
    (``processor_core_0.cc:50-142``).
    """

    def __init__(self, lib: BeatriceLib, model_dir: str, family: int = 2, speaker: int = 0,
                 formant_index: int = 4):
        self.lib, self.family = lib, family
        f = lambda s: lib.fn(family, s)  # noqa: E731
        self.f = f
        enc = lambda s: os.path.join(model_dir, s).encode("utf-8")  # noqa: E731
        self.pe, self.pi, self.wg = f("CreatePhoneExtractor")(), f("CreatePitchEstimator")(), \
            f("CreateWaveformGenerator")()
        self.pc, self.pic, self.wc = f("CreatePhoneContext1")(), f("CreatePitchContext1")(), \
            f("CreateWaveformContext1")()
        self.errors = [
            f("ReadPhoneExtractorParameters")(self.pe, enc("phone_extractor.bin")),
            f("ReadPitchEstimatorParameters")(self.pi, enc("pitch_estimator.bin")),
            f("ReadWaveformGeneratorParameters")(self.wg, enc("waveform_generator.bin")),
        ]
        n = C.c_int(0)
        self.errors.append(f("ReadNSpeakers")(enc("speaker_embeddings.bin"), C.byref(n)))
        self.n_speakers = n.value
        pc = PHONE_CHANNELS[family]
        if family == 2:
            self.es, self.ec = f("CreateEmbeddingSetter")(), f("CreateEmbeddingContext")()
            self.errors.append(f("ReadEmbeddingSetterParameters")(self.es, enc("embedding_setter.bin")))
            ns = max(self.n_speakers, 1)
            self.codebooks = np.zeros((ns, 512, pc), np.float32)
            self.additive = np.zeros((ns, HIDDEN), np.float32)
            self.formant = np.zeros((9, HIDDEN), np.float32)
            self.kv = np.zeros((ns, 384, 128), np.float32)
            self.errors.append(f("ReadSpeakerEmbeddings")(
                enc("speaker_embeddings.bin"), _fp(self.codebooks), _fp(self.additive),
                _fp(self.formant), _fp(self.kv)))
            if all(e == 0 for e in self.errors):
                self.set_speaker(speaker)
                self.set_formant_index(formant_index)
        else:
            ns = max(self.n_speakers, 1)
            self.additive = np.zeros((ns, HIDDEN), np.float32)
            self.formant = np.zeros((9, HIDDEN), np.float32)
            self.errors.append(f("ReadSpeakerEmbeddings")(enc("speaker_embeddings.bin"), _fp(self.additive)))
            self.errors.append(f("ReadSpeakerEmbeddings")(enc("formant_shift_embeddings.bin"), _fp(self.formant)))
            self.speaker_vec = np.ascontiguousarray(self.additive[speaker] + self.formant[formant_index])
        self.ok = all(e == 0 for e in self.errors)

    # -- rc0 set-speaker path: processor_core_2.cc:431-466 + :161-169 --
    def set_speaker(self, speaker: int, kv_blocks_now: bool = True):
        f = self.f
        self._cb = np.ascontiguousarray(self.codebooks[speaker])
        self._add = np.ascontiguousarray(self.additive[speaker])
        self._kv = np.ascontiguousarray(self.kv[speaker])
        f("SetCodebook")(self.pc, _fp(self._cb))
        f("SetAdditiveSpeakerEmbedding")(self.es, _fp(self._add), self.ec, self.wc)
        f("RegisterKeyValueSpeakerEmbedding")(self.es, _fp(self._kv), self.ec)
        if kv_blocks_now:
            for b in range(4):
                f("SetKeyValueSpeakerEmbedding")(self.es, b, self.ec, self.wc)

    def set_kv_block(self, block: int):
        self.f("SetKeyValueSpeakerEmbedding")(self.es, block, self.ec, self.wc)

    def set_formant_index(self, index: int):
        self._fm = np.ascontiguousarray(self.formant[index])
        self.f("SetFormantShiftEmbedding")(self.es, _fp(self._fm), self.ec, self.wc)

    def set_vq(self, n: int):
        self.f("SetVQNumNeighbors")(self.pc, n)

    def set_pitch_range(self, lo: int, hi: int):
        self.f("SetMinQuantizedPitch")(self.pic, lo)
        self.f("SetMaxQuantizedPitch")(self.pic, hi)

    def frame(self, x160: np.ndarray, q_override=None):
        """Returns (phone, q, feat, wave240) for one 10 ms frame."""
        f, fam = self.f, self.family
        x = np.ascontiguousarray(x160, np.float32)
        phone = np.empty(PHONE_CHANNELS[fam], np.float32)
        feat = np.empty(4, np.float32)
        wave = np.empty(OUT_HOP, np.float32)
        q = C.c_int(0)
        f("ExtractPhone1")(self.pe, _fp(x), _fp(phone), self.pc)
        f("EstimatePitch1")(self.pi, _fp(x), C.byref(q), _fp(feat), self.pic)
        qq = C.c_int(q.value if q_override is None else int(q_override))
        if fam == 2:
            f("GenerateWaveform1")(self.wg, _fp(phone), C.byref(qq), _fp(feat), _fp(wave), self.wc)
        else:
            f("GenerateWaveform1")(self.wg, _fp(phone), C.byref(qq), _fp(feat), _fp(self.speaker_vec),
                                   _fp(wave), self.wc)
        return phone, q.value, feat, wave

    def run(self, x16k: np.ndarray):
        n = len(x16k) // IN_HOP
        fam = self.family
        phones = np.empty((n, PHONE_CHANNELS[fam]), np.float32)
        qs = np.empty(n, np.int32)
        feats = np.empty((n, 4), np.float32)
        waves = np.empty((n, OUT_HOP), np.float32)
        for i in range(n):
            phones[i], qs[i], feats[i], waves[i] = self.frame(x16k[i * IN_HOP:(i + 1) * IN_HOP])
        return phones, qs, feats, waves

    def close(self):
        f = self.f
        f("DestroyPhoneContext1")(self.pc)
        f("DestroyPitchContext1")(self.pic)
        f("DestroyWaveformContext1")(self.wc)
        f("DestroyPhoneExtractor")(self.pe)
        f("DestroyPitchEstimator")(self.pi)
        f("DestroyWaveformGenerator")(self.wg)
        if self.family == 2:
            f("DestroyEmbeddingContext")(self.ec)
            f("DestroyEmbeddingSetter")(self.es)
