"""CPU: the oracle against (1) the independent PyTorch model, (2) the committed golden
vectors, plus file-format error codes and ABI export checks for both libraries."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from beatrice_vst_b200 import batch as bbatch
from beatrice_vst_b200 import lib as blib
from beatrice_vst_b200 import model_spec, signals
from conftest import ROOT, oracle_loader, rms

GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("family,vq", [(2, 0), (2, 3), (0, 0), (1, 0)])
def test_oracle_matches_independent_torch_model(oracle, model_dirs, family, vq):
    """Streaming C++ oracle == whole-utterance PyTorch model of spec M0 (pins the oracle)."""
    import torch_model
    x = signals.voice_like(160 * 25, 16000.0, seed=11)
    s = blib.SingleStream(oracle, model_dirs[family], family=family, speaker=2, formant_index=3)
    assert s.ok, s.errors
    s.set_pitch_range(1, 383)
    if family == 2:
        s.set_vq(vq)
    phone, q, feat, wave = s.run(x)
    s.close()
    m = torch_model.Model(model_dirs[family], family)
    tp, tq, tf, tw, _ = m.forward(x, speaker=2, formant_index=3, min_q=1, max_q=383, vq=vq)
    assert np.array_equal(q, tq)
    assert rms(phone, tp) <= 1e-5
    assert rms(feat, tf) <= 1e-5
    assert rms(wave.ravel(), tw) <= 1e-5          # fp32 tolerance; both are fp32 CPU
    assert wave.std() > 0.05                      # not degenerate


@pytest.mark.parametrize("family", [0, 2])
def test_oracle_matches_golden(oracle, model_dirs, family):
    g = np.load(os.path.join(GOLDEN, f"m0_family{family}.npz"))
    s = blib.SingleStream(oracle, model_dirs[family], family=family, speaker=1, formant_index=5)
    s.set_pitch_range(1, 383)
    phone, q, feat, wave = s.run(g["x"])
    s.close()
    assert np.array_equal(q, g["q"])
    assert rms(phone, g["phone"]) <= 1e-6 and rms(wave, g["wave"]) <= 1e-6
    if family == 2:
        s = blib.SingleStream(oracle, model_dirs[2], family=2, speaker=1, formant_index=5)
        s.set_pitch_range(1, 383)
        s.set_vq(4)
        p2, _, _, w2 = s.run(g["x"])
        s.close()
        assert rms(p2, g["phone_vq4"]) <= 1e-6 and rms(w2, g["wave_vq4"]) <= 1e-6


def test_pitch_range_is_respected(oracle, model_dir):
    x = signals.voice_like(160 * 10, 16000.0, seed=2)
    s = blib.SingleStream(oracle, model_dir)
    s.set_pitch_range(100, 120)
    _, q, _, _ = s.run(x)
    s.close()
    assert q.min() >= 100 and q.max() <= 120


def test_spec_flops_match_survey():
    f = model_spec.flops_per_frame(2)
    assert f["total"] == 86_808_064 and f["wavegen"] == 80_910_848   # SURVEY.md App. B totals


# ---------------------------------------------------------------------------------------
# both libraries: ABI export + reader error codes (host logic only, no GPU needed)
# ---------------------------------------------------------------------------------------
def _libs():
    return [("oracle", oracle_loader.ORACLE_SO), ("product", blib.PRODUCT_SO)]


@pytest.mark.parametrize("name,path", _libs())
def test_exports_every_symbol_of_reference_header(name, path, oracle, product):
    dll = C.CDLL(path)
    names = open(os.path.join(GOLDEN, "beatrice_h_symbols.txt")).read().split()
    assert len(names) == 77
    assert sorted(names) == sorted(blib.all_abi_symbols())
    for n in names:
        assert hasattr(dll, n), f"{name} library lacks {n}"


def test_product_exports_batched_api_declared_in_header(product):
    header = open(os.path.join(ROOT, "include", "beatrice_b200.h")).read()
    for n in bbatch.BATCH_SYMBOLS:
        assert n + "(" in header, f"{n} not declared in include/beatrice_b200.h"
        assert hasattr(product.dll, n), f"product library lacks {n}"
    import re
    declared = set(re.findall(r"\b(BeatriceB200_[A-Za-z0-9]+)\(", header))
    assert declared == set(bbatch.BATCH_SYMBOLS)


@pytest.mark.parametrize("which", ["oracle", "product"])
def test_reader_error_codes(which, oracle, product, model_dir, tmp_path):
    """Beatrice_ErrorCode values 0..4 (reference beatrice.h:30-36) on malformed files; the
    product validates on the host before touching CUDA, so this runs without a GPU."""
    L = oracle if which == "oracle" else product
    f = lambda s: L.fn(2, s)  # noqa: E731
    pe = f("CreatePhoneExtractor")()
    src = open(os.path.join(model_dir, "phone_extractor.bin"), "rb").read()
    cases = {
        "missing.bin": (None, 1),
        "short.bin": (src[:-400], 2),
        "long.bin": (src + b"\0" * 64, 3),
        "odd.bin": (src + b"\0", 4),
        "badmagic.bin": (struct.pack("<I", 0xdeadbeef) + src[4:], 4),
        "wrongkind.bin": (open(os.path.join(model_dir, "pitch_estimator.bin"), "rb").read(), 4),
        "tiny.bin": (b"\0" * 8, 2),
    }
    for name, (data, expect) in cases.items():
        p = tmp_path / name
        if data is not None:
            p.write_bytes(data)
        assert f("ReadPhoneExtractorParameters")(pe, str(p).encode()) == expect, name
    n = C.c_int(-1)
    spk = os.path.join(model_dir, "speaker_embeddings.bin").encode()
    assert f("ReadNSpeakers")(spk, C.byref(n)) == 0 and n.value == 8
    raw = open(spk, "rb").read()
    (tmp_path / "spk_short.bin").write_bytes(raw[:-4])
    assert f("ReadNSpeakers")(str(tmp_path / "spk_short.bin").encode(), C.byref(n)) == 2
    f("DestroyPhoneExtractor")(pe)


def test_product_reports_no_device_without_aborting(product):
    n = bbatch.device_count(product)
    assert n >= 0
    if n == 0:
        with pytest.raises(RuntimeError):
            bbatch.Engine(product, 4)
