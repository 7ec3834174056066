"""CPU: the REFERENCE's own call site (src/common compiled in place -> oracle/_ref/) driving
the oracle through the beatrice.h ABI.  These are the reference-derived known answers of
SURVEY.md section 8(c): unloaded behaviour, default pitch range, pitch-shift identities,
formant index, the 4-hop key-value schedule, FIFO block-size independence -- and they pin the
numpy restatement of gain/resample (oracle/hostrate_ref.py) bit-for-bit."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import callsite
import hostrate_ref
from beatrice_vst_b200 import lib as blib
from beatrice_vst_b200 import signals
from conftest import ROOT

pytestmark = pytest.mark.skipif(not callsite.available("oracle"),
                                reason="oracle/_ref not built (needs /root/reference at build time)")
_fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731


class Core2Emu:
    """Process1 of the call site (processor_core_2.cc:181-255) on top of SingleStream, with the
    fp64 pitch transform restated in Python; used as the expected value for the runner."""

    def __init__(self, lib, model_dir):
        self.s = blib.SingleStream(lib, model_dir)
        self.s.set_pitch_range(1, 383)       # defaults 33.125 / 80.875 -> bins 1 / 383
        self.avg, self.inton, self.shift, self.corr, self.ctype = 52.0, 1.0, 0.0, 0.0, 0
        self.kv_pending = 4
        self.q_log = []

    def set_speaker(self, spk):
        self.s.set_speaker(spk, kv_blocks_now=False)
        self.kv_pending = 0

    def transform(self, q):
        bps = 96.0 / 12.0
        t = self.avg + (float(q) - self.avg) * self.inton + bps * self.shift
        if self.corr != 0.0:
            if self.ctype == 0:
                nearest = (np.floor(t / bps) + 0.5) * bps
                nd = (t - nearest) * (2.0 / bps)
                t = nearest if abs(nd) < 1e-4 else nearest + nd * abs(nd) ** (-self.corr) * (bps / 2.0)
            else:
                nearest = np.round(t / bps) * bps   # never hit with ties in these tests
                nd = (t - nearest) * (2.0 / bps)
                if self.corr > 1 - 1e-4:
                    t = nearest
                elif nd >= 0:
                    t = nearest + nd ** (1.0 / (1.0 - self.corr)) * (bps / 2.0)
                else:
                    t = nearest - (-nd) ** (1.0 / (1.0 - self.corr)) * (bps / 2.0)
        r = np.floor(abs(t) + 0.5) * np.sign(t)      # std::round: half away from zero
        return int(min(max(int(r), 1), 447))

    def frame(self, x16):
        s = self.s
        if self.kv_pending < 4:
            s.set_kv_block(self.kv_pending)
            self.kv_pending += 1
        f = s.f
        x = np.ascontiguousarray(x16, np.float32)
        ph, ft, w = np.empty(128, np.float32), np.empty(4, np.float32), np.empty(240, np.float32)
        q = C.c_int(0)
        f("ExtractPhone1")(s.pe, _fp(x), _fp(ph), s.pc)
        f("EstimatePitch1")(s.pi, _fp(x), C.byref(q), _fp(ft), s.pic)
        qq = C.c_int(self.transform(q.value))
        self.q_log.append((q.value, qq.value))
        f("GenerateWaveform1")(s.wg, _fp(ph), C.byref(qq), _fp(ft), _fp(w), s.wc)
        return w


def _toml(model_dir):
    return os.path.join(model_dir, "model.toml")


def test_unloaded_core_outputs_zeros(model_dir):
    x = signals.voice_like(480 * 3, 48000.0, 0)
    y, info = callsite.run("oracle", None, x)
    assert info["version"] == -1 and info["last"] == 0 and not y.any()


def test_load_errors_propagate_as_error_codes(model_dir, tmp_path):
    import shutil
    bad = tmp_path / "bad"
    shutil.copytree(model_dir, bad)
    data = (bad / "waveform_generator.bin").read_bytes()
    (bad / "waveform_generator.bin").write_bytes(data[:-8])
    x = signals.voice_like(480, 48000.0, 0)
    y, info = callsite.run("oracle", str(bad / "model.toml"), x)
    assert info["load"] == 2 and info["version"] == -1 and not y.any()    # kFileTooSmall -> unloaded core
    os.remove(bad / "pitch_estimator.bin")
    _, info = callsite.run("oracle", str(bad / "model.toml"), x)
    assert info["load"] == 1                                               # kFileOpenError


def test_defaults_match_numpy_restatement_bit_exact(oracle, model_dir):
    x = signals.voice_like(480 * 25, 48000.0, 3)
    y, info = callsite.run("oracle", _toml(model_dir), x)
    assert info == {"load": 0, "last": 0, "version": 2}
    emu = Core2Emu(oracle, model_dir)
    hr = hostrate_ref.HostRateRef(emu.frame)
    y2 = np.concatenate([hr.process(x[i * 480:(i + 1) * 480]) for i in range(25)])
    assert np.array_equal(y, y2)
    assert not y[:480].any() and y[481:960].any()      # one-block FIFO delay (resample.h:343-363)
    assert all(a == b for a, b in emu.q_log)             # defaults leave the pitch bin unchanged


def test_parameters_events_match_emulation_bit_exact(oracle, model_dir):
    """gain slews, +12 st pitch shift (q+96, clamped), formant index round(shift*2+4),
    speaker change with the key-value blocks spread over four hops, pitch correction."""
    x = signals.voice_like(480 * 30, 48000.0, 5)
    events = [(-1, "input_gain", -6.0), (-1, "pitch_shift", 12.0), (4, "output_gain", 3.0),
              (8, "voice", 5), (15, "formant_shift", 1.0), (20, "pitch_correction", 0.4),
              (22, "intonation_intensity", 1.5), (24, "average_source_pitch", 60.0), (26, "vq_num_neighbors", 3)]
    y, info = callsite.run("oracle", _toml(model_dir), x, events=events)
    assert info["load"] == 0 and info["last"] == 0
    emu = Core2Emu(oracle, model_dir)
    emu.shift = 12.0
    hr = hostrate_ref.HostRateRef(emu.frame)
    hr.gain_in.target_db = -6.0
    out = []
    for i in range(30):
        if i == 4:
            hr.gain_out.target_db = 3.0
        if i == 8:
            emu.set_speaker(5)
        if i == 15:
            emu.s.set_formant_index(int(round(1.0 * 2 + 4)))
        if i == 20:
            emu.corr = 0.4
        if i == 22:
            emu.inton = 1.5
        if i == 24:
            emu.avg = 60.0
        if i == 26:
            emu.s.set_vq(3)
        out.append(hr.process(x[i * 480:(i + 1) * 480]))
    assert np.array_equal(y, np.concatenate(out))
    raw, used = zip(*emu.q_log[:20])
    assert all(u == min(r + 96, 447) for r, u in zip(raw, used))


def test_block_size_does_not_change_the_stream(model_dir):
    x = signals.voice_like(480 * 12, 48000.0, 9)
    y480, _ = callsite.run("oracle", _toml(model_dir), x, block=480)
    y333, _ = callsite.run("oracle", _toml(model_dir), x, block=333)
    y64, _ = callsite.run("oracle", _toml(model_dir), x, block=64)
    assert np.array_equal(y480, y333) and np.array_equal(y480, y64)


def test_reset_context_restarts_model_state_only(model_dir):
    x = signals.voice_like(480 * 10, 48000.0, 4)
    y, info = callsite.run("oracle", _toml(model_dir), x, events=[(5, "reset", 1)])
    y0, _ = callsite.run("oracle", _toml(model_dir), x)
    assert info["last"] == 0
    assert np.array_equal(y[:480 * 6], y0[:480 * 6])     # output of hop 5 still comes from the FIFO
    assert not np.array_equal(y[480 * 6:], y0[480 * 6:])


def test_speaker_out_of_range_and_bad_correction_type(model_dir):
    x = signals.voice_like(480 * 2, 48000.0, 4)
    # 8 speakers -> ids 0..8 valid (8 = morph slot); 9 is rejected, previous speaker stays
    y, info = callsite.run("oracle", _toml(model_dir), x, events=[(0, "voice", 9)])
    y0, _ = callsite.run("oracle", _toml(model_dir), x)
    assert np.array_equal(y, y0)


def test_committed_callsite_golden_is_reproduced(model_dir):
    g = np.load(os.path.join(ROOT, "tests", "golden", "callsite_48k.npz"))
    events = [tuple(e) for e in json.loads(str(g["events"]))]
    y, info = callsite.run("oracle", _toml(model_dir), g["x"], events=events)
    assert info["load"] == 0
    assert np.array_equal(y, g["y"])


@pytest.mark.skipif(not callsite.available("stub"), reason="oracle/_ref/callsite_runner_stub not built")
def test_stub_echo_runner_agrees_with_the_numpy_adapter(model_dir):
    """The reference call site over the stub library in echo mode (the known-answer source of tests/test_gpu_anyrate.py)
    at 48 kHz / 480 equals the numpy restatement of the adapter around the same echo "model", bit for bit; and at another
    host rate it is a different, non-trivial signal of the block size's length."""
    x = signals.batch_48k(1, 12, seed0=5)[:, 0, :].reshape(-1)
    toml = os.path.join(model_dir, "model.toml")
    y, info = callsite.run("stub", toml, x, 48000.0, 480, echo=True)
    assert info["load"] == 0 and info["last"] == 0

    def echo(x16):
        out = np.zeros(240, np.float32)
        out[:160] = x16
        return out

    hr = hostrate_ref.HostRateRef(echo)
    want = np.concatenate([hr.process(x[h * 480:(h + 1) * 480]) for h in range(12)])
    assert np.array_equal(y, want) and y.std() > 0.01
    y2, _ = callsite.run("stub", toml, x[:4410], 44100.0, 441, echo=True)
    assert y2.shape == (4410,) and y2.std() > 0.01 and not np.array_equal(y2, y[:4410])
