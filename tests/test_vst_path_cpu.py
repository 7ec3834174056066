"""CPU: the reference's UNMODIFIED VST3 processor (src/vst/processor.cc + vendored vst3sdk, compiled in place
into oracle/_ref/vst_harness_*) driven headlessly through IComponent / IConnectionPoint / IAudioProcessor --
SURVEY.md section 8 row (f-2).  The audio must be bit-identical to the src/common call site fed the same plain
parameter values: the VST layer only normalises parameters, mixes the bus down to mono and forwards blocks."""
import os

import numpy as np
import pytest

import callsite
from beatrice_vst_b200 import signals

pytestmark = pytest.mark.skipif(not (callsite.vst_available("oracle") and callsite.available("oracle")),
                                reason="oracle/_ref not built (needs /root/reference at build time)")


def _toml(model_dir):
    return os.path.join(model_dir, "model.toml")


def test_vst_process_equals_callsite_bit_exact(model_dir):
    x = signals.voice_like(480 * 20, 48000.0, 11)
    yv, iv = callsite.run_vst("oracle", _toml(model_dir), x)
    yc, ic = callsite.run("oracle", _toml(model_dir), x)
    assert iv["load"] == 0 and iv["process"] == 0 and ic["load"] == 0
    assert yv.any() and np.array_equal(yv, yc)


def test_vst_parameter_queues_and_reset(model_dir):
    """Normalised parameter points in IParameterChanges (before the model message and mid-stream), a host block
    size that is not a multiple of the hop, and setActive(false/true) -> ResetContext."""
    x = signals.voice_like(480 * 30, 48000.0, 12)
    events = [(-1, "pitch_shift", 7.0), (-1, "input_gain", -3.0), (3, "voice", 4), (9, "formant_shift", -1.0),
              (14, "reset", 1), (18, "pitch_correction", 0.5), (21, "output_gain", 2.0), (25, "vq_num_neighbors", 2)]
    yv, iv = callsite.run_vst("oracle", _toml(model_dir), x, events=events, block=333)
    assert iv["load"] == 0 and iv["process"] == 0
    applied = dict(iv["applied"])
    # what the processor de-normalised is what the call site gets as the plain value
    plain = [(b, n, (1 if n == "reset" else applied[n])) for b, n, _ in events]
    yc, ic = callsite.run("oracle", _toml(model_dir), x, events=plain, block=333)
    assert ic["load"] == 0 and ic["last"] == 0
    assert np.array_equal(yv, yc)
    # the normalisation round trip keeps these values (divisions of the schema), so the test is not vacuous
    assert applied["pitch_shift"] == 7.0 and applied["voice"] == 4 and abs(applied["pitch_correction"] - 0.5) < 1e-9
    y0, _ = callsite.run_vst("oracle", _toml(model_dir), x, block=333)
    assert not np.array_equal(yv, y0)


def test_vst_unloaded_and_silence(model_dir):
    x = signals.voice_like(480 * 4, 48000.0, 13)
    y, info = callsite.run_vst("oracle", None, x)           # no model message: the core stays unloaded
    assert info["process"] == 0 and not y.any()
    z = np.zeros(480 * 4, np.float32)
    y, info = callsite.run_vst("oracle", _toml(model_dir), z)   # silent input: processor.cc:205-218 skips the core
    assert info["load"] == 0 and not y.any()


def test_vst_bad_model_path_keeps_running(tmp_path):
    """processor.cc:274-300 keeps whatever the controller sent, even when the model cannot be loaded: the message is
    acknowledged, the core stays unloaded and the processor hands back silence instead of failing."""
    x = signals.voice_like(480 * 3, 48000.0, 14)
    y, info = callsite.run_vst("oracle", str(tmp_path / "missing" / "model.toml"), x)
    assert info["load"] == 0 and info["process"] == 0 and not y.any()
