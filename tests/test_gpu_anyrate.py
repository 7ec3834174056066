"""Device adapter for ANY host sample rate / block size (BeatriceB200_SetHostSampleRate + _ProcessAnyRate) against the
reference's own ProcessorCore2::Process (oracle/_ref, compiled from /root/reference by oracle/Makefile):

* BIT-EXACT: the reference call site over the stub library in echo mode (a "model" that returns its 160 input samples
  and 80 zeros) vs the device adapter around the same stand-in (BeatriceB200_SetEchoModel): gain_in with slews, the
  rational polyphase resampler both ways with its fraction clocks, the 480-sample block FIFO, gain_out --
  resample.h:25-46, :130-206, :209-230, :343-363, :384-394, gain.h:41-71.  array_equal over the whole run.
* END TO END: the real model behind it vs the reference call site over the CPU oracle, <= 1e-4 RMS (north_star).
"""
import os

import numpy as np
import pytest

import callsite
from beatrice_vst_b200 import batch as bbatch
from beatrice_vst_b200 import signals
from conftest import rms

pytestmark = pytest.mark.gpu
SETTER = dict(input_gain="InputGain", output_gain="OutputGain", pitch_shift="PitchShift", voice="TargetSpeaker")


def _blocks(total, sizes):
    """Cuts [0, total) into consecutive blocks cycling through `sizes`."""
    out, pos, i = [], 0, 0
    while pos < total:
        m = min(sizes[i % len(sizes)], total - pos)
        out.append((pos, m))
        pos += m
        i += 1
    return out


def _host_signal(n, samples, rate, seed0):
    """n streams of `samples` samples at `rate`: the 48 kHz test signal generator, read as if recorded at `rate`."""
    hops = (samples + 479) // 480
    x = signals.batch_48k(n, hops, seed0=seed0)            # [hops][n][480]
    return np.ascontiguousarray(x.transpose(1, 0, 2).reshape(n, -1)[:, :samples])


@pytest.mark.skipif(not callsite.available("stub"), reason="oracle/_ref/callsite_runner_stub not built")
@pytest.mark.parametrize("rate,block", [(48000.0, 480), (48000.0, 333), (44100.0, 441), (44100.0, 512), (96000.0, 1024),
                                        (88200.0, 100), (32000.0, 320), (22050.0, 256), (16000.0, 160), (47999.0, 64)])
def test_any_rate_adapter_is_bit_exact(product, model_dir, rate, block):
    n, seconds = 3, 0.35
    samples = int(rate * seconds)
    x = _host_signal(n, samples, rate, seed0=4100)
    # gain changes in mid-run: rising and falling slews of both gains (block index, name, dB)
    nb = (samples + block - 1) // block
    events = [(nb // 5, "input_gain", 6.0), (nb // 3, "output_gain", -9.0), (nb // 2, "input_gain", -3.0), (2 * nb // 3, "output_gain", 4.0)]
    events.insert(0, (-1, "output_gain", -4.0))      # set before the rate (and, in the reference, before LoadModel): the
    eng = bbatch.Engine(product, n)                  # Gain::Context survives SetSampleRate, so the first block slews to it
    assert eng.load(model_dir) == 0
    assert eng.set("OutputGain", -4.0, -1) == 0
    assert eng.set_host_sample_rate(rate) == 0
    assert eng.set_echo_model(True) == 0
    got = np.zeros_like(x)
    for bi, (pos, m) in enumerate(_blocks(samples, [block])):
        for (at, name, db) in events:
            if at == bi:
                assert eng.set(SETTER[name], db, -1) == 0
        got[:, pos:pos + m] = eng.process_any_rate(x[:, pos:pos + m])
    eng.close()
    toml = os.path.join(model_dir, "model.toml")
    for s in range(n):
        want, info = callsite.run("stub", toml, x[s], rate, block, events=events, echo=True)
        assert info["load"] == 0 and info["last"] == 0
        assert np.array_equal(got[s], want[:samples]), (rate, block, s, int(np.argmax(got[s] != want[:samples])))
    assert got.std() > 0.01


@pytest.mark.skipif(not callsite.available("stub"), reason="oracle/_ref/callsite_runner_stub not built")
def test_any_rate_adapter_ragged_blocks(product, model_dir):
    """Block sizes that change from call to call (1 sample up to several hops in one call): the result depends on the
    samples only, not on how they were cut -- except through the gain's dB round trip per call, so the reference is
    driven with the same cuts: one call-site run per distinct block size is not possible, hence no gain change here and
    the comparison is against a fixed-block run of the reference (resampler + FIFO are block-size independent)."""
    n, rate = 2, 44100.0
    samples = 9000
    x = _host_signal(n, samples, rate, seed0=4200)
    eng = bbatch.Engine(product, n)
    assert eng.load(model_dir) == 0
    assert eng.set_host_sample_rate(rate) == 0 and eng.set_echo_model(True) == 0
    got = np.zeros_like(x)
    for pos, m in _blocks(samples, [1, 7, 441, 2048, 64, 1500, 3]):
        got[:, pos:pos + m] = eng.process_any_rate(x[:, pos:pos + m])
    eng.close()
    toml = os.path.join(model_dir, "model.toml")
    for s in range(n):
        want, _ = callsite.run("stub", toml, x[s], rate, 441, echo=True)
        assert np.array_equal(got[s], want[:samples]), s


@pytest.mark.skipif(not callsite.available("oracle"), reason="oracle/_ref/callsite_runner_oracle not built")
@pytest.mark.parametrize("rate,block", [(44100.0, 441), (96000.0, 512)])
def test_any_rate_matches_reference_callsite(product, model_dir, rate, block):
    """The real model behind the any-rate adapter vs the reference call site over the CPU oracle."""
    n, seconds = 4, 0.6
    samples = int(rate * seconds) // block * block
    x = _host_signal(n, samples, rate, seed0=4300)
    eng = bbatch.Engine(product, n)
    assert eng.load(model_dir) == 0
    assert eng.set_host_sample_rate(rate) == 0
    for s in range(n):
        assert eng.set("TargetSpeaker", s % 8, s) == 0
    assert eng.set("PitchShift", 3.0, 1) == 0
    got = np.concatenate([eng.process_any_rate(x[:, pos:pos + m]) for pos, m in _blocks(samples, [block])], axis=1)
    eng.close()
    toml = os.path.join(model_dir, "model.toml")
    for s in range(n):
        ev = [(-1, "voice", s % 8)] + ([(-1, "pitch_shift", 3.0)] if s == 1 else [])
        want, info = callsite.run("oracle", toml, x[s], rate, block, events=ev)
        assert info["load"] == 0
        assert rms(got[s], want) <= 1e-4, (rate, s, rms(got[s], want))
    assert got.std() > 0.01


def test_any_rate_argument_checks(product, model_dir):
    eng = bbatch.Engine(product, 2)
    assert eng.load(model_dir) == 0
    x = np.zeros((2, 64), np.float32)
    with pytest.raises(RuntimeError):
        eng.process_any_rate(x)                       # no host rate set: the reference's kResamplerNotReady
    assert eng.set_host_sample_rate(0.0) != 0
    assert eng.set_host_sample_rate(-5.0) != 0
    assert eng.set_host_sample_rate(44100.0) == 0
    assert eng.set_host_sample_rate(44100.0) == 0     # same rate: no-op (processor_core_2.cc:422-424)
    assert not eng.process_any_rate(x).any()          # silence in, silence out
    with pytest.raises(RuntimeError):
        eng.process_any_rate(np.zeros((2, 5000), np.float32))   # block larger than the adapter's 4096
    eng.close()
