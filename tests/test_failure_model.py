"""The library never aborts the host process (reference behaviour: the per-frame functions
return void and the call site zeroes the output on every guard, processor_core_2.cc:26-43,
processor_core.h:95-104).  A failure -- here: no usable CUDA device -- is latched, readable through
BeatriceB200_LastError, loaders return an error code, per-frame calls write silence.  There is still no CPU
fallback: nothing is computed.

Runs in a child process because the latch is library-wide."""
import os
import subprocess
import sys
import textwrap

import pytest

from beatrice_vst_b200 import batch as bbatch
from conftest import ROOT

CHILD = textwrap.dedent("""
    import sys
    sys.path.insert(0, {root!r})
    import numpy as np
    from beatrice_vst_b200 import batch as bbatch, lib as blib, signals
    product = blib.load_product()
    assert bbatch.last_error(product) == (0, "")
    s = blib.SingleStream(product, {model_dir!r})
    # host-side validation passed, the upload could not happen: Beatrice_kFileOpenError (1) from the four loaders
    # that need the device; the two host-only readers (ReadNSpeakers, ReadSpeakerEmbeddings) still succeed
    assert s.errors == [1, 1, 1, 0, 1, 0], s.errors
    code, text = bbatch.last_error(product)
    assert code == -101 and "no usable CUDA device" in text, (code, text)
    x = signals.voice_like(160 * 3, 16000.0, seed=1)
    phone, q, feat, wave = s.run(x)
    assert not phone.any() and not feat.any() and not wave.any() and (q == 1).all()
    s.set_speaker(1)            # rc0 setters: no-ops, no crash
    s.set_formant_index(2)
    s.close()
    assert bbatch.Engine.__init__ is not None
    try:
        bbatch.Engine(product, 4)
        raise SystemExit("CreateEngine must fail without a device")
    except RuntimeError:
        pass
    # the entries added later in round 2 reject a missing engine without touching the device
    dll = bbatch.bind(product)
    # (BAD_ARGUMENT, or DEVICE where the latched failure is looked at first)
    assert dll.BeatriceB200_SetHostSampleRate(None, 44100.0) in (-1, -2)
    assert dll.BeatriceB200_ProcessAnyRate(None, None, None, 16) in (-1, -2)
    assert dll.BeatriceB200_SetUpsamplerForm(None, 1) in (-1, -2)
    assert dll.BeatriceB200_SetSkipOps(None, b"wave.mrf") in (-1, -2)
    bbatch.clear_error(product)
    assert bbatch.last_error(product) == (0, "")
    print("child ok")
""")


def _run_child(model_dir, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    p = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT, model_dir=model_dir)], env=env,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.returncode, p.stdout[-2000:], p.stderr[-2000:])   # -6 would be SIGABRT
    assert "child ok" in p.stdout
    assert "ERROR (latched" in p.stderr                                             # loud, once, on stderr
    assert p.stderr.count("ERROR (latched") == 1


def test_no_device_latches_error_and_writes_silence(product, model_dir):
    if bbatch.device_count(product) > 0:
        pytest.skip("a GPU is present; the forced-invalid-device variant below covers this box")
    _run_child(model_dir, {})


@pytest.mark.gpu
def test_invalid_device_latches_error_and_writes_silence(product, model_dir):
    """Same on a GPU box with the device choice forced out of range."""
    _run_child(model_dir, {"BEATRICE_B200_DEVICE": "99"})
