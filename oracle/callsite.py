"""TEST INFRASTRUCTURE ONLY -- driver for ``oracle/_ref/callsite_runner_*``.

Those executables are the reference's unmodified ``src/common`` call site
(``ProcessorProxy`` -> ``ProcessorCore2::Process``) compiled from
``/root/reference`` by ``oracle/Makefile`` and linked against either the CPU
oracle (``which="oracle"``) or the CUDA product library (``which="b200"``).
They run as a separate process so the reference's C++ never shares a symbol
namespace with Python extension modules (doing so crashed inside libstdc++).

``oracle/_ref/vst_harness_*`` (``run_vst``) go one layer further out: the reference's
unmodified VST3 processor ``src/vst/processor.cc`` plus the vendored vst3sdk, driven by a
headless host (``oracle/vst_harness.cc``) -- SURVEY.md section 8 row (f-2).
"""
from __future__ import annotations

import json
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# reference src/common/error.h:11-25
ERROR_NAMES = ["kSuccess", "kFileOpenError", "kFileTooSmall", "kFileTooLarge", "kInvalidFileSize",
               "kTOMLSyntaxError", "kInvalidModelConfig", "kSpeakerIDOutOfRange",
               "kInvalidPitchCorrectionType", "kModelNotLoaded", "kResamplerNotReady", "kGainNotReady",
               "kUnknownError"]


def exe_path(which: str) -> str:
    return os.path.join(_HERE, "_ref", f"callsite_runner_{which}")


def available(which: str) -> bool:
    return os.path.exists(exe_path(which))


def run(which: str, toml_path, x: np.ndarray, sample_rate: float = 48000.0, block: int = 480,
        events=(), timeout: float = 600.0, echo: bool = False):
    """Feeds ``x`` through ProcessorCore::Process in ``block``-sample calls.

    ``events``: iterable of ``(block_index, name, value)``; ``block_index`` -1 applies the
    parameter before LoadModel, ``name="reset"`` calls ResetContext().  ``toml_path=None``
    leaves the proxy unloaded.  Returns ``(y, info)`` with info = dict(load, last, version).
    ``which="stub"`` with ``echo=True``: the call site over ``oracle/stub_beatricelib.cc`` whose "model" returns its
    160 input samples followed by 80 zeros -- bit-exact known answers for the host-rate adapter at any rate / block.
    """
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.f32"), os.path.join(d, "out.f32")
        np.ascontiguousarray(x, "<f4").tofile(fin)
        cmd = [exe_path(which), "run", toml_path or "-", fin, fout, repr(float(sample_rate)), str(int(block))]
        cmd += [f"{int(b)}:{n}={float(v)!r}" for b, n, v in events]
        env = dict(os.environ, STUB_ECHO="1" if echo else "0")
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        if p.returncode != 0:
            raise RuntimeError(f"callsite_runner failed ({p.returncode}): {p.stderr[-2000:]}")
        y = np.fromfile(fout, "<f4")
    info = {}
    for tok in p.stdout.split():
        if "=" in tok:
            k, v = tok.split("=")
            info[k] = int(v)
    return y, info


def vst_exe_path(which: str) -> str:
    return os.path.join(_HERE, "_ref", f"vst_harness_{which}")


def vst_available(which: str) -> bool:
    return os.path.exists(vst_exe_path(which))


def run_vst(which: str, toml_path, x: np.ndarray, sample_rate: float = 48000.0, block: int = 480,
            events=(), timeout: float = 600.0):
    """Same contract as :func:`run`, but through the reference's unmodified VST3 processor
    (``src/vst/processor.cc`` + vendored vst3sdk) driven by the headless host ``oracle/vst_harness.cc``:
    parameters travel as normalised values in ``IParameterChanges``, the model path as the controller's
    ``param_change`` message, audio through ``IAudioProcessor::process``.

    Returns ``(y, info)``; ``info["applied"]`` lists ``(name, plain_value)`` as the processor de-normalised
    them (feed these to :func:`run` for a bit-exact comparison), ``info["load"]`` / ``info["process"]`` are
    the ``tresult`` codes (0 = kResultOk).
    """
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.f32"), os.path.join(d, "out.f32")
        np.ascontiguousarray(x, "<f4").tofile(fin)
        cmd = [vst_exe_path(which), "run", toml_path or "-", fin, fout, repr(float(sample_rate)), str(int(block))]
        cmd += [f"{int(b)}:{n}={float(v)!r}" for b, n, v in events]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError(f"vst_harness failed ({p.returncode}): {p.stderr[-2000:]}")
        y = np.fromfile(fout, "<f4")
    info = {"applied": []}
    for line in p.stdout.splitlines():
        if line.startswith("applied "):
            k, v = line[len("applied "):].split("=")
            info["applied"].append((k, float(v)))
        elif line.startswith("status:"):
            for tok in line.split()[1:]:
                k, v = tok.split("=")
                info[k] = int(v)
    return y, info


def bench(which: str, toml_path: str, signal: np.ndarray, warmup: int = 10, timeout: float = 1800.0):
    """signal [n_threads, n_frames, 480] -> dict(frames_per_s, threads, ...)."""
    signal = np.ascontiguousarray(signal, "<f4")
    nt, nf, hop = signal.shape
    assert hop == 480
    with tempfile.TemporaryDirectory() as d:
        fsig = os.path.join(d, "sig.f32")
        signal.tofile(fsig)
        p = subprocess.run([exe_path(which), "bench", toml_path, fsig, str(nt), str(nf), str(warmup)],
                           capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"callsite_runner bench failed ({p.returncode}): {p.stderr[-2000:]}")
    return json.loads(p.stdout.strip().splitlines()[-1])
