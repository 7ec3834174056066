"""GPU parity tests (``-m gpu``): the CUDA library through its C ABI against the CPU oracle on
identical seeded inputs, against the committed golden vectors, and -- through the reference's
own call site -- as a drop-in.  Tolerance: output audio <= 1e-4 RMS (fp32), the bar
BASELINE.json's north_star states; integer outputs (pitch bins) exact.

Nothing here reads /root/reference; the call site is the prebuilt oracle/_ref runner."""
import ctypes as C
import json
import os
import threading

import numpy as np
import pytest

import callsite
from beatrice_vst_b200 import batch as bbatch
from beatrice_vst_b200 import lib as blib
from beatrice_vst_b200 import signals
from conftest import ROOT, rms

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")
TOL_WAVE = 1e-4     # north_star: <= 1e-4 RMS (fp32)
TOL_FEAT = 2e-5     # fp32 mode: phone / pitch-feature vectors (values of magnitude ~1..4)
TOL_FEAT_TC = 1e-4  # split-bf16 on tcgen05 (the default): ~1e-5 relative on the same vectors


def _tol_feat(precision):
    return TOL_FEAT if precision == 0 else TOL_FEAT_TC


@pytest.fixture(params=[2, 0], ids=["bf16x3", "f32"])
def abi_precision(request, product):
    """Arithmetic of the per-stream contexts behind beatrice.h: the default (split-bf16 on tcgen05) and the opt-in
    fp32 CUDA-core mode.  Contexts build on their first per-frame call, so the switch is set around the test."""
    bbatch.set_default_precision(product, request.param)
    yield request.param
    bbatch.set_default_precision(product, -1)


def _pair(product, oracle, model_dir, family=2, **kw):
    a = blib.SingleStream(product, model_dir, family=family, **kw)
    b = blib.SingleStream(oracle, model_dir, family=family, **kw)
    assert a.ok and b.ok, (a.errors, b.errors)
    return a, b


def test_native_library_is_the_one_loaded(product):
    assert bbatch.device_count(product) >= 1
    assert any("libbeatrice_b200.so" in l for l in open("/proc/self/maps"))


@pytest.mark.parametrize("family", [2, 0, 1])
def test_single_stream_abi_matches_oracle(product, oracle, model_dirs, family, abi_precision):
    x = signals.voice_like(160 * 40, 16000.0, seed=21)
    a, b = _pair(product, oracle, model_dirs[family], family, speaker=3, formant_index=6)
    a.set_pitch_range(1, 383)
    b.set_pitch_range(1, 383)
    pa, qa, fa, wa = a.run(x)
    pb, qb, fb, wb = b.run(x)
    a.close()
    b.close()
    assert np.array_equal(qa, qb)
    assert rms(pa, pb) <= _tol_feat(abi_precision) and rms(fa, fb) <= _tol_feat(abi_precision), (rms(pa, pb), rms(fa, fb))
    assert rms(wa, wb) <= TOL_WAVE, rms(wa, wb)
    assert np.abs(wa - wb).max() <= 1e-3


@pytest.mark.parametrize("family", [0, 2])
def test_single_stream_abi_matches_golden(product, model_dirs, family, abi_precision):
    g = np.load(os.path.join(GOLDEN, f"m0_family{family}.npz"))
    s = blib.SingleStream(product, model_dirs[family], family=family, speaker=1, formant_index=5)
    s.set_pitch_range(1, 383)
    phone, q, feat, wave = s.run(g["x"])
    s.close()
    assert np.array_equal(q, g["q"])
    assert rms(phone, g["phone"]) <= _tol_feat(abi_precision) and rms(wave, g["wave"]) <= TOL_WAVE
    if family == 2:
        s = blib.SingleStream(product, model_dirs[2], family=2, speaker=1, formant_index=5)
        s.set_pitch_range(1, 383)
        s.set_vq(4)
        p2, _, _, w2 = s.run(g["x"])
        s.close()
        assert rms(p2, g["phone_vq4"]) <= _tol_feat(abi_precision) and rms(w2, g["wave_vq4"]) <= TOL_WAVE


def test_vocoder_stage_taps_match_oracle(product, oracle, model_dir):
    """Per-stage activations of the vocoder (hidden, pre, 4 stage outputs); fp32 mode, where every intermediate
    lives in an fp32 ring (the tensor-core modes keep hidden / pre inside the fused chain kernel)."""
    x = signals.voice_like(160 * 6, 16000.0, seed=4)
    bbatch.set_default_precision(product, 0)
    a, b = _pair(product, oracle, model_dir)
    a.run(x)
    b.run(x)
    product.dll.BeatriceB200_WaveformTap.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int]
    oracle.dll.BeatriceOracle_WaveformTap.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int]
    for which, n in enumerate([256, 256, 5 * 128, 20 * 64, 80 * 32, 240 * 16]):
        ta, tb = np.zeros(n, np.float32), np.zeros(n, np.float32)
        fp = lambda v: v.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
        assert product.dll.BeatriceB200_WaveformTap(a.wc, which, fp(ta), n) == n
        assert oracle.dll.BeatriceOracle_WaveformTap(b.wc, which, fp(tb), n) == n
        scale = max(float(np.sqrt(np.mean(tb ** 2))), 1e-6)
        assert rms(ta, tb) / scale <= 2e-5, (which, rms(ta, tb), scale)
    a.close()
    b.close()
    bbatch.set_default_precision(product, -1)


def test_vq_speaker_switch_and_formant(product, oracle, model_dir, abi_precision):
    """kNN-VQ on/off, set-speaker with the key-value blocks one per hop, formant change."""
    x = signals.voice_like(160 * 30, 16000.0, seed=8).reshape(30, 160)
    a, b = _pair(product, oracle, model_dir)
    outs = []
    for s in (a, b):
        w = []
        for i in range(30):
            if i == 5:
                s.set_vq(8)
            if i == 10:
                s.set_speaker(6, kv_blocks_now=False)
            if 10 <= i < 14:
                s.set_kv_block(i - 10)
            if i == 18:
                s.set_formant_index(0)
            if i == 22:
                s.set_vq(0)
                s.set_pitch_range(50, 200)
            w.append(s.frame(x[i]))
        outs.append(w)
    for (pa, qa, fa, wa), (pb, qb, fb, wb) in zip(*outs):
        assert qa == qb
        assert rms(pa, pb) <= _tol_feat(abi_precision) and rms(wa, wb) <= TOL_WAVE
    a.close()
    b.close()


def test_context_recreation_gives_fresh_state(product, model_dir):
    """ResetContext destroys and re-creates contexts (processor_core_2.cc:258-266)."""
    x = signals.voice_like(160 * 8, 16000.0, seed=3)
    s1 = blib.SingleStream(product, model_dir)
    r1 = s1.run(x)
    s1.close()
    s2 = blib.SingleStream(product, model_dir)
    r2 = s2.run(x)
    s2.close()
    assert np.array_equal(r1[3], r2[3]) and np.array_equal(r1[1], r2[1])


def test_independent_instances_on_threads(product, oracle, model_dir):
    """DAWs run many plug-in instances on different threads (SURVEY.md 8b threading)."""
    n = 4
    xs = [signals.voice_like(160 * 12, 16000.0, seed=30 + i) for i in range(n)]
    res = [None] * n

    def work(i):
        s = blib.SingleStream(product, model_dir, speaker=i)
        res[i] = s.run(xs[i])
        s.close()

    th = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(n):
        o = blib.SingleStream(oracle, model_dir, speaker=i)
        ref = o.run(xs[i])
        o.close()
        assert np.array_equal(res[i][1], ref[1])
        assert rms(res[i][3], ref[3]) <= TOL_WAVE


# ---------------------------------------------------------------------------------------
# batched engine
# ---------------------------------------------------------------------------------------
def _oracle_stream(oracle, model_dir, x, speaker=0, formant_index=4, vq=0, shift_bins=0, lo=1, hi=383):
    s = blib.SingleStream(oracle, model_dir, speaker=speaker, formant_index=formant_index)
    s.set_pitch_range(lo, hi)
    s.set_vq(vq)
    out, qs = [], []
    for i in range(len(x) // 160):
        xi = x[i * 160:(i + 1) * 160]
        # run the encoders once to learn q, then the vocoder with the transformed bin
        f = s.f
        fp = lambda v: v.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
        xi = np.ascontiguousarray(xi, np.float32)
        ph, ft, w = np.empty(128, np.float32), np.empty(4, np.float32), np.empty(240, np.float32)
        q = C.c_int(0)
        f("ExtractPhone1")(s.pe, fp(xi), fp(ph), s.pc)
        f("EstimatePitch1")(s.pi, fp(xi), C.byref(q), fp(ft), s.pic)
        qq = C.c_int(min(max(q.value + shift_bins, 1), 447))
        f("GenerateWaveform1")(s.wg, fp(ph), C.byref(qq), fp(ft), fp(w), s.wc)
        out.append(w)
        qs.append((q.value, qq.value))
    s.close()
    return np.stack(out), qs


@pytest.mark.parametrize("engine_precision", [2, 0], ids=["bf16x3", "f32"])
def test_batched_frames_match_oracle_per_stream(product, oracle, model_dir, engine_precision):
    n, hops = 6, 12
    xs = signals.batch_16k(n, hops, seed0=100)
    eng = bbatch.Engine(product, n, precision=engine_precision)
    assert eng.load(model_dir) == 0 and eng.n_speakers == 8
    spk = [0, 3, 7, 1, 2, 5]
    shift = [0.0, 12.0, -12.0, 3.0, 0.0, 24.0]
    fidx = [4, 0, 8, 4, 6, 2]
    vq = [0, 0, 4, 0, 8, 0]
    for s in range(n):
        assert eng.set("TargetSpeaker", spk[s], s) == 0
        assert eng.set("PitchShift", shift[s], s) == 0
        assert eng.set("FormantShift", (fidx[s] - 4) / 2.0, s) == 0
        assert eng.set("VQNumNeighbors", vq[s], s) == 0
    eng.reset_stream(-1)   # applies all four key-value blocks at once, like ResetContext
    got = np.stack([eng.process_frames(xs[h]).copy() for h in range(hops)], axis=1)   # [n][hops][240]
    _, q_raw, q_used, _ = eng.last_intermediates()
    for s in range(n):
        ref, qs = _oracle_stream(oracle, model_dir, xs[:, s, :].reshape(-1), spk[s], fidx[s], vq[s],
                                 int(round(shift[s] * 8)))
        assert (q_raw[s], q_used[s]) == qs[-1]
        assert rms(got[s], ref) <= TOL_WAVE, (s, rms(got[s], ref))
    assert eng.kernel_launches() > 0
    eng.close()


def test_batched_rejects_bad_arguments(product, model_dir):
    eng = bbatch.Engine(product, 2)
    assert eng.set("PitchShift", 1.0, 0) == 9            # kModelNotLoaded before LoadModel
    assert eng.load(model_dir) == 0
    assert eng.set("TargetSpeaker", 8, 0) == 0           # n_speakers = the morphing slot (processor_core_2.cc:436)
    assert eng.set("TargetSpeaker", 9, 0) == 7           # beyond it -> kSpeakerIDOutOfRange
    assert eng.set("TargetSpeaker", -1, 0) == 7
    assert eng.set("PitchCorrectionType", 2, 0) == 8     # kInvalidPitchCorrectionType
    assert eng.set("PitchShift", 1.0, 5) == -1           # no such stream
    assert eng.load(os.path.join(model_dir, "nope")) == 1
    eng.close()


@pytest.mark.parametrize("engine_precision", [2, 0], ids=["bf16x3", "f32"])
@pytest.mark.skipif(not callsite.available("oracle"), reason="oracle/_ref not built")
def test_batched_48k_matches_reference_callsite(product, model_dir, engine_precision):
    """Process48k == ProcessorCore2::Process at 48 kHz / 480-sample blocks, per stream, incl.
    gain slews, pitch shift, correction and a speaker change (4-hop key-value schedule)."""
    n, hops = 3, 24
    x = signals.batch_48k(n, hops, seed0=40)
    plans = [
        [(-1, "input_gain", -6.0), (-1, "pitch_shift", 7.0), (5, "output_gain", 3.0), (9, "voice", 4)],
        [(-1, "pitch_correction", 0.5), (3, "formant_shift", -1.5), (12, "vq_num_neighbors", 4)],
        [(-1, "pitch_correction_type", 1), (-1, "pitch_correction", 0.3), (7, "intonation_intensity", 0.5),
         (10, "min_source_pitch", 45.0), (10, "max_source_pitch", 70.0), (15, "voice", 2)],
    ]
    setter = dict(input_gain="InputGain", output_gain="OutputGain", pitch_shift="PitchShift", voice="TargetSpeaker",
                  pitch_correction="PitchCorrection", formant_shift="FormantShift", vq_num_neighbors="VQNumNeighbors",
                  pitch_correction_type="PitchCorrectionType", intonation_intensity="IntonationIntensity",
                  min_source_pitch="MinSourcePitch", max_source_pitch="MaxSourcePitch")
    eng = bbatch.Engine(product, n, precision=engine_precision)
    assert eng.load(model_dir) == 0
    out = np.empty((n, hops, 480), np.float32)
    for h in range(hops):
        for s in range(n):
            for (b, name, v) in plans[s]:
                if b == h or (b == -1 and h == 0):
                    val = int(v) if name in ("voice", "vq_num_neighbors", "pitch_correction_type") else float(v)
                    assert eng.set(setter[name], val, s) == 0
        out[:, h, :] = eng.process_48k(x[h])
    eng.close()
    for s in range(n):
        y, info = callsite.run("oracle", os.path.join(model_dir, "model.toml"), x[:, s, :].reshape(-1),
                               events=plans[s])
        assert info["load"] == 0 and info["last"] == 0
        e = rms(out[s].reshape(-1), y)
        assert e <= TOL_WAVE, (s, e)


@pytest.mark.parametrize("family", [0, 1])
@pytest.mark.skipif(not callsite.available("oracle"), reason="oracle/_ref not built")
def test_batched_48k_legacy_families_match_reference_callsite(product, model_dirs, family):
    """The batched engine for the 2.0.0-alpha.2 / -beta.1 model families (256 phone channels, 384 pitch bins, the
    speaker vector = speaker embedding + formant embedding formed per call, no embedding setter / VQ / key-value
    path) against ProcessorCore0 / ProcessorCore1 (reference src/common/processor_core_{0,1}.cc:24-142) over the
    CPU oracle: speaker and formant changes, pitch shift / correction, gains; both pipeline depths."""
    n, hops = 4, 22
    x = signals.batch_48k(n, hops, seed0=640 + family)
    plans = [
        [(0, "voice", 3), (0, "pitch_shift", 5.0), (6, "formant_shift", 1.5), (11, "voice", 6)],
        [(0, "voice", 1), (0, "input_gain", -4.0), (4, "pitch_correction", 0.6), (9, "output_gain", 3.0)],
        [(0, "formant_shift", -2.0), (7, "reset", 1), (12, "voice", 7), (12, "pitch_shift", -7.0)],
        [(3, "pitch_correction_type", 1), (3, "pitch_correction", 0.35), (8, "intonation_intensity", 1.4)],
    ]
    setter = dict(input_gain="InputGain", output_gain="OutputGain", pitch_shift="PitchShift", voice="TargetSpeaker",
                  pitch_correction="PitchCorrection", formant_shift="FormantShift", pitch_correction_type="PitchCorrectionType",
                  intonation_intensity="IntonationIntensity")
    toml = os.path.join(model_dirs[family], "model.toml")
    refs = []
    for s in range(n):
        y, info = callsite.run("oracle", toml, x[:, s, :].reshape(-1), events=plans[s])
        assert info == {"load": 0, "last": 0, "version": family}
        refs.append(y)
    for depth in (1, 2):
        eng = bbatch.Engine(product, n)
        assert eng.load(model_dirs[family]) == 0 and eng.family == family and eng.n_speakers == 8
        assert eng.set("TargetSpeaker", 8, 0) == 7            # no morphing slot in the legacy engine
        assert eng.set_pipeline_depth(depth) == 0
        outs = []
        for h in range(hops):
            for s in range(n):
                for (b, name, v) in plans[s]:
                    if b != h:
                        continue
                    if name == "reset":
                        assert eng.reset_stream(s) == 0
                    else:
                        assert eng.set(setter[name], int(v) if name in ("voice", "pitch_correction_type") else float(v), s) == 0
            outs.append(eng.process_48k(x[h]).copy())
        if depth == 2:
            outs = outs[1:] + [eng.drain()]
        eng.close()
        got = np.stack(outs, axis=1)
        for s in range(n):
            e = rms(got[s].reshape(-1), refs[s])
            assert refs[s].std() > 0.01 and e <= TOL_WAVE, (family, depth, s, e)


def test_48k_host_and_device_entries_interleave(product, model_dir):
    """The host-buffer entry computes the block it hands back on a side stream, with the hop index mirrored on
    the host; the device-buffer entry does it inside the hop graph with the device counter.  Alternating the
    two must give exactly what either gives alone (same ring slots, same counters), gain slew included."""
    n, hops = 2, 14
    x = signals.batch_48k(n, hops, seed0=90)
    ref_eng = bbatch.Engine(product, n)
    assert ref_eng.load(model_dir) == 0
    mix = bbatch.Engine(product, n)
    assert mix.load(model_dir) == 0
    d_in, d_out = mix.dev_alloc("t_in48", n * 480), mix.dev_alloc("t_out48", n * 480)
    for h in range(hops):
        if h == 3:
            for e in (ref_eng, mix):
                assert e.set("OutputGain", -4.0, 1) == 0 and e.set("InputGain", 2.0, 0) == 0
        want = ref_eng.process_48k(x[h]).copy()
        if h % 3 == 1:
            mix.to_device(d_in, x[h])
            assert mix.process_48k_device(d_in, d_out) == 0
            mix.synchronize()
            got = mix.to_host(d_out, (n, 480))
        else:
            got = mix.process_48k(x[h]).copy()
        assert np.array_equal(got, want), h
    assert want.std() > 0.01
    ref_eng.close()
    mix.close()


def test_batched_48k_matches_committed_callsite_golden(product, model_dir):
    g = np.load(os.path.join(GOLDEN, "callsite_48k.npz"))
    events = [tuple(e) for e in json.loads(str(g["events"]))]
    setter = dict(input_gain="InputGain", output_gain="OutputGain", pitch_shift="PitchShift", voice="TargetSpeaker",
                  pitch_correction="PitchCorrection", formant_shift="FormantShift")
    x = g["x"].reshape(-1, 480)
    eng = bbatch.Engine(product, 1)
    assert eng.load(model_dir) == 0
    out = []
    for h in range(len(x)):
        for (b, name, v) in events:
            if b == h or (b == -1 and h == 0):
                assert eng.set(setter[name], int(v) if name == "voice" else float(v), 0) == 0
        out.append(eng.process_48k(x[h][None, :])[0].copy())
    eng.close()
    assert rms(np.concatenate(out), g["y"]) <= TOL_WAVE


@pytest.mark.skipif(not (callsite.available("oracle") and callsite.available("b200")), reason="oracle/_ref not built")
def test_dropin_through_reference_callsite(model_dir):
    """The reference's unmodified ProcessorProxy/ProcessorCore2 linked against the CUDA library (its default
    arithmetic: split-bf16 on tcgen05) produces the same audio as the same call site linked against the CPU oracle."""
    x = signals.voice_like(480 * 40, 48000.0, seed=77)
    events = [(-1, "pitch_shift", -5.0), (6, "voice", 2), (11, "vq_num_neighbors", 4), (20, "reset", 1),
              (25, "formant_shift", 2.0)]
    toml = os.path.join(model_dir, "model.toml")
    ya, ia = callsite.run("b200", toml, x, events=events, block=333)
    yb, ib = callsite.run("oracle", toml, x, events=events, block=333)
    assert ia == ib == {"load": 0, "last": 0, "version": 2}
    assert rms(ya, yb) <= TOL_WAVE, rms(ya, yb)


@pytest.mark.skipif(not (callsite.vst_available("oracle") and callsite.vst_available("b200")),
                    reason="oracle/_ref VST harness not built")
def test_dropin_through_reference_vst_processor(model_dir):
    """the reference's unmodified VST3 processor (src/vst/processor.cc + vst3sdk), driven by a
    headless host through IAudioProcessor::process with parameter queues and the model-load message, produces
    the same audio linked against the CUDA library as linked against the CPU oracle."""
    x = signals.voice_like(480 * 40, 48000.0, seed=78)
    events = [(-1, "pitch_shift", 4.0), (5, "voice", 3), (12, "vq_num_neighbors", 4), (19, "reset", 1),
              (24, "formant_shift", -1.5), (30, "output_gain", -2.0)]
    toml = os.path.join(model_dir, "model.toml")
    ya, ia = callsite.run_vst("b200", toml, x, events=events, block=256)
    yb, ib = callsite.run_vst("oracle", toml, x, events=events, block=256)
    assert ia["load"] == ib["load"] == 0 and ia["process"] == ib["process"] == 0
    assert ia["applied"] == ib["applied"]
    assert yb.std() > 0.01 and rms(ya, yb) <= TOL_WAVE, rms(ya, yb)


def test_full_size_batch_properties(product, model_dir):
    """256 streams (BASELINE.json configs[1]): determinism, stream independence (a stream's
    output does not depend on what the other 255 carry) and equality with the batch-of-1 path."""
    n, hops = 256, 4
    xs = signals.batch_16k(8, hops, seed0=500)
    xs = np.tile(xs, (1, n // 8, 1))                      # stream s carries signal s % 8
    eng = bbatch.Engine(product, n)
    assert eng.load(model_dir) == 0
    got = np.stack([eng.process_frames(xs[h]).copy() for h in range(hops)], axis=1)
    eng.close()
    assert np.isfinite(got).all() and got.std() > 0.05
    for s in range(8, n):
        assert np.array_equal(got[s], got[s % 8]), s      # same input -> bitwise same output
    eng = bbatch.Engine(product, n)
    eng.load(model_dir)
    xs2 = xs.copy()
    xs2[:, 8:, :] = signals.batch_16k(n - 8, hops, seed0=900)
    got2 = np.stack([eng.process_frames(xs2[h]).copy() for h in range(hops)], axis=1)
    eng.close()
    assert np.array_equal(got2[:8], got[:8])              # independence
    one = blib.SingleStream(product, model_dir)
    one.set_pitch_range(1, 383)
    _, _, _, w = one.run(xs[:, 3, :].reshape(-1))
    one.close()
    assert rms(w, got[3]) <= 1e-6


# ---------------------------------------------------------------------------------------
# tensor-core (tcgen05) precisions of the batched engine
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol", [(2, TOL_WAVE), (1, 2e-2)])
def test_tensor_core_precisions_against_oracle(product, oracle, model_dir, precision, tol):
    """precision 2 = split-bf16 tcgen05 (hi/lo operands, fp32 accumulate) must still meet the
    1e-4 RMS bar; precision 1 = plain bf16 operands on the vocoder convs is reported against a
    looser, documented bound (bf16 has 8 mantissa bits).  The pitch bins stay exact in both:
    plain bf16 keeps the encoders in fp32."""
    n, hops = 5, 10
    xs = signals.batch_16k(n, hops, seed0=300)
    eng = bbatch.Engine(product, n, precision=precision)
    assert eng.load(model_dir) == 0
    for s in range(n):
        eng.set("TargetSpeaker", s, s)
    eng.reset_stream(-1)
    got = np.stack([eng.process_frames(xs[h]).copy() for h in range(hops)], axis=1)
    _, q_raw, _, _ = eng.last_intermediates()
    eng.close()
    worst = 0.0
    for s in range(n):
        ref, qs = _oracle_stream(oracle, model_dir, xs[:, s, :].reshape(-1), speaker=s)
        assert q_raw[s] == qs[-1][0]
        worst = max(worst, rms(got[s], ref))
    assert worst <= tol, worst


@pytest.mark.parametrize("precision,tol", [(2, TOL_WAVE), (1, 2e-2)])
def test_fused_mrf_stream_groups_and_reset(product, oracle, model_dir, precision, tol):
    """The fused MRF kernel keeps its conv histories per GROUP of streams (6 / 3 / 1 streams per
    CTA): 20 streams leave partial groups in two stages, and resetting ONE stream mid-run must
    clear exactly that stream's history (ResetContext semantics, processor_core_2.cc:258-266)."""
    n, hops, reset_at, victim = 20, 8, 5, 7
    xs = signals.batch_16k(n, hops, seed0=700)
    eng = bbatch.Engine(product, n, precision=precision)
    assert eng.load(model_dir) == 0
    got = []
    for h in range(hops):
        if h == reset_at:
            eng.reset_stream(victim)
        got.append(eng.process_frames(xs[h]).copy())
    got = np.stack(got, axis=1)
    eng.close()
    worst = 0.0
    for s in (0, 5, 6, 7, 8, 17, 18, 19):
        if s == victim:
            a, _ = _oracle_stream(oracle, model_dir, xs[:reset_at, s, :].reshape(-1))
            b, _ = _oracle_stream(oracle, model_dir, xs[reset_at:, s, :].reshape(-1))
            ref = np.concatenate([a, b])
        else:
            ref, _ = _oracle_stream(oracle, model_dir, xs[:, s, :].reshape(-1))
        worst = max(worst, rms(got[s], ref))
    assert worst <= tol, worst


def test_reset_all_streams_and_cluster_determinism(product, oracle, model_dir):
    """ResetStream(-1) (one memset per arena) must leave every stream as a fresh context, and the
    K-split cluster kernels (DSMEM reduce-scatter, rank-ordered sums) must be bitwise reproducible:
    40 streams span three 16-stream cluster groups of vocoder stage 0, the last one partial."""
    n, hops = 40, 5
    xs = signals.batch_16k(n, hops, seed0=900)
    eng = bbatch.Engine(product, n, precision=2)
    assert eng.load(model_dir) == 0
    first = np.stack([eng.process_frames(xs[h]).copy() for h in range(hops)], axis=1)
    eng.reset_stream(-1)
    again = np.stack([eng.process_frames(xs[h]).copy() for h in range(hops)], axis=1)
    eng.close()
    assert np.array_equal(first, again)          # fresh state + identical arithmetic, bit for bit
    for s in (0, 15, 16, 31, 32, 39):
        ref, _ = _oracle_stream(oracle, model_dir, xs[:, s, :].reshape(-1))
        assert rms(first[s], ref) <= TOL_WAVE, (s, rms(first[s], ref))


@pytest.mark.parametrize("form", [0, 1])
@pytest.mark.parametrize("plan", ["11|7,3;11@128/7|3;11,7,3", "7,3|11;3|7|11;11/7/3"])
def test_mrf_launch_plans_do_not_change_a_sample(product, model_dir, plan, form):
    """The fused MRF stages can be cut into launches and CTA classes in other ways than the default "11|7|3" (branch
    lists run back to back in one CTA, a stage as several launches chained as a PDL pair, a minimum shared-memory request:
    `MrfLaunchPlan`, BEATRICE_B200_MRF_PLAN, read when an engine loads).  Which CTA runs a branch must not change its
    arithmetic: every sample is bit-identical to the default plan, with either upsampler form (own launches / in the
    prologue of the fused kernel, where a CTA with several branches computes the upsampler once per branch)."""
    n, hops = 20, 6
    xs = signals.batch_16k(n, hops, seed0=1500)

    def run(env_plan):
        old = os.environ.pop("BEATRICE_B200_MRF_PLAN", None)
        if env_plan:
            os.environ["BEATRICE_B200_MRF_PLAN"] = env_plan
        try:
            eng = bbatch.Engine(product, n, precision=2)
            assert eng.load(model_dir) == 0
        finally:
            os.environ.pop("BEATRICE_B200_MRF_PLAN", None)
            if old is not None:
                os.environ["BEATRICE_B200_MRF_PLAN"] = old
        assert eng.set_upsampler_form(form) == 0
        out = np.stack([eng.process_frames(xs[h]).copy() for h in range(hops)], axis=1)
        eng.close()
        return out

    base, other = run(None), run(plan)
    assert base.std() > 0.01
    assert np.array_equal(base, other)
