"""ctypes binding of the batched engine (``include/beatrice_b200.h`` part 2).

Setter names and meaning mirror ``ProcessorCore2`` (reference
``src/common/processor_core_2.cc:431-590``); ``stream=-1`` addresses all streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .lib import BeatriceLib

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)
_vp = C.c_void_p

BATCH_SYMBOLS = {
    "BeatriceB200_DeviceCount": (C.c_int, []),
    "BeatriceB200_Version": (C.c_char_p, []),
    "BeatriceB200_SetDefaultPrecision": (None, [C.c_int]),
    "BeatriceB200_LastError": (C.c_int, []),
    "BeatriceB200_LastErrorString": (C.c_char_p, []),
    "BeatriceB200_ClearError": (None, []),
    "BeatriceB200_CreateEngine": (_vp, [C.c_int, C.c_int, C.c_int]),
    "BeatriceB200_DestroyEngine": (None, [_vp]),
    "BeatriceB200_LoadModel": (C.c_int, [_vp, C.c_char_p]),
    "BeatriceB200_LoadModelFromMemory": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    "BeatriceB200_NumSpeakers": (C.c_int, [_vp]),
    "BeatriceB200_NumStreams": (C.c_int, [_vp]),
    "BeatriceB200_ModelFamily": (C.c_int, [_vp]),
    "BeatriceB200_PhoneChannels": (C.c_int, [_vp]),
    "BeatriceB200_SetTargetSpeaker": (C.c_int, [_vp, C.c_int, C.c_int]),
    "BeatriceB200_SetSpeakerMorphingWeights": (C.c_int, [_vp, C.c_int, _f32p, C.c_int]),
    "BeatriceB200_SeedMorphLottery": (C.c_int, [_vp, C.c_uint]),
    "BeatriceB200_GetMorphState": (C.c_int, [_vp, C.c_int, _f32p, _f32p, _i32p]),
    "BeatriceB200_SetFormantShift": (C.c_int, [_vp, C.c_int, C.c_double]),
    "BeatriceB200_SetPitchShift": (C.c_int, [_vp, C.c_int, C.c_double]),
    "BeatriceB200_SetAverageSourcePitch": (C.c_int, [_vp, C.c_int, C.c_double]),
    "BeatriceB200_SetIntonationIntensity": (C.c_int, [_vp, C.c_int, C.c_double]),
    "BeatriceB200_SetPitchCorrection": (C.c_int, [_vp, C.c_int, C.c_double]),
    "BeatriceB200_SetPitchCorrectionType": (C.c_int, [_vp, C.c_int, C.c_int]),
    "BeatriceB200_SetMinSourcePitch": (C.c_int, [_vp, C.c_int, C.c_double]),
    "BeatriceB200_SetMaxSourcePitch": (C.c_int, [_vp, C.c_int, C.c_double]),
    "BeatriceB200_SetVQNumNeighbors": (C.c_int, [_vp, C.c_int, C.c_int]),
    "BeatriceB200_SetInputGain": (C.c_int, [_vp, C.c_int, C.c_double]),
    "BeatriceB200_SetOutputGain": (C.c_int, [_vp, C.c_int, C.c_double]),
    "BeatriceB200_ResetStream": (C.c_int, [_vp, C.c_int]),
    "BeatriceB200_ProcessFrames": (C.c_int, [_vp, _vp, _vp]),
    "BeatriceB200_ProcessFramesDevice": (C.c_int, [_vp, _vp, _vp]),
    "BeatriceB200_Process48k": (C.c_int, [_vp, _vp, _vp]),
    "BeatriceB200_Process48kDevice": (C.c_int, [_vp, _vp, _vp]),
    "BeatriceB200_Synchronize": (None, [_vp]),
    "BeatriceB200_SetPipelineDepth": (C.c_int, [_vp, C.c_int]),
    "BeatriceB200_PipelineDepth": (C.c_int, [_vp]),
    "BeatriceB200_SetPipelinePlan": (C.c_int, [_vp, C.c_char_p]),
    "BeatriceB200_SetUpsamplerForm": (C.c_int, [_vp, C.c_int]),
    "BeatriceB200_SetSkipOps": (C.c_int, [_vp, C.c_char_p]),
    "BeatriceB200_SetHostSampleRate": (C.c_int, [_vp, C.c_double]),
    "BeatriceB200_ProcessAnyRate": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "BeatriceB200_SetEchoModel": (C.c_int, [_vp, C.c_int]),
    "BeatriceB200_DrainPipeline": (C.c_int, [_vp, _vp, _vp]),
    "BeatriceB200_AllocPinned": (_vp, [C.c_size_t]),
    "BeatriceB200_FreePinned": (None, [_vp]),
    "BeatriceB200_AllocDevice": (_vp, [_vp, C.c_size_t]),
    "BeatriceB200_FreeDevice": (None, [_vp, _vp]),
    "BeatriceB200_CopyToDevice": (None, [_vp, _vp, _vp, C.c_size_t]),
    "BeatriceB200_CopyToHost": (None, [_vp, _vp, _vp, C.c_size_t]),
    "BeatriceB200_Stream": (_vp, [_vp]),
    "BeatriceB200_GetLastIntermediates": (C.c_int, [_vp, _f32p, _i32p, _i32p, _f32p]),
    "BeatriceB200_TransformPitchBins": (C.c_int, [_vp, _i32p, _i32p]),
    "BeatriceB200_AdapterOnly48k": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "BeatriceB200_ResidentBytes": (C.c_size_t, [_vp]),
    "BeatriceB200_KernelLaunchCount": (C.c_uint64, [_vp]),
    "BeatriceB200_ProfileHop": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int]),
    "BeatriceB200_WaveformTap": (C.c_int, [_vp, C.c_int, _f32p, C.c_int]),
}


class KernelRecord(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms", C.c_float), ("flops", C.c_double), ("bytes", C.c_double)]


def bind(lib: BeatriceLib):
    for name, (res, args) in BATCH_SYMBOLS.items():
        fn = getattr(lib.dll, name)
        fn.restype = res
        fn.argtypes = args
    return lib.dll


def device_count(lib: BeatriceLib) -> int:
    return bind(lib).BeatriceB200_DeviceCount()


def set_default_precision(lib: BeatriceLib, precision: int) -> None:
    """Precision of beatrice.h contexts built from now on (0 f32, 1 bf16, 2 bf16x3, -1 default)."""
    bind(lib).BeatriceB200_SetDefaultPrecision(precision)


def last_error(lib: BeatriceLib):
    """(code, text) of the latched library failure; (0, "") when none."""
    dll = bind(lib)
    return dll.BeatriceB200_LastError(), (dll.BeatriceB200_LastErrorString() or b"").decode()


def clear_error(lib: BeatriceLib) -> None:
    bind(lib).BeatriceB200_ClearError()


class Engine:
    """n_streams voice streams on one GPU."""

    def __init__(self, lib: BeatriceLib, n_streams: int, device: int = 0, precision: int = 2):
        self.dll = bind(lib)
        self.n = n_streams
        self.h = self.dll.BeatriceB200_CreateEngine(device, n_streams, precision)
        if not self.h:
            raise RuntimeError(
                f"BeatriceB200_CreateEngine(device={device}, n_streams={n_streams}, precision={precision}) failed: "
                "no usable CUDA device or unsupported precision (there is no CPU fallback)")
        self._pinned = {}
        self._dev = {}

    # ---- loading ----
    def load(self, model_dir: str) -> int:
        return self.dll.BeatriceB200_LoadModel(self.h, model_dir.encode("utf-8"))

    def load_from_memory(self, images) -> int:
        """images: five ``bytes``/uint8 arrays in the order phone, pitch, wavegen, setter, speakers."""
        bufs = [np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else
                np.ascontiguousarray(b, np.uint8) for b in images]
        ptrs = (_vp * 5)(*[b.ctypes.data for b in bufs])
        sizes = (C.c_size_t * 5)(*[b.size for b in bufs])
        return self.dll.BeatriceB200_LoadModelFromMemory(self.h, ptrs, sizes)

    @property
    def family(self) -> int:
        return self.dll.BeatriceB200_ModelFamily(self.h)

    @property
    def n_speakers(self) -> int:
        return self.dll.BeatriceB200_NumSpeakers(self.h)

    # ---- per-stream parameters (ProcessorCore2 setters) ----
    def set(self, name: str, value, stream: int = -1) -> int:
        fn = getattr(self.dll, "BeatriceB200_Set" + name)
        return fn(self.h, stream, value)

    def set_morph_weights(self, weights, stream: int = -1) -> int:
        """ProcessorCore2::SetSpeakerMorphingWeights; select the morphing slot with set("TargetSpeaker", n_speakers)."""
        w = np.ascontiguousarray(weights, np.float32)
        return self.dll.BeatriceB200_SetSpeakerMorphingWeights(self.h, stream, w.ctypes.data_as(_f32p), int(w.size))

    def seed_morph_lottery(self, seed: int) -> int:
        return self.dll.BeatriceB200_SeedMorphLottery(self.h, seed)

    def morph_state(self, stream: int):
        """(additive average [256], registered key-value embedding [384,128], lottery pick) of a stream's morphing slot."""
        add = np.empty(256, np.float32)
        kv = np.empty((384, 128), np.float32)
        pick = C.c_int(-1)
        rc = self.dll.BeatriceB200_GetMorphState(self.h, stream, add.ctypes.data_as(_f32p), kv.ctypes.data_as(_f32p),
                                                 C.byref(pick))
        if rc != 0:
            raise RuntimeError(f"BeatriceB200_GetMorphState -> {rc}")
        return add, kv, pick.value

    def reset_stream(self, stream: int = -1) -> int:
        return self.dll.BeatriceB200_ResetStream(self.h, stream)

    # ---- processing with host buffers (H2D + D2H inside the call) ----
    def pinned(self, key: str, shape):
        """A pinned float32 host array owned by the engine wrapper."""
        n = int(np.prod(shape))
        if key not in self._pinned or self._pinned[key][1] < n:
            p = self.dll.BeatriceB200_AllocPinned(n * 4)
            self._pinned[key] = (p, n)
        p, _ = self._pinned[key]
        return np.ctypeslib.as_array(C.cast(p, _f32p), shape=(n,)).reshape(shape)

    def process_frames(self, x: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """x [n_streams,160] @16 kHz -> [n_streams,240] @24 kHz."""
        x = np.ascontiguousarray(x, np.float32)
        assert x.shape == (self.n, 160)
        if out is None:
            out = np.empty((self.n, 240), np.float32)
        rc = self.dll.BeatriceB200_ProcessFrames(self.h, x.ctypes.data, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"BeatriceB200_ProcessFrames -> {rc}")
        return out

    def process_48k(self, x: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """x [n_streams,480] @48 kHz -> [n_streams,480] @48 kHz (one 10 ms hop)."""
        x = np.ascontiguousarray(x, np.float32)
        assert x.shape == (self.n, 480)
        if out is None:
            out = np.empty((self.n, 480), np.float32)
        rc = self.dll.BeatriceB200_Process48k(self.h, x.ctypes.data, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"BeatriceB200_Process48k -> {rc}")
        return out

    # ---- device-resident processing ----
    def dev_alloc(self, key: str, n_floats: int):
        if key not in self._dev:
            self._dev[key] = self.dll.BeatriceB200_AllocDevice(self.h, n_floats * 4)
        return self._dev[key]

    def to_device(self, dev_ptr, host: np.ndarray):
        host = np.ascontiguousarray(host, np.float32)
        self.dll.BeatriceB200_CopyToDevice(self.h, dev_ptr, host.ctypes.data, host.nbytes)

    def to_host(self, dev_ptr, shape) -> np.ndarray:
        out = np.empty(shape, np.float32)
        self.dll.BeatriceB200_CopyToHost(self.h, out.ctypes.data, dev_ptr, out.nbytes)
        return out

    def process_frames_device(self, in_dev, out_dev) -> int:
        return self.dll.BeatriceB200_ProcessFramesDevice(self.h, in_dev, out_dev)

    def process_48k_device(self, in_dev, out_dev) -> int:
        return self.dll.BeatriceB200_Process48kDevice(self.h, in_dev, out_dev)

    def set_pipeline_depth(self, depth: int) -> int:
        """1: a call returns the hop it was given; 2: vocoder of the previous hop || encoders of this one."""
        return self.dll.BeatriceB200_SetPipelineDepth(self.h, depth)

    def set_pipeline_plan(self, plan: str) -> int:
        return self.dll.BeatriceB200_SetPipelinePlan(self.h, plan.encode("utf-8"))

    def set_skip_ops(self, names: str) -> int:
        """Measurement aid: drop the hop ops whose names contain one of the comma-separated substrings ("" = none)."""
        return self.dll.BeatriceB200_SetSkipOps(self.h, names.encode("utf-8"))

    def set_host_sample_rate(self, rate: float) -> int:
        """ProcessorCore2::SetSampleRate for :meth:`process_any_rate`."""
        return self.dll.BeatriceB200_SetHostSampleRate(self.h, float(rate))

    def process_any_rate(self, x: np.ndarray) -> np.ndarray:
        """x [n][m] at the host rate -> [n][m] (ProcessorCore2::Process per stream, any block size m)."""
        x = np.ascontiguousarray(x, np.float32)
        assert x.ndim == 2 and x.shape[0] == self.n
        out = np.empty_like(x)
        rc = self.dll.BeatriceB200_ProcessAnyRate(self.h, x.ctypes.data, out.ctypes.data, x.shape[1])
        if rc != 0:
            raise RuntimeError(f"BeatriceB200_ProcessAnyRate -> {rc}")
        return out

    def set_echo_model(self, on: bool) -> int:
        return self.dll.BeatriceB200_SetEchoModel(self.h, 1 if on else 0)

    def set_upsampler_form(self, form: int) -> int:
        """1 = in the fused MRF kernels' prologue, 0 = own launches, -1 = by pipeline depth (default)."""
        return self.dll.BeatriceB200_SetUpsamplerForm(self.h, form)

    def drain(self, model_rate: bool = False) -> np.ndarray:
        """Depth 2: the blocks of the hop still in flight ([n,480] @48 kHz, or [n,240] @24 kHz)."""
        out = np.empty((self.n, 240 if model_rate else 480), np.float32)
        rc = self.dll.BeatriceB200_DrainPipeline(self.h, out.ctypes.data if model_rate else None,
                                                 None if model_rate else out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"BeatriceB200_DrainPipeline -> {rc}")
        return out

    def synchronize(self):
        self.dll.BeatriceB200_Synchronize(self.h)

    @property
    def cuda_stream(self) -> int:
        return self.dll.BeatriceB200_Stream(self.h)

    # ---- introspection ----
    def last_intermediates(self):
        phone = np.empty((self.n, self.dll.BeatriceB200_PhoneChannels(self.h)), np.float32)
        q_raw = np.empty(self.n, np.int32)
        q_used = np.empty(self.n, np.int32)
        feat = np.empty((self.n, 4), np.float32)
        self.dll.BeatriceB200_GetLastIntermediates(
            self.h, phone.ctypes.data_as(_f32p), q_raw.ctypes.data_as(_i32p), q_used.ctypes.data_as(_i32p),
            feat.ctypes.data_as(_f32p))
        return phone, q_raw, q_used, feat

    def transform_pitch_bins(self, bins_raw: np.ndarray) -> np.ndarray:
        """The call-site pitch transform with every stream's current parameters (device kernel)."""
        q = np.ascontiguousarray(bins_raw, np.int32)
        assert q.shape == (self.n,)
        out = np.empty(self.n, np.int32)
        rc = self.dll.BeatriceB200_TransformPitchBins(self.h, q.ctypes.data_as(_i32p), out.ctypes.data_as(_i32p))
        if rc != 0:
            raise RuntimeError(f"BeatriceB200_TransformPitchBins -> {rc}")
        return out

    def adapter_only_48k(self, x48: np.ndarray, model24: np.ndarray):
        """One hop of the device-side 48 kHz adapter with the model replaced by ``model24``; returns (x16, out48)."""
        x48 = np.ascontiguousarray(x48, np.float32)
        model24 = np.ascontiguousarray(model24, np.float32)
        assert x48.shape == (self.n, 480) and model24.shape == (self.n, 240)
        x16 = np.empty((self.n, 160), np.float32)
        out = np.empty((self.n, 480), np.float32)
        rc = self.dll.BeatriceB200_AdapterOnly48k(self.h, x48.ctypes.data, model24.ctypes.data, x16.ctypes.data,
                                                  out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"BeatriceB200_AdapterOnly48k -> {rc}")
        return x16, out

    def resident_bytes(self) -> int:
        return int(self.dll.BeatriceB200_ResidentBytes(self.h))

    def kernel_launches(self) -> int:
        return int(self.dll.BeatriceB200_KernelLaunchCount(self.h))

    def profile_hop(self, in_dev, out_dev, capacity: int = 256):
        recs = (KernelRecord * capacity)()
        n = self.dll.BeatriceB200_ProfileHop(self.h, in_dev, out_dev, C.cast(recs, _vp), capacity)
        return [dict(name=recs[i].name.decode(), ms=float(recs[i].ms), flops=float(recs[i].flops),
                     bytes=float(recs[i].bytes)) for i in range(max(n, 0))]

    def close(self):
        if self.h:
            for p in self._dev.values():
                self.dll.BeatriceB200_FreeDevice(self.h, p)
            self.dll.BeatriceB200_DestroyEngine(self.h)
            for p, _ in self._pinned.values():
                self.dll.BeatriceB200_FreePinned(p)
            self.h = None
