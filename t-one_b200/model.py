"""Host side of the drop-in: ctypes binding of the C ABI and the reference-shaped model class.

``B200StreamingCTCModel`` mirrors ``tone.onnx_wrapper.StreamingCTCModel``
(reference: tone/onnx_wrapper.py:17-123): same class constants, same ``forward(audio_chunk,
state) -> (logprobs, state_next)`` contract, same validation and exception types, so it plugs
into the unchanged ``StreamingCTCPipeline`` constructor (reference: tone/pipeline.py:100-109,
143-147).  There is no CPU fallback: if the CUDA library cannot be loaded, construction fails.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Optional

import numpy as np

from .arch import DEFAULT_ARCH
from . import weights as _weights

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TONE_B200_LIB") or os.path.join(_HERE, "libtone_b200.so")

TONE_OK, TONE_EINVAL, TONE_ENOMEM, TONE_ECUDA, TONE_ESTATE, TONE_ERANGE = 0, -1, -2, -3, -4, -5

# every symbol include/tone_b200.h declares
SYMBOLS = (
    "tone_create", "tone_destroy", "tone_get_info", "tone_last_error", "tone_load_weight",
    "tone_finalize_weights", "tone_alloc_slots", "tone_release_slots", "tone_reset_slots", "tone_step",
    "tone_stage", "tone_step_staged", "tone_fetch", "tone_fetch_greedy", "tone_sync", "tone_step_device", "tone_host_buffers", "tone_export_state",
    "tone_import_state", "tone_step_debug", "tone_selftest_gemm", "tone_cluster_prof_read", "tone_step_features",
)


class ToneConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("chunk_samples", C.c_int32), ("max_slots", C.c_int32),
                ("max_batch", C.c_int32), ("gemm_impl", C.c_int32), ("use_graph", C.c_int32),
                ("cluster_max_batch", C.c_int32)]


class ToneInfo(C.Structure):
    _fields_ = [("chunk_samples", C.c_int32), ("frames_out", C.c_int32), ("n_classes", C.c_int32),
                ("state_size", C.c_int32), ("max_slots", C.c_int32), ("max_batch", C.c_int32),
                ("launches_per_step", C.c_int32), ("n_taps", C.c_int32),
                ("state_bytes_per_slot", C.c_int64), ("weight_bytes", C.c_int64)]


_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """dlopen the CUDA library (built in-tree by ``build.py``).  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: the CUDA library is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs nvcc); there is no CPU fallback for the acoustic-model step.")
    lib = C.CDLL(p)
    vp, i32p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)
    lib.tone_create.argtypes = [C.POINTER(ToneConfig), C.POINTER(vp)]
    lib.tone_destroy.argtypes = [vp]
    lib.tone_destroy.restype = None
    lib.tone_get_info.argtypes = [vp, C.POINTER(ToneInfo)]
    lib.tone_last_error.argtypes = []
    lib.tone_last_error.restype = C.c_char_p
    lib.tone_load_weight.argtypes = [vp, C.c_char_p, f32p, C.POINTER(C.c_int64), C.c_int32]
    lib.tone_finalize_weights.argtypes = [vp]
    lib.tone_alloc_slots.argtypes = [vp, C.c_int32, i32p]
    lib.tone_release_slots.argtypes = [vp, C.c_int32, i32p]
    lib.tone_reset_slots.argtypes = [vp, C.c_int32, i32p]
    lib.tone_step.argtypes = [vp, C.c_int32, i32p, i32p, f32p, i32p]
    lib.tone_stage.argtypes = [vp, C.c_int32, i32p, i32p]
    lib.tone_step_staged.argtypes = [vp, C.c_int32, vp]
    lib.tone_fetch.argtypes = [vp, C.c_int32, f32p, i32p]
    lib.tone_sync.argtypes = [vp]
    lib.tone_fetch_greedy.argtypes = [vp, C.c_int32, i32p, f32p]
    lib.tone_step_device.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp]
    lib.tone_host_buffers.argtypes = [vp, C.POINTER(i32p), C.POINTER(i32p), C.POINTER(f32p), C.POINTER(i32p)]
    lib.tone_export_state.argtypes = [vp, C.c_int32, C.POINTER(C.c_uint16)]
    lib.tone_import_state.argtypes = [vp, C.c_int32, C.POINTER(C.c_uint16)]
    lib.tone_step_debug.argtypes = [vp, C.c_int32, i32p, i32p, f32p, i32p, f32p]
    lib.tone_selftest_gemm.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, f32p, f32p, f32p, C.c_int32]
    for s in SYMBOLS:
        if s not in ("tone_destroy", "tone_last_error"):
            getattr(lib, s).restype = C.c_int
    if path is None:
        _lib = lib
    return lib


class ToneError(RuntimeError):
    pass


def _raise(lib, code: int):
    msg = (lib.tone_last_error() or b"").decode("utf-8", "replace")
    if code in (TONE_EINVAL, TONE_ERANGE):
        raise ValueError(msg)
    if code == TONE_ENOMEM:
        raise MemoryError(msg)
    raise ToneError(f"tone_b200 error {code}: {msg}")


def _i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Engine:
    """One engine per GPU: weights, resident stream slots, step."""

    def __init__(self, weights=None, chunk_samples: int = 2400, max_slots: int = 64, max_batch: Optional[int] = None,
                 device: int = 0, gemm_impl: int = 0, use_graph: bool = True, cluster_max_batch: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        cfg = ToneConfig(device, chunk_samples, max_slots, max_batch or max_slots, gemm_impl, int(use_graph),
                         int(cluster_max_batch))
        rc = self._lib.tone_create(C.byref(cfg), C.byref(self._h))
        if rc:
            self._h = C.c_void_p()
            _raise(self._lib, rc)
        self._finalizer = weakref.finalize(self, self._lib.tone_destroy, self._h)
        self.info = self._get_info()
        self.T = self.info.frames_out
        self.chunk_samples = chunk_samples
        self._wrap_host_buffers()
        if weights is not None:
            self.load_weights(weights)

    # -- plumbing
    def _ck(self, rc):
        if rc:
            _raise(self._lib, rc)

    def _get_info(self) -> ToneInfo:
        info = ToneInfo()
        self._ck(self._lib.tone_get_info(self._h, C.byref(info)))
        return info

    def _wrap_host_buffers(self):
        i32p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_float)
        s, p, l, t = i32p(), i32p(), f32p(), i32p()
        self._ck(self._lib.tone_host_buffers(self._h, C.byref(s), C.byref(p), C.byref(l), C.byref(t)))
        mb = self.info.max_batch
        self.h_slots = np.ctypeslib.as_array(s, shape=(mb,))
        self.h_pcm = np.ctypeslib.as_array(p, shape=(mb, self.chunk_samples))
        self.h_logprobs = np.ctypeslib.as_array(l, shape=(mb, 13, 35)).reshape(-1)
        self.h_tokens = np.ctypeslib.as_array(t, shape=(mb, 13)).reshape(-1)

    def close(self):
        self._finalizer()

    # -- weights
    def load_weights(self, weights) -> None:
        """weights: mapping reference-state_dict-name -> array (numpy or torch), fp32 shapes as in the reference."""
        w = _weights.from_state_dict(weights)
        for name, arr in w.items():
            a = np.ascontiguousarray(arr, dtype=np.float32)
            shp = (C.c_int64 * a.ndim)(*a.shape)
            self._ck(self._lib.tone_load_weight(self._h, name.encode(), _f32p(a), shp, a.ndim))
        self._ck(self._lib.tone_finalize_weights(self._h))
        self.info = self._get_info()

    # -- slots
    def alloc_slots(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.int32)
        self._ck(self._lib.tone_alloc_slots(self._h, n, _i32p(out)))
        return out

    def release_slots(self, slots) -> None:
        s = np.ascontiguousarray(slots, dtype=np.int32)
        self._ck(self._lib.tone_release_slots(self._h, len(s), _i32p(s)))

    def reset_slots(self, slots) -> None:
        s = np.ascontiguousarray(slots, dtype=np.int32)
        self._ck(self._lib.tone_reset_slots(self._h, len(s), _i32p(s)))

    # -- step
    def step(self, slots, pcm, want_tokens: bool = True):
        """pcm int32 (B, chunk) -> (logprobs fp32 (B,T,35), tokens int32 (B,T)); state advances in the slots."""
        s = np.ascontiguousarray(slots, dtype=np.int32)
        x = np.ascontiguousarray(pcm, dtype=np.int32)
        B = len(s)
        if x.shape != (B, self.chunk_samples):
            raise ValueError(f"pcm must be ({B}, {self.chunk_samples}), got {x.shape}")
        lp = np.empty((B, self.T, 35), dtype=np.float32)
        tk = np.empty((B, self.T), dtype=np.int32)
        self._ck(self._lib.tone_step(self._h, B, _i32p(s), _i32p(x), _f32p(lp), _i32p(tk) if want_tokens else None))
        return lp, tk

    def step_features(self, slots, feats, want_tokens: bool = True):
        """Feature-input mode (reference ``skip_preprocessor=True``): feats (B, 64, F) log-mel, rounded to fp16."""
        s = np.ascontiguousarray(slots, dtype=np.int32)
        f = np.ascontiguousarray(feats, dtype=np.float16)
        B = len(s)
        if f.shape != (B, 64, self.chunk_samples // 80):
            raise ValueError(f"feats must have shape {(B, 64, self.chunk_samples // 80)}, got {f.shape}")
        lp = np.empty((B, self.T, 35), dtype=np.float32)
        tk = np.empty((B, self.T), dtype=np.int32)
        self._ck(self._lib.tone_step_features(self._h, B, _i32p(s), f.ctypes.data_as(C.c_void_p), _f32p(lp), _i32p(tk)))
        return (lp, tk) if want_tokens else lp

    def step_debug(self, slots, pcm):
        s = np.ascontiguousarray(slots, dtype=np.int32)
        x = np.ascontiguousarray(pcm, dtype=np.int32)
        B = len(s)
        lp = np.empty((B, self.T, 35), dtype=np.float32)
        tk = np.empty((B, self.T), dtype=np.int32)
        taps = np.zeros((self.info.n_taps, B * self.T, 384), dtype=np.float32)
        self._ck(self._lib.tone_step_debug(self._h, B, _i32p(s), _i32p(x), _f32p(lp), _i32p(tk), _f32p(taps)))
        return lp, tk, taps

    # zero-copy staged API (benchmarks): write self.h_slots / self.h_pcm, then stage(), step_staged(), fetch()
    def stage(self, B: int) -> None:
        self._ck(self._lib.tone_stage(self._h, B, _i32p(self.h_slots), _i32p(self.h_pcm)))

    def step_staged(self, B: int, cuda_stream: int = 0) -> None:
        self._ck(self._lib.tone_step_staged(self._h, B, C.c_void_p(cuda_stream) if cuda_stream else None))

    def step_device(self, B: int, d_slots: int, d_pcm: int, d_logprobs: int = 0, d_tokens: int = 0,
                    cuda_stream: int = 0) -> None:
        """Asynchronous step on raw device pointers (e.g. torch tensors' data_ptr())."""
        vp = C.c_void_p
        self._ck(self._lib.tone_step_device(self._h, B, vp(d_slots or None), vp(d_pcm or None),
                                            vp(d_logprobs or None), vp(d_tokens or None), vp(cuda_stream or None)))

    def fetch(self, B: int, tokens: bool = True):
        self._ck(self._lib.tone_fetch(self._h, B, _f32p(self.h_logprobs), _i32p(self.h_tokens) if tokens else None))
        lp = self.h_logprobs[: B * self.T * 35].reshape(B, self.T, 35)
        tk = self.h_tokens[: B * self.T].reshape(B, self.T)
        return lp, tk

    def step_greedy(self, slots, pcm):
        """Step and fetch only what greedy decoding needs: tokens int32 (B,T) and the (space, blank) log-probs
        fp32 (B,T,2) that the phrase splitter thresholds on."""
        s = np.ascontiguousarray(slots, dtype=np.int32)
        x = np.ascontiguousarray(pcm, dtype=np.int32)
        B = len(s)
        if x.shape != (B, self.chunk_samples):
            raise ValueError(f"pcm must be ({B}, {self.chunk_samples}), got {x.shape}")
        self._ck(self._lib.tone_stage(self._h, B, _i32p(s), _i32p(x)))
        self._ck(self._lib.tone_step_staged(self._h, B, None))
        tk = np.empty((B, self.T), dtype=np.int32)
        sil = np.empty((B, self.T, 2), dtype=np.float32)
        self._ck(self._lib.tone_fetch_greedy(self._h, B, _i32p(tk), _f32p(sil)))
        return tk, sil

    def step_pinned(self, B: int):
        """H2D of the pinned inputs + step + D2H into the pinned outputs, synchronous (the e2e path)."""
        self._ck(self._lib.tone_step(self._h, B, _i32p(self.h_slots), _i32p(self.h_pcm), _f32p(self.h_logprobs),
                                     _i32p(self.h_tokens)))
        return (self.h_logprobs[: B * self.T * 35].reshape(B, self.T, 35), self.h_tokens[: B * self.T].reshape(B, self.T))

    def sync(self) -> None:
        self._ck(self._lib.tone_sync(self._h))

    # -- state wire format
    def export_state(self, slot: int) -> np.ndarray:
        out = np.empty(self.info.state_size, dtype=np.float16)
        self._ck(self._lib.tone_export_state(self._h, int(slot), out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out

    def import_state(self, slot: int, state: np.ndarray) -> None:
        a = np.ascontiguousarray(state, dtype=np.float16)
        if a.shape != (self.info.state_size,):
            raise ValueError(f"state must be ({self.info.state_size},), got {a.shape}")
        self._ck(self._lib.tone_import_state(self._h, int(slot), a.ctypes.data_as(C.POINTER(C.c_uint16))))

    def selftest_gemm(self, A: np.ndarray, W: np.ndarray, block_n: int = 64) -> np.ndarray:
        A = np.ascontiguousarray(A, dtype=np.float32)
        W = np.ascontiguousarray(W, dtype=np.float32)
        M, K = A.shape
        N = W.shape[0]
        out = np.empty((M, N), dtype=np.float32)
        self._ck(self._lib.tone_selftest_gemm(self._h, M, N, K, _f32p(A), _f32p(W), _f32p(out), block_n))
        return out


class StreamSlots:
    """Opaque model state for the device-resident mode: the slot ids of B streams.

    The pipeline never inspects the model state (reference: tone/pipeline.py:143-147,172; the Triton client's
    state is a bare counter, dev/triton/client_wer.py:138-207), so a handle is a valid state."""

    def __init__(self, engine: Engine, slots: np.ndarray):
        self.engine, self.slots = engine, slots
        self._fin = weakref.finalize(self, StreamSlots._release, weakref.ref(engine), slots.copy())

    @staticmethod
    def _release(engine_ref, slots):
        eng = engine_ref()
        if eng is not None and eng._finalizer.alive:
            try:
                eng.release_slots(slots)
            except Exception:
                pass

    def release(self):
        self._fin()

    def __len__(self):
        return len(self.slots)


class B200StreamingCTCModel:
    """Acoustic model with the reference's interface, running on one B200.

    state_mode="numpy": ``forward`` takes/returns the reference's flat fp16 state ``(B, 219729)`` - each call
    imports the state into scratch slots, steps, and exports it again (parity / migration path; state crosses PCIe).
    state_mode="device": state is a :class:`StreamSlots` handle; per-stream state stays in HBM (throughput path).
    """

    SAMPLE_RATE = 8000            # reference: tone/onnx_wrapper.py:30-34
    MEAN_TIME_BIAS = 0.33
    AUDIO_CHUNK_SAMPLES = 2400
    FRAME_SIZE = 0.03
    STATE_SIZE = DEFAULT_ARCH.state_size

    def __init__(self, weights, *, state_mode: str = "numpy", max_streams: int = 64, device: int = 0,
                 chunk_samples: Optional[int] = None, gemm_impl: int = 0, use_graph: bool = True):
        if state_mode not in ("numpy", "device"):
            raise ValueError("state_mode must be 'numpy' or 'device'")
        self.state_mode = state_mode
        self.AUDIO_CHUNK_SAMPLES = int(chunk_samples or type(self).AUDIO_CHUNK_SAMPLES)
        self.engine = Engine(weights, chunk_samples=self.AUDIO_CHUNK_SAMPLES, max_slots=max_streams,
                             max_batch=max_streams, device=device, gemm_impl=gemm_impl, use_graph=use_graph)
        self._scratch = None

    # -- factories mirroring tone/onnx_wrapper.py:38-78
    @classmethod
    def from_seed(cls, seed: int = 0, **kw) -> "B200StreamingCTCModel":
        """Seeded synthetic weights of the configs/streaming_acoustic architecture (no checkpoint is available offline)."""
        return cls(_weights.init_weights(seed), **kw)

    @classmethod
    def from_local(cls, model_path, **kw) -> "B200StreamingCTCModel":
        """Load a ToneForCTC / Tone state_dict saved as .npz, .pt/.bin (torch) or .safetensors."""
        p = str(model_path)
        if p.endswith(".npz"):
            sd = dict(np.load(p))
        elif p.endswith(".safetensors"):
            from safetensors.numpy import load_file
            sd = load_file(p)
        else:
            import torch
            sd = torch.load(p, map_location="cpu")
            sd = sd.get("state_dict", sd)
        return cls(sd, **kw)

    @classmethod
    def from_hugging_face(cls, **kw) -> "B200StreamingCTCModel":
        """Use the HF checkpoint t-tech/T-one only if it is already in the local cache (no network here)."""
        from huggingface_hub import hf_hub_download
        path = hf_hub_download("t-tech/T-one", "model.safetensors", local_files_only=True)
        return cls.from_local(path, **kw)

    # -- the step (reference: tone/onnx_wrapper.py:84-123)
    def forward(self, audio_chunk, state=None):
        if not isinstance(audio_chunk, np.ndarray):
            raise TypeError(f"Incorrect 'audio_chunk' type: expected np.ndarray, but got {type(audio_chunk)}")
        if audio_chunk.ndim != 3 or audio_chunk.shape[1:] != (self.AUDIO_CHUNK_SAMPLES, 1):
            raise ValueError(
                f"Shape of 'audio_chunk' must be (B, {self.AUDIO_CHUNK_SAMPLES}, 1), but got {audio_chunk.shape}")
        if audio_chunk.dtype != np.int32:
            raise ValueError(f"Incorrect dtype of 'audio_chunk': expected np.int32, but got {audio_chunk.dtype}")
        if audio_chunk.min() < -32768 or audio_chunk.max() > 32767:
            raise ValueError("Samples in 'audio_chunk' must be in range [-32768; 32767], "
                             f"but it is in range [{audio_chunk.min()}; {audio_chunk.max()}]")
        B = audio_chunk.shape[0]
        pcm = audio_chunk[:, :, 0]
        eng = self.engine
        if self.state_mode == "device":
            if state is None:
                state = StreamSlots(eng, eng.alloc_slots(B))
            if not isinstance(state, StreamSlots):
                raise TypeError(f"Incorrect 'state' type: expected StreamSlots or None, but got {type(state)}")
            if len(state) != B:
                raise ValueError(f"'state' holds {len(state)} streams, but the batch has {B}")
            logprobs, _ = eng.step(state.slots, pcm, want_tokens=False)
            return [logprobs, state]
        # numpy mode
        if state is None:
            state = np.zeros((B, self.STATE_SIZE), dtype=np.float16)
        if not isinstance(state, np.ndarray):
            raise TypeError(f"Incorrect 'state' type: expected np.ndarray or None, but got {type(state)}")
        if state.shape != (B, self.STATE_SIZE):
            raise ValueError(f"Shape of 'state' must be ({B}, {self.STATE_SIZE}), but got {state.shape}")
        if state.dtype != np.float16:
            raise ValueError(f"Incorrect dtype of 'state': expected np.float16, but got {state.dtype}")
        if self._scratch is None or len(self._scratch) < B:
            if self._scratch is not None:
                eng.release_slots(self._scratch)
            self._scratch = eng.alloc_slots(B)
        slots = self._scratch[:B]
        for b in range(B):
            eng.import_state(int(slots[b]), state[b])
        logprobs, _ = eng.step(slots, pcm, want_tokens=False)
        state_next = np.stack([eng.export_state(int(slots[b])) for b in range(B)], 0)
        return [logprobs, state_next]
