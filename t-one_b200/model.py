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
PCM_I32, PCM_I16 = 0, 1
OUT_LOGPROBS, OUT_TOKENS, OUT_SIL, OUT_PHRASES = 1, 2, 4, 8
FLAG_NO_PDL, FLAG_NO_FUSED_VATT = 1, 2

# every symbol include/tone_b200.h declares
SYMBOLS = (
    "tone_create", "tone_destroy", "tone_get_info", "tone_last_error", "tone_load_weight",
    "tone_finalize_weights", "tone_alloc_slots", "tone_release_slots", "tone_reset_slots", "tone_step",
    "tone_submit", "tone_wait", "tone_next_staging", "tone_ticket_phrases", "tone_step_features",
    "tone_stage", "tone_step_staged", "tone_fetch", "tone_sync", "tone_fetch_greedy", "tone_step_device",
    "tone_export_states", "tone_import_states", "tone_export_states_triton", "tone_import_states_triton",
    "tone_step_debug", "tone_selftest_gemm", "tone_selftest_phrases",
    "tone_server_create", "tone_server_destroy", "tone_server_push", "tone_server_poll", "tone_server_get_stats",
)


class ToneConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("chunk_samples", C.c_int32), ("max_slots", C.c_int32),
                ("max_batch", C.c_int32), ("gemm_impl", C.c_int32), ("use_graph", C.c_int32),
                ("lanes", C.c_int32), ("lane_min_batch", C.c_int32), ("persist_min_tiles", C.c_int32),
                ("persist_mode", C.c_int32), ("split_k", C.c_int32), ("flags", C.c_int32),
                ("fused_ff", C.c_int32), ("fused_ff_min_rows", C.c_int32), ("att_block_min_rows", C.c_int32),
                ("lazy_norm_min_rows", C.c_int32), ("dw_pipe_min_batch", C.c_int32), ("att_pipe_min_batch", C.c_int32),
                ("rowgemm_min_rows", C.c_int32), ("persist_ctas", C.c_int32)]


class ToneInfo(C.Structure):
    _fields_ = [("chunk_samples", C.c_int32), ("frames_out", C.c_int32), ("n_classes", C.c_int32),
                ("state_size", C.c_int32), ("max_slots", C.c_int32), ("max_batch", C.c_int32),
                ("launches_per_step", C.c_int32), ("n_taps", C.c_int32),
                ("state_bytes_per_slot", C.c_int64), ("weight_bytes", C.c_int64),
                ("pipeline_depth", C.c_int32), ("max_phrases_per_step", C.c_int32)]


class TonePhrase(C.Structure):
    _fields_ = [("batch_index", C.c_int32), ("start_frame", C.c_int32), ("end_frame", C.c_int32),
                ("text_offset", C.c_int32), ("text_len", C.c_int32)]


class ToneServerConfig(C.Structure):
    _fields_ = [("max_batch", C.c_int32), ("max_queue_delay_us", C.c_int32), ("idle_timeout_ms", C.c_int32),
                ("queue_depth", C.c_int32), ("outputs", C.c_int32), ("prewarm", C.c_int32)]


class ToneServerStats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("chunks", C.c_int64), ("phrases", C.c_int64), ("streams_opened", C.c_int64),
                ("streams_closed", C.c_int64), ("streams_reclaimed", C.c_int64), ("rejected", C.c_int64),
                ("open_streams", C.c_int32), ("queued_chunks", C.c_int32), ("mean_batch", C.c_double),
                ("latency_ms_p50", C.c_double), ("latency_ms_p99", C.c_double), ("latency_ms_max", C.c_double),
                ("queue_ms_p50", C.c_double), ("queue_ms_p99", C.c_double)]


STREAM_PHRASE_DTYPE = np.dtype([("stream_id", "<u8"), ("start_frame", "<i4"), ("end_frame", "<i4"),
                                ("text_offset", "<i4"), ("text_len", "<i4")])

PHRASE_DTYPE = np.dtype([("batch_index", "<i4"), ("start_frame", "<i4"), ("end_frame", "<i4"),
                         ("text_offset", "<i4"), ("text_len", "<i4")])

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
    u8p, u16p, i16p, i64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_int16), C.POINTER(C.c_int64)
    lib.tone_create.argtypes = [C.POINTER(ToneConfig), C.POINTER(vp)]
    lib.tone_destroy.argtypes = [vp]
    lib.tone_destroy.restype = None
    lib.tone_get_info.argtypes = [vp, C.POINTER(ToneInfo)]
    lib.tone_last_error.argtypes = []
    lib.tone_last_error.restype = C.c_char_p
    lib.tone_load_weight.argtypes = [vp, C.c_char_p, f32p, i64p, C.c_int32]
    lib.tone_finalize_weights.argtypes = [vp]
    lib.tone_alloc_slots.argtypes = [vp, C.c_int32, i32p]
    lib.tone_release_slots.argtypes = [vp, C.c_int32, i32p]
    lib.tone_reset_slots.argtypes = [vp, C.c_int32, i32p]
    lib.tone_step.argtypes = [vp, C.c_int32, i32p, i32p, f32p, i32p]
    lib.tone_submit.argtypes = [vp, C.c_int32, i32p, vp, C.c_int32, u8p, C.c_int32, i32p]
    lib.tone_wait.argtypes = [vp, C.c_int32, f32p, i32p, f32p]
    lib.tone_next_staging.argtypes = [vp, C.POINTER(i32p), C.POINTER(i16p), C.POINTER(u8p)]
    lib.tone_ticket_phrases.argtypes = [vp, C.c_int32, C.POINTER(C.POINTER(TonePhrase)), i32p, C.POINTER(u8p), i32p]
    lib.tone_step_features.argtypes = [vp, C.c_int32, i32p, vp, f32p, i32p]
    lib.tone_stage.argtypes = [vp, C.c_int32, i32p, i32p]
    lib.tone_step_staged.argtypes = [vp, C.c_int32, vp]
    lib.tone_fetch.argtypes = [vp, C.c_int32, f32p, i32p]
    lib.tone_sync.argtypes = [vp]
    lib.tone_fetch_greedy.argtypes = [vp, C.c_int32, i32p, f32p]
    lib.tone_step_device.argtypes = [vp, C.c_int32, i32p, vp, C.c_int32, vp, vp, vp]
    lib.tone_export_states.argtypes = [vp, C.c_int32, i32p, u16p]
    lib.tone_import_states.argtypes = [vp, C.c_int32, i32p, u16p]
    lib.tone_export_states_triton.argtypes = [vp, C.c_int32, i32p, u16p, u16p, i64p]
    lib.tone_import_states_triton.argtypes = [vp, C.c_int32, i32p, u16p, u16p, i64p]
    lib.tone_step_debug.argtypes = [vp, C.c_int32, i32p, i32p, f32p, i32p, f32p]
    lib.tone_selftest_gemm.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, f32p, f32p, f32p, C.c_int32]
    lib.tone_selftest_phrases.argtypes = [vp, C.c_int32, i32p, C.c_int32, i32p, f32p, u8p]
    u64p = C.POINTER(C.c_uint64)
    lib.tone_server_create.argtypes = [vp, C.POINTER(ToneServerConfig), C.POINTER(vp)]
    lib.tone_server_destroy.argtypes = [vp]
    lib.tone_server_destroy.restype = None
    lib.tone_server_push.argtypes = [vp, C.c_int32, u64p, i16p, u8p]
    lib.tone_server_poll.argtypes = [vp, C.c_int32, i32p, u64p, i32p, f32p, f32p, vp, C.c_int32, i32p, u8p, C.c_int32, i32p]
    lib.tone_server_get_stats.argtypes = [vp, C.POINTER(ToneServerStats)]
    for s in SYMBOLS:
        if s not in ("tone_destroy", "tone_last_error", "tone_server_destroy"):
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


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _u16p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint16))


class Ticket:
    """One pipelined step in flight (``Engine.submit``): what ``Engine.wait`` needs to shape its results."""
    __slots__ = ("id", "B", "outputs")

    def __init__(self, id_: int, B: int, outputs: int):
        self.id, self.B, self.outputs = id_, B, outputs


class Engine:
    """One engine per GPU: weights, resident stream slots, step."""

    def __init__(self, weights=None, chunk_samples: int = 2400, max_slots: int = 64, max_batch: Optional[int] = None,
                 device: int = 0, gemm_impl: int = 0, use_graph: bool = True, *, lanes: int = 0, lane_min_batch: int = 0,
                 persist_min_tiles: int = 0, persist_mode: int = 0, split_k: int = 0, flags: int = 0,
                 fused_ff: int = 0, fused_ff_min_rows: int = 0, att_block_min_rows: int = 0, lazy_norm_min_rows: int = 0,
                 dw_pipe_min_batch: int = 0, att_pipe_min_batch: int = 0, rowgemm_min_rows: int = 0,
                 persist_ctas: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        cfg = ToneConfig(device, chunk_samples, max_slots, max_batch or max_slots, gemm_impl, int(use_graph),
                         lanes, lane_min_batch, persist_min_tiles, persist_mode, split_k, flags, fused_ff, fused_ff_min_rows,
                         att_block_min_rows, lazy_norm_min_rows, dw_pipe_min_batch, att_pipe_min_batch, rowgemm_min_rows,
                         persist_ctas)
        rc = self._lib.tone_create(C.byref(cfg), C.byref(self._h))
        if rc:
            self._h = C.c_void_p()
            _raise(self._lib, rc)
        self._finalizer = weakref.finalize(self, self._lib.tone_destroy, self._h)
        self.info = self._get_info()
        self.T = self.info.frames_out
        self.chunk_samples = chunk_samples
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

    # pipelined form: up to info.pipeline_depth tickets in flight; copies of neighbouring steps overlap the kernels
    def next_staging(self, B: int):
        """Pinned staging of the set the next ``submit`` uses: (slots int32 (B,), pcm int16 (B, chunk), is_last uint8 (B,)).
        Fill them in place and call ``submit_staged`` - no intermediate host copy."""
        sp, pp, lp = C.POINTER(C.c_int32)(), C.POINTER(C.c_int16)(), C.POINTER(C.c_uint8)()
        self._ck(self._lib.tone_next_staging(self._h, C.byref(sp), C.byref(pp), C.byref(lp)))
        return (np.ctypeslib.as_array(sp, shape=(B,)), np.ctypeslib.as_array(pp, shape=(B, self.chunk_samples)),
                np.ctypeslib.as_array(lp, shape=(B,)))

    def submit(self, slots, pcm, outputs: int = OUT_LOGPROBS | OUT_TOKENS, is_last=None) -> Ticket:
        """Enqueue one step (asynchronous).  pcm int16 or int32 (B, chunk); is_last uint8/bool (B,) for OUT_PHRASES."""
        s = np.ascontiguousarray(slots, dtype=np.int32)
        x = np.ascontiguousarray(pcm)
        B = len(s)
        if x.dtype not in (np.int16, np.int32) or x.shape != (B, self.chunk_samples):
            raise ValueError(f"pcm must be int16/int32 ({B}, {self.chunk_samples}), got {x.dtype} {x.shape}")
        last = None if is_last is None else np.ascontiguousarray(is_last, dtype=np.uint8)
        if last is not None and last.shape != (B,):
            raise ValueError(f"is_last must have shape ({B},)")
        t = C.c_int32(-1)
        self._ck(self._lib.tone_submit(self._h, B, _i32p(s), x.ctypes.data_as(C.c_void_p),
                                       PCM_I16 if x.dtype == np.int16 else PCM_I32,
                                       _u8p(last) if last is not None else None, outputs, C.byref(t)))
        return Ticket(t.value, B, outputs)

    def wait(self, ticket: Ticket) -> dict:
        """Block until the ticket's outputs are on the host -> dict with the requested of 'logprobs' (B,T,35),
        'tokens' (B,T), 'sil' (B,T,2), 'phrases' (list of (batch_index, start_frame, end_frame, label_ids))."""
        B, o = ticket.B, ticket.outputs
        lp = np.empty((B, self.T, 35), dtype=np.float32) if o & OUT_LOGPROBS else None
        tk = np.empty((B, self.T), dtype=np.int32) if o & OUT_TOKENS else None
        sl = np.empty((B, self.T, 2), dtype=np.float32) if o & OUT_SIL else None
        self._ck(self._lib.tone_wait(self._h, ticket.id, _f32p(lp) if lp is not None else None,
                                     _i32p(tk) if tk is not None else None, _f32p(sl) if sl is not None else None))
        out = {}
        if lp is not None:
            out["logprobs"] = lp
        if tk is not None:
            out["tokens"] = tk
        if sl is not None:
            out["sil"] = sl
        if o & OUT_PHRASES:
            out["phrases"] = self.ticket_phrases(ticket.id)
        return out

    def ticket_phrases(self, ticket_id: int):
        """Finished phrases of a waited ticket: list of (batch_index, start_frame, end_frame, label ids uint8 array)."""
        ph, n = C.POINTER(TonePhrase)(), C.c_int32()
        pool, npool = C.POINTER(C.c_uint8)(), C.c_int32()
        self._ck(self._lib.tone_ticket_phrases(self._h, ticket_id, C.byref(ph), C.byref(n), C.byref(pool), C.byref(npool)))
        if n.value == 0:
            return []
        rec = np.frombuffer((C.c_char * (n.value * C.sizeof(TonePhrase))).from_address(C.addressof(ph.contents)),
                            dtype=PHRASE_DTYPE).copy()
        text = np.ctypeslib.as_array(pool, shape=(max(npool.value, 1),)).copy()
        return [(int(r["batch_index"]), int(r["start_frame"]), int(r["end_frame"]),
                 text[r["text_offset"]: r["text_offset"] + r["text_len"]]) for r in rec]

    def step_phrases(self, slots, pcm, is_last=None):
        """Synchronous step that returns only the phrases finished by this chunk (device-side splitter + greedy decode)."""
        return self.wait(self.submit(slots, pcm, OUT_PHRASES, is_last))["phrases"]

    def selftest_phrases(self, slots, tokens, sil, is_last=None):
        """Debug: the device-side splitter alone on per-frame (tokens (B,n), sil (B,n,2)), n <= 13 frames per call."""
        s = np.ascontiguousarray(slots, dtype=np.int32)
        tk = np.ascontiguousarray(tokens, dtype=np.int32)
        sl = np.ascontiguousarray(sil, dtype=np.float32)
        B, n = tk.shape
        last = None if is_last is None else np.ascontiguousarray(is_last, dtype=np.uint8)
        self._ck(self._lib.tone_selftest_phrases(self._h, B, _i32p(s), n, _i32p(tk), _f32p(sl),
                                                 _u8p(last) if last is not None else None))
        return self.ticket_phrases(-1)

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

    # staged API: stage once, step any number of times on the staged chunk, fetch
    def stage(self, slots, pcm) -> None:
        s = np.ascontiguousarray(slots, dtype=np.int32)
        x = np.ascontiguousarray(pcm, dtype=np.int32)
        self._ck(self._lib.tone_stage(self._h, len(s), _i32p(s), _i32p(x)))

    def step_staged(self, B: int, cuda_stream: int = 0) -> None:
        self._ck(self._lib.tone_step_staged(self._h, B, C.c_void_p(cuda_stream) if cuda_stream else None))

    def step_device(self, slots, d_pcm: int, pcm_format: int = PCM_I16, d_logprobs: int = 0, d_tokens: int = 0,
                    cuda_stream: int = 0) -> None:
        """Asynchronous step on raw device pointers (e.g. torch tensors' data_ptr()); slot ids are host ints."""
        s = np.ascontiguousarray(slots, dtype=np.int32)
        vp = C.c_void_p
        self._ck(self._lib.tone_step_device(self._h, len(s), _i32p(s), vp(d_pcm or None), pcm_format,
                                            vp(d_logprobs or None), vp(d_tokens or None), vp(cuda_stream or None)))

    def fetch(self, B: int, tokens: bool = True):
        lp = np.empty((B, self.T, 35), dtype=np.float32)
        tk = np.empty((B, self.T), dtype=np.int32)
        self._ck(self._lib.tone_fetch(self._h, B, _f32p(lp), _i32p(tk) if tokens else None))
        return lp, tk

    def step_greedy(self, slots, pcm):
        """Step and fetch only what greedy decoding needs: tokens int32 (B,T) and the (space, blank) log-probs
        fp32 (B,T,2) that the phrase splitter thresholds on."""
        r = self.wait(self.submit(slots, pcm, OUT_TOKENS | OUT_SIL))
        return r["tokens"], r["sil"]

    def sync(self) -> None:
        self._ck(self._lib.tone_sync(self._h))

    # -- state wire formats
    def export_states(self, slots) -> np.ndarray:
        """-> (n, 219729) float16, the reference's flat state of each slot."""
        s = np.ascontiguousarray(slots, dtype=np.int32)
        out = np.empty((len(s), self.info.state_size), dtype=np.float16)
        self._ck(self._lib.tone_export_states(self._h, len(s), _i32p(s), _u16p(out)))
        return out

    def import_states(self, slots, states) -> None:
        s = np.ascontiguousarray(slots, dtype=np.int32)
        a = np.ascontiguousarray(states, dtype=np.float16)
        if a.shape != (len(s), self.info.state_size):
            raise ValueError(f"states must be ({len(s)}, {self.info.state_size}), got {a.shape}")
        self._ck(self._lib.tone_import_states(self._h, len(s), _i32p(s), _u16p(a)))

    def export_state(self, slot: int) -> np.ndarray:
        return self.export_states([int(slot)])[0]

    def import_state(self, slot: int, state: np.ndarray) -> None:
        a = np.ascontiguousarray(state, dtype=np.float16)
        if a.shape != (self.info.state_size,):
            raise ValueError(f"state must be ({self.info.state_size},), got {a.shape}")
        self.import_states([int(slot)], a[None])

    def export_states_triton(self, slots):
        """-> (cache_last_time (n,18,384,30) f16, cache_last_channel (n,32,8,50) f16, cache_last_chan_len (n,) i64);
        reference: tone/scripts/export.py:335-376."""
        s = np.ascontiguousarray(slots, dtype=np.int32)
        n = len(s)
        tm = np.empty((n, 18, 384, 30), dtype=np.float16)
        ch = np.empty((n, 32, 8, 50), dtype=np.float16)
        ln = np.empty((n,), dtype=np.int64)
        self._ck(self._lib.tone_export_states_triton(self._h, n, _i32p(s), _u16p(tm), _u16p(ch),
                                                     ln.ctypes.data_as(C.POINTER(C.c_int64))))
        return tm, ch, ln

    def import_states_triton(self, slots, cache_last_time, cache_last_channel, cache_last_chan_len) -> None:
        s = np.ascontiguousarray(slots, dtype=np.int32)
        n = len(s)
        tm = np.ascontiguousarray(cache_last_time, dtype=np.float16)
        ch = np.ascontiguousarray(cache_last_channel, dtype=np.float16)
        ln = np.ascontiguousarray(cache_last_chan_len, dtype=np.int64)
        if tm.shape != (n, 18, 384, 30) or ch.shape != (n, 32, 8, 50) or ln.shape != (n,):
            raise ValueError("expected cache_last_time (n,18,384,30), cache_last_channel (n,32,8,50), cache_last_chan_len (n,)")
        self._ck(self._lib.tone_import_states_triton(self._h, n, _i32p(s), _u16p(tm), _u16p(ch),
                                                     ln.ctypes.data_as(C.POINTER(C.c_int64))))

    def selftest_gemm(self, A: np.ndarray, W: np.ndarray, block_n: int = 64) -> np.ndarray:
        A = np.ascontiguousarray(A, dtype=np.float32)
        W = np.ascontiguousarray(W, dtype=np.float32)
        M, K = A.shape
        N = W.shape[0]
        out = np.empty((M, N), dtype=np.float32)
        self._ck(self._lib.tone_selftest_gemm(self._h, M, N, K, _f32p(A), _f32p(W), _f32p(out), block_n))
        return out


class _SlotCore:
    """Shared by every handle of one group of streams: the slots, their owner and the current generation."""

    def __init__(self, engine: Engine, slots: np.ndarray):
        self.engine, self.slots, self.gen, self.closed = engine, slots, 0, False
        self._fin = weakref.finalize(self, _SlotCore._release, weakref.ref(engine), slots.copy())

    @staticmethod
    def _release(engine_ref, slots):
        eng = engine_ref()
        if eng is not None and eng._finalizer.alive:
            try:
                eng.release_slots(slots)
            except Exception:
                pass

    def close(self):
        self.closed = True
        self._fin()


class StreamSlots:
    """Opaque model state for the device-resident mode: a handle on the slots of B streams.

    The pipeline never inspects the model state (reference: tone/pipeline.py:143-147,172; the Triton client's
    state is a bare counter, dev/triton/client_wer.py:138-207), so a handle is a valid state.

    Contract (differs from the reference's value semantics, by design): the per-stream state lives in HBM and is
    advanced IN PLACE by ``forward``.  Every ``forward`` returns a NEW handle with the next generation number; a handle
    is single-use - passing an older one again (retry after an exception, branching, replaying) raises ``ValueError``
    instead of silently advancing the stream twice.  The slots go back to the pool when the last handle of the group is
    garbage collected, or deterministically through ``release()`` / use as a context manager."""

    def __init__(self, engine: Engine, slots: np.ndarray, _core: Optional[_SlotCore] = None):
        self._core = _core or _SlotCore(engine, slots)
        self.gen = self._core.gen

    @property
    def engine(self) -> Engine:
        return self._core.engine

    @property
    def slots(self) -> np.ndarray:
        return self._core.slots

    def _advance(self) -> "StreamSlots":
        c = self._core
        if c.closed:
            raise ValueError("this state handle was released")
        if self.gen != c.gen:
            raise ValueError(f"stale state handle (generation {self.gen}, the streams are at {c.gen}): "
                             "device-resident state is advanced in place and a handle is single-use")
        c.gen += 1
        return StreamSlots(c.engine, c.slots, c)

    def release(self):
        self._core.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()

    def __len__(self):
        return len(self._core.slots)


class B200StreamingCTCModel:
    """Acoustic model with the reference's interface, running on one B200.

    state_mode="numpy": ``forward`` takes/returns the reference's flat fp16 state ``(B, 219729)`` - each call
    imports the state into scratch slots, steps, and exports it again (parity / migration path; state crosses PCIe).
    state_mode="device": state is a :class:`StreamSlots` handle; per-stream state stays in HBM (throughput path) and
    is advanced in place - see :class:`StreamSlots` for the single-use handle contract.
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
        # the sample range [-32768; 32767] is checked by the C ABI while it narrows the PCM (TONE_ERANGE -> ValueError)
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
            nxt = state._advance()
            logprobs, _ = eng.step(state.slots, pcm, want_tokens=False)
            return [logprobs, nxt]
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
        eng.import_states(slots, state)
        logprobs, _ = eng.step(slots, pcm, want_tokens=False)
        return [logprobs, eng.export_states(slots)]
