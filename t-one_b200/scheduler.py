"""Stream scheduler / sequence batcher above the engine.

Two forms.  :class:`StreamServer` is the serving loop: the native, threaded ``tone_server`` of the C ABI (producers push
chunks from any thread; one native thread forms batches, keeps two tickets in flight and hands back completed batches).
:class:`StreamScheduler` is the same batching policy as a small synchronous Python class over ``Engine.step`` - handy for
tests and for callers that want to drive the steps themselves.

The reference has no scheduler of its own: it delegates batching and per-stream state residency to Triton -
``sequence_batching { oldest { max_candidate_sequences 4096 } , max_sequence_idle_microseconds 15 s }``
(reference: triton/model/config.pbtxt:26-31) and ``dynamic_batching { max_queue_delay_microseconds 10000 }`` with
``max_batch_size 16`` (reference: configs/streaming_acoustic/config.pbtxt:3,35-37).  This class plays that role in
process: stream -> slot lifecycle, per-stream FIFO of chunks, oldest-first batch formation with a queue-delay
window, idle reclaim.  It is single-threaded by design (one stepping thread per engine, like the C ABI).
"""
from __future__ import annotations

import time
from collections import OrderedDict, deque
from typing import Callable, Dict, Hashable, List, Optional, Tuple

import numpy as np


class SchedulerFull(RuntimeError):
    pass


class _Stream:
    __slots__ = ("slot", "queue", "last_active", "ending")

    def __init__(self, slot: int, now: float):
        self.slot, self.queue, self.last_active, self.ending = slot, deque(), now, False


class StreamScheduler:
    def __init__(self, engine, max_batch: Optional[int] = None, max_queue_delay_s: float = 0.010,
                 idle_timeout_s: float = 15.0, clock: Callable[[], float] = time.monotonic):
        self.engine = engine
        self.max_batch = int(max_batch or engine.info.max_batch)
        self.max_queue_delay_s = float(max_queue_delay_s)
        self.idle_timeout_s = float(idle_timeout_s)
        self.clock = clock
        self.streams: "OrderedDict[Hashable, _Stream]" = OrderedDict()
        self.steps = 0
        self.chunks = 0

    # ---- lifecycle (Triton: sequence_start / sequence_end flags, dev/triton/client_wer.py:180-187)
    def open(self, stream_id: Hashable) -> None:
        if stream_id in self.streams:
            raise KeyError(f"stream {stream_id!r} is already open")
        try:
            slot = int(self.engine.alloc_slots(1)[0])
        except MemoryError as ex:
            raise SchedulerFull(str(ex)) from ex
        self.streams[stream_id] = _Stream(slot, self.clock())

    def close(self, stream_id: Hashable) -> None:
        st = self.streams.pop(stream_id)
        self.engine.release_slots(np.array([st.slot], dtype=np.int32))

    def submit(self, stream_id: Hashable, chunk: np.ndarray, *, end: bool = False) -> None:
        """Queue one chunk (int32, chunk_samples) for a stream; opens the stream on first use.  ``end`` closes the
        stream after this chunk has been stepped."""
        chunk = np.asarray(chunk)
        if chunk.shape != (self.engine.chunk_samples,) or chunk.dtype != np.int32:
            raise ValueError(f"chunk must be int32 ({self.engine.chunk_samples},), got {chunk.dtype} {chunk.shape}")
        if stream_id not in self.streams:
            self.open(stream_id)
        st = self.streams[stream_id]
        if st.ending:
            raise RuntimeError(f"stream {stream_id!r} already received its last chunk")
        now = self.clock()
        st.queue.append((now, chunk))
        st.last_active = now
        st.ending = end

    # ---- batching
    def pending(self) -> int:
        return sum(1 for s in self.streams.values() if s.queue)

    def ready(self) -> bool:
        """A step is due when a full batch is waiting or the oldest queued chunk has waited max_queue_delay."""
        n, oldest = 0, None
        for s in self.streams.values():
            if s.queue:
                n += 1
                t = s.queue[0][0]
                oldest = t if oldest is None or t < oldest else oldest
        if n == 0:
            return False
        return n >= self.max_batch or (self.clock() - oldest) >= self.max_queue_delay_s

    def step(self) -> Dict[Hashable, Tuple[np.ndarray, np.ndarray]]:
        """Form one batch - at most one chunk per stream (the step is a recurrence), oldest head-of-queue first -
        run it, and return {stream_id: (logprobs (T,35), tokens (T,))}."""
        cand = [(s.queue[0][0], sid) for sid, s in self.streams.items() if s.queue]
        if not cand:
            return {}
        cand.sort(key=lambda x: x[0])
        ids = [sid for _, sid in cand[: self.max_batch]]
        slots = np.array([self.streams[sid].slot for sid in ids], dtype=np.int32)
        pcm = np.stack([self.streams[sid].queue.popleft()[1] for sid in ids], 0)
        logprobs, tokens = self.engine.step(slots, pcm)
        self.steps += 1
        self.chunks += len(ids)
        now = self.clock()
        out = {}
        for i, sid in enumerate(ids):
            out[sid] = (logprobs[i], tokens[i])
            st = self.streams[sid]
            st.last_active = now
            if st.ending and not st.queue:
                self.close(sid)
        return out

    def drain(self) -> List[Dict[Hashable, Tuple[np.ndarray, np.ndarray]]]:
        res = []
        while self.pending():
            res.append(self.step())
        return res

    def reclaim_idle(self) -> List[Hashable]:
        """Close streams with nothing queued that have been silent for idle_timeout_s (Triton: 15 s)."""
        now = self.clock()
        dead = [sid for sid, s in self.streams.items() if not s.queue and now - s.last_active >= self.idle_timeout_s]
        for sid in dead:
            self.close(sid)
        return dead


class StreamServer:
    """The native stream server (``tone_server_*`` of include/tone_b200.h) above one :class:`Engine`.

    ``push(stream_ids, pcm, last=None)`` is thread-safe and returns at once; ``poll()`` returns the next completed batch
    as a dict (``stream_ids``, ``seq``, ``latency_ms``, optional ``logprobs`` and ``phrases`` = list of (stream_id,
    start_frame, end_frame, label ids)); ``stats()`` the served counters and the push -> result latency percentiles.
    While a server exists the engine must not be stepped directly."""

    def __init__(self, engine, max_batch: int = 0, max_queue_delay_s: float = 0.010, idle_timeout_s: float = 15.0,
                 queue_depth: int = 4, outputs: Optional[int] = None, prewarm: bool = False):
        from . import model as M
        self._M, self.engine, self._lib = M, engine, engine._lib
        self.outputs = M.OUT_PHRASES if outputs is None else outputs
        cfg = M.ToneServerConfig(max_batch, int(max_queue_delay_s * 1e6), int(idle_timeout_s * 1e3), queue_depth, self.outputs, int(prewarm))
        self._h = M.C.c_void_p()
        rc = self._lib.tone_server_create(engine._h, M.C.byref(cfg), M.C.byref(self._h))
        if rc:
            M._raise(self._lib, rc)
        mb = max_batch or engine.info.max_batch
        self._ids = np.empty(mb, dtype=np.uint64)
        self._seq = np.empty(mb, dtype=np.int32)
        self._lat = np.empty(mb, dtype=np.float32)
        self._lp = np.empty((mb, engine.T, 35), dtype=np.float32) if self.outputs & M.OUT_LOGPROBS else None
        self._ph = np.empty(4 * mb, dtype=M.STREAM_PHRASE_DTYPE)
        self._text = np.empty(mb * 2304, dtype=np.uint8)

    def push(self, stream_ids, pcm, last=None) -> None:
        C = self._M.C
        ids = np.ascontiguousarray(stream_ids, dtype=np.uint64)
        x = np.ascontiguousarray(pcm, dtype=np.int16)
        if x.shape != (len(ids), self.engine.chunk_samples):
            raise ValueError(f"pcm must be int16 ({len(ids)}, {self.engine.chunk_samples}), got {x.shape}")
        fl = None if last is None else np.ascontiguousarray(last, dtype=np.uint8)
        rc = self._lib.tone_server_push(self._h, len(ids), ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                        x.ctypes.data_as(C.POINTER(C.c_int16)),
                                        fl.ctypes.data_as(C.POINTER(C.c_uint8)) if fl is not None else None)
        if rc:
            self._M._raise(self._lib, rc)

    def poll(self, timeout_s: float = 0.1):
        """-> dict for the next completed batch, or None on time-out."""
        C = self._M.C
        n, nph, ntext = C.c_int32(), C.c_int32(), C.c_int32()
        rc = self._lib.tone_server_poll(
            self._h, int(timeout_s * 1e3), C.byref(n), self._ids.ctypes.data_as(C.POINTER(C.c_uint64)),
            self._seq.ctypes.data_as(C.POINTER(C.c_int32)), self._lat.ctypes.data_as(C.POINTER(C.c_float)),
            self._lp.ctypes.data_as(C.POINTER(C.c_float)) if self._lp is not None else None,
            self._ph.ctypes.data_as(C.c_void_p), len(self._ph), C.byref(nph),
            self._text.ctypes.data_as(C.POINTER(C.c_uint8)), len(self._text), C.byref(ntext))
        if rc:
            self._M._raise(self._lib, rc)
        if n.value == 0:
            return None
        B = n.value
        out = {"stream_ids": self._ids[:B].copy(), "seq": self._seq[:B].copy(), "latency_ms": self._lat[:B].copy()}
        if self._lp is not None:
            out["logprobs"] = self._lp[:B].copy()
        if self.outputs & self._M.OUT_PHRASES:
            text = self._text[: ntext.value]
            out["phrases"] = [(int(r["stream_id"]), int(r["start_frame"]), int(r["end_frame"]),
                               text[r["text_offset"]: r["text_offset"] + r["text_len"]].copy()) for r in self._ph[: nph.value]]
        return out

    def stats(self) -> dict:
        st = self._M.ToneServerStats()
        rc = self._lib.tone_server_get_stats(self._h, self._M.C.byref(st))
        if rc:
            self._M._raise(self._lib, rc)
        return {k: getattr(st, k) for k, _ in st._fields_}

    def close(self) -> None:
        if self._h:
            self._lib.tone_server_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
