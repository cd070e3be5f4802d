"""Stream scheduler: slot lifecycle, oldest-first batching with a delay window, per-stream FIFO, idle reclaim."""
import types

import numpy as np
import pytest


class FakeEngine:
    """Records what it is asked to step; logprobs[b, :, 0] = slot id, [b, :, 1] = first sample of the chunk."""

    def __init__(self, max_slots=4, max_batch=3, chunk=2400, T=10):
        self.info = types.SimpleNamespace(max_batch=max_batch, max_slots=max_slots)
        self.chunk_samples, self.T = chunk, T
        self.free = list(range(max_slots))[::-1]
        self.batches = []

    def alloc_slots(self, n):
        if n > len(self.free):
            raise MemoryError("no free slots")
        return np.array([self.free.pop() for _ in range(n)], dtype=np.int32)

    def release_slots(self, slots):
        self.free.extend(int(s) for s in slots)

    def step(self, slots, pcm):
        self.batches.append(list(map(int, slots)))
        lp = np.zeros((len(slots), self.T, 35), dtype=np.float32)
        lp[:, :, 0] = np.asarray(slots)[:, None]
        lp[:, :, 1] = pcm[:, :1]
        return lp, np.zeros((len(slots), self.T), dtype=np.int32)


class Clock:
    def __init__(self):
        self.t = 0.0

    def __call__(self):
        return self.t


def chunk(v, n=2400):
    return np.full(n, v, dtype=np.int32)


def test_batching_window_and_oldest_first(tb):
    clk, eng = Clock(), FakeEngine(max_slots=5, max_batch=3)
    s = tb.scheduler.StreamScheduler(eng, max_queue_delay_s=0.010, clock=clk)
    s.submit("a", chunk(1))
    clk.t = 0.002
    s.submit("b", chunk(2))
    assert not s.ready()                       # 2 < max_batch and the oldest has waited 2 ms < 10 ms
    clk.t = 0.011
    assert s.ready()                           # delay window expired
    clk.t = 0.012
    s.submit("c", chunk(3))
    s.submit("d", chunk(4))
    assert s.ready() and s.pending() == 4
    out = s.step()                             # oldest three heads: a, b, c
    assert sorted(out) == ["a", "b", "c"]
    assert out["b"][0][0, 1] == 2 and out["b"][0].shape == (10, 35)
    assert s.pending() == 1
    out = s.step()
    assert list(out) == ["d"]
    assert s.step() == {}


def test_one_chunk_per_stream_per_step_and_fifo(tb):
    clk, eng = Clock(), FakeEngine()
    s = tb.scheduler.StreamScheduler(eng, clock=clk)
    for i in range(3):
        s.submit("x", chunk(10 + i))
    s.submit("y", chunk(99))
    outs = s.drain()
    assert [sorted(o) for o in outs] == [["x", "y"], ["x"], ["x"]]
    assert [o["x"][0][0, 1] for o in outs] == [10, 11, 12]          # FIFO per stream
    assert all(len(set(b)) == len(b) for b in eng.batches)          # a slot never appears twice in a batch


def test_slot_lifecycle_capacity_end_flag_and_idle_reclaim(tb):
    clk, eng = Clock(), FakeEngine(max_slots=2, max_batch=2)
    s = tb.scheduler.StreamScheduler(eng, idle_timeout_s=15.0, clock=clk)
    s.submit(1, chunk(1))
    s.submit(2, chunk(2), end=True)
    with pytest.raises(tb.scheduler.SchedulerFull):
        s.submit(3, chunk(3))
    s.step()
    assert 2 not in s.streams and len(eng.free) == 1                # end flag released the slot after its last chunk
    s.submit(3, chunk(3))                                           # the freed slot is reusable
    with pytest.raises(RuntimeError):
        s.submit(3, chunk(4), end=True) or s.submit(3, chunk(5))
    s.drain()
    clk.t = 14.0
    assert s.reclaim_idle() == []
    clk.t = 30.0
    assert s.reclaim_idle() == [1]                                  # 15 s idle -> reclaimed (Triton default)
    assert len(eng.free) == 2
    with pytest.raises(ValueError):
        s.submit(9, np.zeros(100, dtype=np.int32))


@pytest.mark.gpu
def test_ragged_arrivals_match_isolated_streams(tb, weights):
    """Streams joining/leaving at different times through the scheduler == each stream stepped alone."""
    C = 2400
    eng = tb.Engine(weights, chunk_samples=C, max_slots=8, max_batch=4)
    pcm = tb.synth.telephony_pcm(3, C * 5, seed=12)
    # isolated reference runs
    want = {}
    for sid in range(3):
        slot = eng.alloc_slots(1)
        want[sid] = [eng.step(slot, pcm[sid:sid + 1, i * C:(i + 1) * C])[0][0].copy() for i in range(5)]
        eng.release_slots(slot)
    sch = tb.scheduler.StreamScheduler(eng, max_batch=4, max_queue_delay_s=0.0)
    got = {0: [], 1: [], 2: []}
    arrivals = [[0], [0, 1], [0, 1, 2], [0, 1, 2], [0, 1, 2], [1, 2], [2]]     # stream k starts k steps late
    pos = {0: 0, 1: 0, 2: 0}
    for active in arrivals:
        for sid in active:
            sch.submit(sid, pcm[sid, pos[sid] * C:(pos[sid] + 1) * C].astype(np.int32), end=(pos[sid] == 4))
            pos[sid] += 1
        for sid, (lp, _) in sch.step().items():
            got[sid].append(lp.copy())
    assert not sch.streams                                           # every stream ended and released its slot
    for sid in range(3):
        assert len(got[sid]) == 5
        assert np.abs(np.stack(got[sid]) - np.stack(want[sid])).max() < 3e-2
    eng.close()
