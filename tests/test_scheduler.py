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


@pytest.mark.gpu
def test_native_stream_server_matches_isolated_streams(tb, weights):
    """tone_server (native batcher + stepping thread, two tickets in flight): ragged arrivals from two producer threads.
    Per stream, the log-probs come back in order and equal the stream stepped alone; the phrases are exactly what the
    reference's splitter + greedy decoder give on those log-probs; end flags close streams; idle streams are reclaimed."""
    import threading
    import pipeline_oracle as po
    C, n_streams, n_chunks = 2400, 37, 9
    eng = tb.Engine(weights, chunk_samples=C, max_slots=48, max_batch=16)      # 6 of the 48 slots become the server's padding streams
    M = tb.model
    pcm = tb.synth.telephony_pcm(n_streams, C * n_chunks, seed=12)
    want = {}
    for sid in range(0, n_streams, 6):                       # isolated runs of a sample of the streams
        slot = eng.alloc_slots(1)
        want[sid] = [eng.step(slot, pcm[sid:sid + 1, i * C:(i + 1) * C])[0][0].copy() for i in range(n_chunks)]
        eng.release_slots(slot)
    srv = tb.scheduler.StreamServer(eng, max_batch=16, max_queue_delay_s=0.002, idle_timeout_s=0.3, queue_depth=4,
                                    outputs=M.OUT_LOGPROBS | M.OUT_PHRASES)
    try:
        def producer(ids):
            rng = np.random.default_rng(ids[0])
            pos = {i: 0 for i in ids}
            while any(p < n_chunks for p in pos.values()):
                pick = [i for i in ids if pos[i] < n_chunks and rng.random() < 0.6]
                if pick:
                    x = np.stack([pcm[i, pos[i] * C:(pos[i] + 1) * C] for i in pick]).astype(np.int16)
                    last = np.array([pos[i] == n_chunks - 1 and i % 2 == 0 for i in pick], dtype=np.uint8)   # even streams end
                    while True:
                        try:
                            srv.push(np.array(pick, dtype=np.uint64) + 1000, x, last)
                            break
                        except MemoryError:                  # a stream's queue is full: back off
                            import time
                            time.sleep(0.001)
                    for i in pick:
                        pos[i] += 1
        th = [threading.Thread(target=producer, args=(list(range(k, n_streams, 2)),)) for k in range(2)]
        for t in th:
            t.start()
        got = {i: [] for i in range(n_streams)}
        phrases = {i: [] for i in range(n_streams)}
        total, idle_polls = 0, 0
        while total < n_streams * n_chunks and idle_polls < 100:
            r = srv.poll(0.1)
            if r is None:
                idle_polls += 1
                continue
            assert len(set(r["stream_ids"].tolist())) == len(r["stream_ids"])       # one chunk per stream per step
            for k, sid in enumerate(r["stream_ids"].tolist()):
                assert r["seq"][k] == len(got[sid - 1000])                         # per-stream FIFO
                got[sid - 1000].append(r["logprobs"][k])
            for sid, a, b, ids in r["phrases"]:
                phrases[sid - 1000].append(("".join(po.LABELS[t] for t in ids), a, b))
            total += len(r["stream_ids"])
        for t in th:
            t.join()
        assert total == n_streams * n_chunks
        for sid, w in want.items():
            assert np.abs(np.stack(got[sid]) - np.stack(w)).max() < 3e-2           # batch composition changes tile selection
        for sid in range(n_streams):                                               # phrases: bit-exact on the served log-probs
            st, ref = None, []
            for i in range(n_chunks):
                ph, st = po.split(got[sid][i], st, is_last=(i == n_chunks - 1 and sid % 2 == 0))
                ref += [(po.greedy(p), a, b) for p, a, b in ph]
            assert phrases[sid] == ref
        st = srv.stats()
        assert st["chunks"] == n_streams * n_chunks and st["streams_opened"] == n_streams
        assert st["streams_closed"] == (n_streams + 1) // 2                        # the even streams sent an end flag
        assert 1.0 <= st["mean_batch"] <= 16 and st["latency_ms_p99"] > 0
        import time
        time.sleep(0.8)                                                            # odd streams: idle for > 0.3 s
        st = srv.stats()
        assert st["open_streams"] == 0 and st["streams_reclaimed"] == n_streams // 2
        # capacity: 43 new streams do not fit into 48 - 6 slots - all or nothing
        with pytest.raises(MemoryError):
            srv.push(np.arange(43, dtype=np.uint64), np.zeros((43, C), dtype=np.int16))
        assert srv.stats()["open_streams"] == 0
    finally:
        srv.close()
        eng.close()
