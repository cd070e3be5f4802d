"""Ragged-arrival serving benchmark: N real-time telephony streams, each delivering a 300 ms chunk at
Poisson-distributed instants (mean period 300 ms), pushed from a producer thread into the native stream server
(tone_server: 10 ms batching window, oldest-first, two tickets in flight, device-side phrase splitter), results polled by
a consumer thread.  Reports the served throughput (audio-seconds per second = concurrent real-time streams sustained),
the push -> result latency percentiles and the mean batch.

    python tools/gpu_serve_bench.py [streams ...] [max_batch=1024] [seconds=4] [window_ms=10] [saturate=0|1]
"""
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tb = importlib.import_module("t-one_b200")


def run(n_streams, max_batch=1024, seconds=4.0, window_ms=10.0, saturate=False, chunk=2400):
    eng = tb.Engine(tb.weights.init_weights(0), chunk_samples=chunk, max_slots=max(n_streams + 64, max_batch), max_batch=max_batch)
    srv = tb.scheduler.StreamServer(eng, max_batch=max_batch, max_queue_delay_s=window_ms / 1e3, queue_depth=4, prewarm=True)
    pool = tb.synth.telephony_pcm(256, chunk * 2, seed=7).reshape(512, chunk).astype(np.int16)
    period = chunk / 8000.0
    stop = threading.Event()
    pushed = [0, 0]     # chunks accepted, chunks refused (queue full)

    def producer():
        rng = np.random.default_rng(1)
        tick = 0.002
        t_next = time.perf_counter()
        while not stop.is_set():
            now = time.perf_counter()
            if now < t_next and not saturate:
                time.sleep(max(0.0, t_next - now))
            t_next += tick
            lam = n_streams * tick / period * (3.0 if saturate else 1.0)       # saturate: offer 3x real time
            k = min(int(rng.poisson(lam)), n_streams)
            if k == 0:
                continue
            ids = np.unique(rng.integers(0, n_streams, size=k)).astype(np.uint64)   # distinct streams, Poisson arrivals
            x = pool[rng.integers(0, len(pool), size=len(ids))]
            try:
                srv.push(ids, x)
                pushed[0] += len(ids)
            except MemoryError:
                pushed[1] += len(ids)

    served = [0, 0]     # chunks, phrases

    def consumer():
        while not stop.is_set() or True:
            r = srv.poll(0.05)
            if r is None:
                if stop.is_set():
                    return
                continue
            served[0] += len(r["stream_ids"])
            served[1] += len(r["phrases"])

    tp, tc = threading.Thread(target=producer), threading.Thread(target=consumer)
    t0 = time.perf_counter()
    tp.start()
    tc.start()
    time.sleep(1.0)                       # warm-up: graphs for the batch sizes in use, slots allocated
    s0, c0, t1 = srv.stats(), served[0], time.perf_counter()
    time.sleep(seconds)
    s1, c1, t2 = srv.stats(), served[0], time.perf_counter()
    stop.set()
    tp.join()
    tc.join()
    st = srv.stats()
    res = {
        "streams": n_streams, "max_batch": max_batch, "window_ms": window_ms, "mode": "saturated (3x offered)" if saturate else "real time",
        "offered_rtfx": (3.0 if saturate else 1.0) * n_streams,
        "served_rtfx": (c1 - c0) * period / (t2 - t1),
        "steps_per_s": (s1["steps"] - s0["steps"]) / (t2 - t1),
        "mean_batch": st["mean_batch"], "latency_ms_p50": st["latency_ms_p50"], "latency_ms_p99": st["latency_ms_p99"],
        "latency_ms_max": st["latency_ms_max"], "queue_ms_p50": st["queue_ms_p50"], "queue_ms_p99": st["queue_ms_p99"],
        "refused_chunks": pushed[1], "accepted_chunks": pushed[0], "phrases": served[1], "open_streams": st["open_streams"],
    }
    print(json.dumps(res), flush=True)
    srv.close()
    eng.close()
    return res


if __name__ == "__main__":
    ns, kw = [], {}
    for a in sys.argv[1:]:
        if "=" in a:
            k, v = a.split("=")
            kw[k] = float(v) if k in ("seconds", "window_ms") else int(v)
        else:
            ns.append(int(a))
    kw["saturate"] = bool(kw.get("saturate", 0))
    out = [run(n, **kw) for n in (ns or [20000, 60000])]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "serve_bench.jsonl"), "a") as f:
        for r in out:
            f.write(json.dumps(r) + "\n")
