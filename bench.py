#!/usr/bin/env python
"""Benchmark of the streaming acoustic-model step (BASELINE.json metric: streaming RTFx, audio-seconds per second).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path, one process per GPU
    python bench.py --impl reference --steps K --warmup W    # the reference's own torch model on the host cores

Workload (BASELINE.json configs[2]; at 8 GPUs this is configs[3]): 1024 concurrent streams x 300 ms chunks per GPU, full
step = log-mel + 16-layer Conformer + CTC log-softmax/argmax, bf16 tensor-core GEMMs with fp32 accumulation, synthetic
telephony audio, seeded random-init weights of the configs/streaming_acoustic architecture.  A "step" advances every
stream of the batch by one chunk.  Streams shard across GPUs with no collective (weak scaling: 1024 streams per GPU);
torch.distributed (NCCL) is used only for the barrier and the max-over-ranks of the timed region.

Prints ONE JSON line (rank 0):
  value          device-timed throughput, int16 PCM already resident in HBM (CUDA events on the launching stream)
  e2e            the same steps through the C ABI with pinned HOST buffers: tone_submit / tone_wait, two tickets in
                 flight (H2D of step i+1 and D2H of step i-1 overlap the kernels of step i); every step copies its int16
                 PCM + slot ids to the device and its full log-probs + argmax tokens back
  e2e.sync       the synchronous reference-shaped call tone_step (int32 PCM in, log-probs out), per-chunk p50 / p99
  e2e.greedy     pipelined, only finished phrases come back (device-side splitter + greedy decoder)
  latency_64     BASELINE configs[1] (64 streams per GPU): device ms/step and synchronous per-chunk latency
  roofline       whole step against the measured sustained bf16 tensor peak (MEASURED_PEAKS.json)
  cpu_baseline   the reference torch model (kind "reference", from baseline/_ref) or the oracle port on the host cores
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")     # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line

FLOP_PER_CHUNK = {2400: 1_287_738_880, 3200: 1_631_636_224}   # SURVEY.md §8d / BASELINE.md §4 (2 x MAC, reference graph)
METRIC = "streaming RTFx (audio-sec/sec)"
UNIT = "audio-s/s"


def workload_config(streams: int, chunk: int) -> dict:
    """Identical in both arms (the driver compares the `config` objects of the two lines)."""
    return {"workload": f"{streams} concurrent streams x {chunk * 1000 // 8000} ms chunks per GPU, log-mel + 16-layer Conformer "
                        "step + CTC log-softmax / greedy argmax (BASELINE.json configs[2]; x8 GPUs = configs[3])",
            "streams_per_gpu": streams, "chunk_samples": chunk, "frames_out": 10 if chunk == 2400 else 13,
            "weights": "seeded random init, configs/streaming_acoustic architecture (71.7M params)"}


def measured_traffic(streams: int, chunk: int):
    """DRAM bytes of one step (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu capture of this
    workload, if there is one.  Preferred: the LIVE figure (`ncu --replay-mode app-range --cache-control none` over a range
    of consecutive steps: warm caches, lanes concurrent; tools/summarize_range.py).  Fallback: the sum over the per-launch
    replay list, where ncu runs every kernel alone with cold caches - an upper bound of the reads."""
    for name, how in ((f"r02_dram_live_B{streams}.json", "live range of consecutive steps, ncu --replay-mode app-range, warm caches"),
                      (f"r02_dram_traffic_B{streams}.json", "sum over ncu's per-launch cold-cache replay"),
                      (f"r01_dram_traffic_B{streams}.json", "sum over ncu's per-launch cold-cache replay")):
        p = os.path.join(ROOT, "profiles", name)
        if chunk == 2400 and os.path.exists(p):
            with open(p) as f:
                return float(json.load(f)["dram_total_bytes"]), f"{name}: {how}"
    return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops"))), "hbm": float(d["hbm_gbs"]),
                "src": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1590.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.6 * (mx or 1)]
        return {"sm_mhz": float(np.median(busy if busy else sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_throughput(streams: int, chunk: int, steps=None, warmup: int = 1, budget_s: float = 15.0, sample_streams: int = 256):
    """Time the reference's CPU implementation of the step on the host cores, on a bounded sample of the workload:
    each step advances `min(streams, sample_streams)` of the workload's streams by one chunk (throughput is per
    audio-second, so the sample is representative).  kind "reference" = the UNMODIFIED reference torch model
    (tone.nn.model.Tone.forward_for_export, fp32) from baseline/_ref; kind "port" = oracle/tone_oracle.py when no
    reference tree is reachable."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refimport
    tb = importlib.import_module("t-one_b200")
    try:
        ncores = len(os.sched_getaffinity(0))   # torchrun exports OMP_NUM_THREADS=1, which would pin us to one core
    except AttributeError:
        ncores = os.cpu_count() or 1
    torch.set_num_threads(max(1, ncores))
    weights = tb.weights.init_weights(0)
    S = min(streams, sample_streams)
    n_chunks = 4
    pcm = tb.synth.telephony_pcm(S, chunk * n_chunks, seed=1234)
    kind = "port"
    model = None
    if chunk == 2400:       # the reference's public interface is the 300 ms chunk (tone/onnx_wrapper.py:32)
        try:
            model = refimport.ReferenceStreamingModel(weights)
            kind = "reference"
        except Exception:
            model = None
    if model is not None:
        state = [None]

        def one(i):
            c = pcm[:, (i % n_chunks) * chunk:(i % n_chunks + 1) * chunk]
            _, state[0] = model.forward(np.ascontiguousarray(c[:, :, None]).astype(np.int32), state[0])
    else:
        import tone_oracle as orc
        W = orc.to_torch(weights)
        state = [orc.zero_state(S)]

        def one(i):
            c = pcm[:, (i % n_chunks) * chunk:(i % n_chunks + 1) * chunk]
            with torch.no_grad():
                _, state[0] = orc.step(W, torch.from_numpy(c), state[0])
    for i in range(warmup):
        one(i)
    times, i, t_all = [], 0, time.perf_counter()
    while True:
        t0 = time.perf_counter()
        one(warmup + i)
        times.append(time.perf_counter() - t0)
        i += 1
        if steps is not None:
            if i >= steps:
                break
        elif time.perf_counter() - t_all > budget_s or i >= 50:
            break
    total = float(np.sum(times))
    return {"value": S * chunk / 8000.0 * len(times) / total, "steps": len(times), "ms_per_step": 1e3 * total / len(times),
            "cores": int(torch.get_num_threads()), "kind": kind,
            "sample": f"{len(times)} steps of {S} of the {streams} streams x {chunk} samples, "
                      + ("reference torch model Tone.forward_for_export fp32 (baseline/_ref)" if kind == "reference"
                         else "oracle port (torch fp32 restatement)")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    chunk = args.chunk
    r = cpu_throughput(args.streams, chunk, steps=max(1, min(args.steps, 100)), warmup=max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.streams, chunk),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "ONNX Runtime and model.onnx cannot be installed offline; this arm runs the torch graph that "
                "tone/scripts/export.py traces into model.onnx, on all host cores",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def bench_engine(tb, torch, B, chunk, K, Wm, local, rank, barrier, greedy=True):
    """All legs for one engine size.  Returns a dict of raw timings (per rank)."""
    G = 2 if B >= 512 else 8                                   # rotating stream sets: state working set >> 126 MB L2
    eng = tb.Engine(tb.weights.init_weights(0), chunk_samples=chunk, max_slots=B * G, max_batch=B, device=local)
    M = tb.model
    T = eng.T
    groups = [eng.alloc_slots(B) for _ in range(G)]
    n_distinct = 4
    pcm_all = tb.synth.telephony_pcm(min(B, 256), chunk * n_distinct, seed=1234 + rank).reshape(-1, n_distinct, chunk)
    pcm_all = np.ascontiguousarray(np.tile(pcm_all, ((B + pcm_all.shape[0] - 1) // pcm_all.shape[0], 1, 1))[:B].transpose(1, 0, 2))
    pcm16 = pcm_all.astype(np.int16)                           # (n_distinct, B, chunk)
    d_pcm = torch.from_numpy(pcm16).cuda()
    d_lp = torch.empty((B, T, 35), dtype=torch.float32, device="cuda")
    d_tk = torch.empty((B, T), dtype=torch.int32, device="cuda")
    tstream = torch.cuda.Stream()     # a real (non-legacy) stream: events recorded on it bracket exactly the launched steps
    stream = tstream.cuda_stream
    assert stream != 0
    out = {}

    def dev_step(i):
        eng.step_device(groups[i % G], d_pcm[i % n_distinct].data_ptr(), M.PCM_I16, d_lp.data_ptr(), d_tk.data_ptr(), stream)

    # ---- leg 1: device-resident inputs, CUDA events on the launching stream
    with torch.cuda.stream(tstream):
        for i in range(Wm):
            dev_step(i)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        ev0.record(tstream)
        for i in range(K):
            dev_step(Wm + i)
        ev1.record(tstream)
        barrier()
        out["wall_s"] = time.perf_counter() - t_wall
        out["dev_ms"] = ev0.elapsed_time(ev1)
    assert torch.isfinite(d_lp).all()
    out["launches"] = int(eng._get_info().launches_per_step)

    # ---- leg 2: end to end, pipelined tickets with pinned host buffers (H2D int16 PCM + slots, D2H logprobs + tokens)
    def pipelined(outputs, n_steps, warm):
        pending, lat, t_sub = None, [], {}
        res = None
        for i in range(warm + n_steps):
            if i == warm:
                if pending is not None:
                    eng.wait(pending)
                    pending = None
                barrier()
                t0 = time.perf_counter()
            s, p, l = eng.next_staging(B)
            s[:] = groups[i % G]
            p[:] = pcm16[i % n_distinct]                       # the serving loop writes the chunk into the pinned staging
            l[:] = 0
            t = eng.submit(s, p, outputs, l)
            t_sub[t.id] = time.perf_counter()
            if pending is not None:
                res = eng.wait(pending)
                lat.append(time.perf_counter() - t_sub.pop(pending.id))
            pending = t
        res = eng.wait(pending)
        lat.append(time.perf_counter() - t_sub.pop(pending.id))
        barrier()
        return time.perf_counter() - t0, lat[-n_steps:], res

    e2e_s, lat_pipe, res = pipelined(M.OUT_LOGPROBS | M.OUT_TOKENS, K, Wm)
    assert np.isfinite(res["logprobs"]).all()
    out["e2e_s"], out["lat_pipe"] = e2e_s, lat_pipe
    if greedy:
        g_s, _, _ = pipelined(M.OUT_PHRASES, K, Wm)
        out["greedy_s"] = g_s

    # ---- leg 3: the synchronous reference-shaped call (int32 PCM in, logprobs + tokens out), per-chunk latency
    pcm32 = pcm_all.astype(np.int32)
    lat = []
    for i in range(Wm + K):
        t1 = time.perf_counter()
        eng.step(groups[i % G], pcm32[i % n_distinct])
        if i >= Wm:
            lat.append(time.perf_counter() - t1)
    out["lat_sync"] = lat
    out["info"] = {"weight_bytes": int(eng.info.weight_bytes), "state_bytes_per_slot": int(eng.info.state_bytes_per_slot),
                   "T": T, "G": G}
    eng.close()
    return out


def serving_leg(tb, B, chunk, local, barrier, seconds=2.0):
    """Saturated throughput THROUGH THE STREAM SERVER (tone_server: native batcher thread, 10 ms window, two tickets in
    flight, device-side phrase splitter): 4 B open streams, a producer thread pushing chunks as fast as the queues take
    them, a consumer thread polling completed batches.  -> (served audio-s/s, server stats)."""
    import threading
    eng = tb.Engine(tb.weights.init_weights(0), chunk_samples=chunk, max_slots=4 * B + 64, max_batch=B, device=local)
    srv = tb.scheduler.StreamServer(eng, max_batch=B, max_queue_delay_s=0.010, queue_depth=4)
    pool = tb.synth.telephony_pcm(min(B, 256), chunk * 2, seed=7).reshape(-1, chunk).astype(np.int16)
    pool = np.ascontiguousarray(np.tile(pool, ((B + len(pool) - 1) // len(pool), 1))[:B])
    stop, served = threading.Event(), [0]

    def producer():
        k = 0
        while not stop.is_set():
            ids = (np.arange(B, dtype=np.uint64) + (k % 4) * B)
            try:
                srv.push(ids, pool)
                k += 1
            except MemoryError:
                time.sleep(0.0005)

    def consumer():
        while True:
            r = srv.poll(0.05)
            if r is None:
                if stop.is_set():
                    return
                continue
            served[0] += len(r["stream_ids"])

    tp, tc = threading.Thread(target=producer), threading.Thread(target=consumer)
    tp.start()
    tc.start()
    time.sleep(0.7)
    barrier()
    c0, t0 = served[0], time.perf_counter()
    time.sleep(seconds)
    c1, t1 = served[0], time.perf_counter()
    stop.set()
    tp.join()
    tc.join()
    st = srv.stats()
    srv.close()
    eng.close()
    return (c1 - c0) * (chunk / 8000.0) / (t1 - t0), st


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    tb = importlib.import_module("t-one_b200")
    B, chunk = args.streams, args.chunk
    K, Wm = args.steps, max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    r = bench_engine(tb, torch, B, chunk, K, Wm, local, rank, barrier)
    clocks = sampler.stop() if rank == 0 else None
    r64 = bench_engine(tb, torch, 64, chunk, max(K, 50), Wm, local, rank, barrier, greedy=False) if B != 64 else None
    served, served_stats = None, None
    if not args.no_serving:
        try:
            served, served_stats = serving_leg(tb, B, chunk, local, (lambda: dist.barrier()) if world > 1 else (lambda: None))
        except Exception as ex:  # noqa: BLE001  (the serving leg never takes the headline down with it)
            served, served_stats = None, {"error": f"{type(ex).__name__}: {ex}"}

    def maxr(*vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    dev_ms, e2e_s, greedy_s = maxr(r["dev_ms"], r["e2e_s"], r["greedy_s"])
    if world > 1:
        t = torch.tensor([served if served is not None else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        served_total = float(t[0])
    else:
        served_total = served
    audio_s = world * B * (chunk / 8000.0) * K
    value = audio_s / (dev_ms / 1e3)
    T, info = r["info"]["T"], r["info"]
    launches = r["launches"]

    if rank == 0:
        peaks = measured_peaks()
        flops = FLOP_PER_CHUNK[chunk] * B                     # algorithmic FLOP of one step launch on one GPU
        step_s = dev_ms / 1e3 / K
        achieved = flops / step_s / 1e12
        traffic, traffic_src = measured_traffic(B, chunk)
        pct = lambda a, q: float(np.percentile(a, q) * 1e3)   # noqa: E731
        cfg = workload_config(B, chunk)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": cfg,
            "parallelism": f"streams sharded x{world} ({world * B} streams), no collective",
            "l2": f"no explicit flush: per-step working set = {info['weight_bytes'] / 1e6:.0f} MB weights + {info['G']} rotating "
                  f"{B}-stream state sets ({info['G'] * B * info['state_bytes_per_slot'] / 1e6:.0f} MB) > 126 MB L2",
            "e2e": {"value": audio_s / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": B * chunk * 2 + B * 4 + B, "d2h_bytes_per_step": B * T * 35 * 4 + B * T * 4,
                    "mode": "tone_submit / tone_wait, 2 tickets in flight, pinned int16 PCM in, fp32 logprobs + tokens out",
                    "ticket_latency_ms": {"p50": pct(r["lat_pipe"], 50), "p99": pct(r["lat_pipe"], 99)},
                    "frac_of_device": (audio_s / e2e_s) / value,
                    "sync": {"value": B * (chunk / 8000.0) / float(np.mean(r["lat_sync"])) * world,
                             "mode": "tone_step (int32 PCM in, logprobs + tokens out), one chunk at a time",
                             "latency_ms": {"p50": pct(r["lat_sync"], 50), "p99": pct(r["lat_sync"], 99)}},
                    "greedy": {"value": audio_s / greedy_s, "d2h_bytes_per_step": 16 + B * 4 * 20 + 32768,
                               "mode": "pipelined, device-side phrase splitter + greedy decode, only finished phrases return"}},
            "gpu_launches": launches * K,
            "launches_per_step": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tflops"], "traffic": traffic,
                         "traffic_note": (f"DRAM bytes per step launch (profiles/{traffic_src}); " if traffic else "")
                                         + f"algorithmic: {info['weight_bytes'] / 1e6:.0f} MB weights + {B * 889_916 / 1e6:.0f} MB state/io",
                         "peak_source": peaks["src"],
                         "kernel": f"whole step graph (one launch = one {B}-stream step, {launches} kernels); per-kernel shares, "
                                   "ncu --set full of the GEMM kinds and the DRAM traffic in profiles/"},
            "clocks": clocks,
            "wall_s_timed_region": r["wall_s"],
        }
        if r64 is not None:
            k64 = len(r64["lat_sync"])
            line["latency_64"] = {
                "workload": "64 concurrent streams x 300 ms chunks per GPU (BASELINE.json configs[1])" if chunk == 2400 else "64 streams",
                "ms_per_step_device": r64["dev_ms"] / k64, "value": 64 * (chunk / 8000.0) * k64 / (r64["dev_ms"] / 1e3),
                "e2e_pipelined": 64 * (chunk / 8000.0) * k64 / r64["e2e_s"],
                "sync_latency_ms": {"p50": pct(r64["lat_sync"], 50), "p99": pct(r64["lat_sync"], 99)},
                "launches_per_step": r64["launches"],
                "roofline_frac": FLOP_PER_CHUNK[chunk] * 64 / (r64["dev_ms"] / 1e3 / k64) / 1e12 / peaks["tflops"]}
        if served_stats is not None:
            line["serving"] = {"value": served_total, "unit": UNIT, "frac_of_device": (served_total or 0.0) / value,
                               "mode": f"tone_server (native batcher thread, 10 ms window, oldest-first, two tickets in flight, device-side "
                                       f"phrase splitter), {4 * B} open streams per GPU, saturated producer; sum over ranks",
                               "mean_batch": served_stats.get("mean_batch"), "error": served_stats.get("error")}
        if world == 1 and not args.no_cpu_baseline:
            c = cpu_throughput(B, chunk, budget_s=args.cpu_budget)
            line["cpu_baseline"] = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"], "sample": c["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="concurrent streams per GPU (BASELINE configs[2]: 1024)")
    ap.add_argument("--chunk", type=int, default=2400, choices=[2400, 3200])
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-serving", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
