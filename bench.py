#!/usr/bin/env python
"""Benchmark of the streaming acoustic-model step (BASELINE.json metric: streaming RTFx, audio-seconds per second).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path, one process per GPU
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (oracle port)

Workload (BASELINE.json configs[1]): 64 concurrent streams x 300 ms chunks per GPU, full step = log-mel + 16-layer
Conformer + CTC log-softmax/argmax, bf16 tensor-core GEMMs with fp32 accumulation, synthetic telephony audio, seeded
random-init weights of the configs/streaming_acoustic architecture.  A "step" advances every stream of the batch by
one chunk.  Streams shard across GPUs with no collective (weak scaling: 64 streams per GPU); torch.distributed (NCCL)
is used only for the barrier and the max-over-ranks of the timed region.

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with the PCM already in HBM; `e2e` = the same steps
through the C-ABI call with pinned HOST buffers (H2D of the PCM and D2H of logprobs+tokens inside the timed region).
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

FLOP_PER_CHUNK = {2400: 1_287_738_880, 3200: 1_631_636_224}   # SURVEY.md §8d / BASELINE.md §4 (2 x MAC, reference graph)
METRIC = "streaming RTFx (audio-sec/sec)"
UNIT = "audio-s/s"


def measured_traffic(streams: int, chunk: int):
    """DRAM bytes of one step from the committed ncu capture (profiles/r01_dram_traffic_B64.json:
    sum over the step's launches of dram__bytes_read.sum + dram__bytes_write.sum).  Only quoted for the workload it was
    measured on; ncu replays every launch with cold caches, so this is an upper bound of the live traffic."""
    p = os.path.join(ROOT, "profiles", "r01_dram_traffic_B64.json")
    if streams == 64 and chunk == 2400 and os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["dram_total_bytes"])
    return None


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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_port_throughput(streams: int, chunk: int, budget_s: float, warmup: int = 1, max_steps: int = 50, steps=None):
    """Time the oracle (CPU restatement of the reference algorithm) on the host cores: bounded sample of the workload."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tone_oracle as orc
    tb = importlib.import_module("t-one_b200")
    # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would pin us to one)
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    torch.set_num_threads(max(1, ncores))
    W = orc.to_torch(tb.weights.init_weights(0))
    n_chunks = 4
    pcm = tb.synth.telephony_pcm(streams, chunk * n_chunks, seed=1234)
    st = orc.zero_state(streams)
    times = []
    i = 0
    with torch.no_grad():
        for _ in range(warmup):
            _, st = orc.step(W, torch.from_numpy(pcm[:, :chunk]), st)
        t_all = time.perf_counter()
        while True:
            c = pcm[:, (i % n_chunks) * chunk:(i % n_chunks + 1) * chunk]
            t0 = time.perf_counter()
            _, st = orc.step(W, torch.from_numpy(c), st)
            times.append(time.perf_counter() - t0)
            i += 1
            if steps is not None:
                if i >= steps:
                    break
            elif time.perf_counter() - t_all > budget_s or i >= max_steps:
                break
    total = float(np.sum(times))
    return {"value": streams * chunk / 8000.0 * len(times) / total, "steps": len(times), "ms_per_step": 1e3 * total / len(times),
            "cores": int(torch.get_num_threads())}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    chunk = args.chunk
    r = cpu_port_throughput(args.streams, chunk, budget_s=1e9, warmup=max(1, min(args.warmup, 2)), steps=args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.streams} concurrent streams x {chunk * 1000 // 8000} ms chunks, reference algorithm "
                               "(torch fp32 CPU port of Tone.forward_for_export; ORT/model.onnx are not installable offline)",
                   "streams": args.streams, "chunk_samples": chunk},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": f"{r['steps']} steps of {args.streams} streams x {chunk} samples"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
    B, chunk, G = args.streams, args.chunk, args.groups
    K, Wm = args.steps, max(3, args.warmup)
    eng = tb.Engine(tb.weights.init_weights(0), chunk_samples=chunk, max_slots=B * G, max_batch=B, device=local)
    T = eng.T
    groups = [eng.alloc_slots(B) for _ in range(G)]
    n_distinct = 8
    pcm_all = tb.synth.telephony_pcm(B, chunk * n_distinct, seed=1234 + rank).reshape(B, n_distinct, chunk)
    pcm_all = np.ascontiguousarray(pcm_all.transpose(1, 0, 2))                  # (n_distinct, B, chunk)
    d_pcm = torch.from_numpy(pcm_all).cuda()
    d_slots = torch.from_numpy(np.stack(groups, 0)).cuda()
    d_lp = torch.empty((B, T, 35), dtype=torch.float32, device="cuda")
    d_tk = torch.empty((B, T), dtype=torch.int32, device="cuda")
    # a real (non-legacy) stream: events recorded on it bracket exactly the launched steps
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def dev_step(i):
        eng.step_device(B, d_slots[i % G].data_ptr(), d_pcm[i % n_distinct].data_ptr(), d_lp.data_ptr(), d_tk.data_ptr(),
                        stream)

    def host_step(i):
        eng.h_slots[:B] = groups[i % G]
        eng.h_pcm[:B] = pcm_all[i % n_distinct]
        return eng.step_pinned(B)

    # ---- leg 1: device-resident inputs, CUDA events on the launching stream
    for i in range(Wm):
        dev_step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    ev0.record()
    for i in range(K):
        dev_step(Wm + i)
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    dev_ms = ev0.elapsed_time(ev1)
    assert torch.isfinite(d_lp).all()

    # ---- leg 2: end to end through the C-ABI call with host buffers (H2D + step + D2H, synchronous)
    for i in range(Wm):
        host_step(i)
    barrier()
    lat = []
    t0 = time.perf_counter()
    for i in range(K):
        t1 = time.perf_counter()
        lp, tk = host_step(Wm + i)
        lat.append(time.perf_counter() - t1)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    assert np.isfinite(lp).all()

    if world > 1:
        t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0]), float(t[1])
    audio_s = world * B * (chunk / 8000.0) * K
    value = audio_s / (dev_ms / 1e3)
    e2e = audio_s / e2e_s
    launches = int(eng.info.launches_per_step if eng.info.launches_per_step else eng._get_info().launches_per_step)

    if rank == 0:
        peaks = measured_peaks()
        flops = FLOP_PER_CHUNK[chunk] * B                     # algorithmic FLOP of one step launch on one GPU
        step_s = dev_ms / 1e3 / K
        achieved = flops / step_s / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{B} concurrent streams x {chunk * 1000 // 8000} ms chunks per GPU, bf16 Conformer step + "
                                   "log-mel + CTC greedy (BASELINE.json configs[1])",
                       "streams_per_gpu": B, "chunk_samples": chunk, "frames_out": T, "parallelism": f"streams sharded x{world}, no collective",
                       "weights": "seeded random init, configs/streaming_acoustic architecture (71.7M params)",
                       "l2": f"no explicit flush: per-step working set = {eng.info.weight_bytes / 1e6:.0f} MB weights + "
                             f"{G} rotating {B}-stream state sets ({G * B * eng.info.state_bytes_per_slot / 1e6:.0f} MB) > 126 MB L2"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * chunk * 4 + B * 4,
                    "d2h_bytes_per_step": B * T * 35 * 4 + B * T * 4,
                    "latency_ms": {"p50": float(np.percentile(lat, 50) * 1e3), "p99": float(np.percentile(lat, 99) * 1e3)}},
            "gpu_launches": launches * K,
            "launches_per_step": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tflops"], "traffic": measured_traffic(B, chunk),
                         "traffic_note": "DRAM bytes per step launch, ncu cold-cache replay (profiles/r01_dram_traffic_B64.md); "
                                         f"algorithmic: {eng.info.weight_bytes / 1e6:.0f} MB weights + {B * 889_916 / 1e6:.0f} MB state/io",
                         "peak_source": peaks["src"],
                         "kernel": f"whole step graph (one launch = one {B}-stream step, {launches} kernels); per-kernel shares, "
                                   "ncu --set full of the GEMM kinds and the DRAM traffic in profiles/"},
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_port_throughput(B, chunk, budget_s=args.cpu_budget)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": f"{r['steps']} steps of {B} streams x {chunk} samples (oracle, torch fp32 CPU)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=64, help="concurrent streams per GPU (BASELINE configs[1]: 64)")
    ap.add_argument("--chunk", type=int, default=2400, choices=[2400, 3200])
    ap.add_argument("--groups", type=int, default=8, help="rotating stream sets so the state working set exceeds L2")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = min(args.steps, 400)   # ~0.4 s per 64-stream step on 8 host cores: keep the run to minutes
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
