"""Host-side logic on CPU: C-ABI library exports, model-class validation, stream sharding over 2 gloo ranks."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol(tb):
    header = open(os.path.join(ROOT, "include", "tone_b200.h")).read()
    declared = set(re.findall(r"\b(tone_[a-z_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = tb.load_library()
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/tone_b200.h but not exported"
    assert declared == set(tb.model.SYMBOLS)
    assert isinstance(lib, ctypes.CDLL)


def test_product_path_fails_loudly_without_gpu(tb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Exception) as ei:
        tb.Engine(None, max_slots=1)
    assert "CUDA" in str(ei.value) or "cuda" in str(ei.value)


def test_missing_library_is_an_error_not_a_fallback(tb, tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tb.load_library(str(tmp_path / "nope.so"))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "t-one_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "tone_oracle" not in src and "oracle/" not in src, f"{f} references the oracle"


def test_model_class_constants_mirror_reference(tb):
    m = tb.B200StreamingCTCModel
    # tone/onnx_wrapper.py:30-34
    assert (m.SAMPLE_RATE, m.MEAN_TIME_BIAS, m.AUDIO_CHUNK_SAMPLES, m.FRAME_SIZE, m.STATE_SIZE) == \
        (8000, 0.33, 2400, 0.03, 219729)


def test_weights_loader_accepts_tone_prefix_and_checks_shapes(tb, weights):
    pref = {"tone." + k: v for k, v in weights.items()}
    got = tb.weights.from_state_dict(pref)
    assert list(got) == list(weights)
    bad = dict(weights)
    bad["decoder.decoder_layers.0.bias"] = np.zeros(3, np.float32)
    with pytest.raises(ValueError):
        tb.weights.from_state_dict(bad)
    del bad["decoder.decoder_layers.0.bias"]
    with pytest.raises(KeyError):
        tb.weights.from_state_dict(bad)


def test_partition_covers_every_stream_once(tb):
    sh = tb.sharding
    for n, g in [(8192, 8), (8192, 4), (1000, 3), (5, 8)]:
        parts = sh.partition(n, g)
        allids = np.concatenate(parts)
        assert sorted(allids.tolist()) == list(range(n))
        assert all(sh.owner_of(s, g) == r for r, p in enumerate(parts) for s in p[:5])
    b = sh.batches(np.arange(2500), 1024)
    assert [len(x) for x in b] == [1024, 1024, 452]


_WORKER = r"""
import importlib, os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
tb = importlib.import_module("t-one_b200")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
mine = tb.sharding.local_streams(101, rank, world)
gathered = [None] * world
dist.all_gather_object(gathered, mine.tolist())
assert sorted(sum(gathered, [])) == list(range(101))
# each rank "processed" len(mine) streams x 0.3 s x 10 steps in (1 + rank) seconds
thr, total, tmax = tb.sharding.aggregate_throughput(len(mine) * 0.3 * 10, 1.0 + rank, dist)
assert abs(total - 101 * 3.0) < 1e-9 and tmax == float(world) and abs(thr - 101 * 3.0 / world) < 1e-9
dist.barrier()
dist.destroy_process_group()
open(os.path.join({out!r}, "rank%d.ok" % rank), "w").write("ok")
"""


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, out=str(tmp_path)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists()
