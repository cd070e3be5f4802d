"""State wire formats of one stream (host side, numpy): conversions between

* the flat fp16 vector of the public model (``state`` / ``state_next``, 219,729 elements;
  reference: tone/onnx_wrapper.py:34, configs/streaming_acoustic/config.pbtxt:12-33), whose element order is the
  seven tensors of ``get_initial_state`` (reference: tone/nn/model.py:259-267), and
* the newer three-tensor Triton cache layout (reference: tone/scripts/export.py:293-376,
  triton/model/config.pbtxt:44-66): ``cache_last_time (18,384,30)``, ``cache_last_channel (32,8,50)``,
  ``cache_last_chan_len`` int64.

``Engine.export_state`` / ``import_state`` speak the flat format; these helpers let the engine sit behind the newer
Triton ensemble as well.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .arch import DEFAULT_ARCH, ToneArch

KEYS = ("preproc", "mhsa", "conv", "mhsa_len", "sub1", "sub2", "reduction")


def split_flat(flat: np.ndarray, arch: ToneArch = DEFAULT_ARCH) -> Dict[str, np.ndarray]:
    """(B, 219729) -> the seven state tensors (views, batch first)."""
    flat = np.asarray(flat)
    if flat.ndim != 2 or flat.shape[1] != arch.state_size:
        raise ValueError(f"flat state must be (B, {arch.state_size}), got {flat.shape}")
    out, o = {}, 0
    for name, shp in arch.state_layout():
        n = int(np.prod(shp))
        out[name] = flat[:, o:o + n].reshape((flat.shape[0],) + tuple(shp))
        o += n
    return out


def join_flat(parts: Dict[str, np.ndarray], arch: ToneArch = DEFAULT_ARCH) -> np.ndarray:
    B = parts["preproc"].shape[0]
    return np.concatenate([np.asarray(parts[k], dtype=np.float16).reshape(B, -1) for k in KEYS], axis=1)


def _tail_geometry(arch: ToneArch) -> Tuple[int, int, int, int]:
    c1, c2, tbase = arch.sub_channels[0], arch.sub2_rows, arch.sub_f1            # (32, 8, 44)
    need = arch.pre_state + arch.sub1_rows * arch.n_mels + arch.d_model * arch.red_state   # 80 + 640 + 384
    tpad = -(-need // (c1 * c2)) or 2                                            # export.py:222-225
    tpad += tpad % 2
    return c1, c2, tbase, tpad


def flat_to_triton(flat: np.ndarray, arch: ToneArch = DEFAULT_ARCH):
    """(B, 219729) fp16 -> (cache_last_time (B,18,384,30) fp16, cache_last_channel (B,32,8,50) fp16,
    cache_last_chan_len (B,) int64); reference: tone/scripts/export.py:335-376."""
    p = split_flat(np.asarray(flat, dtype=np.float16), arch)
    B = p["preproc"].shape[0]
    time = np.concatenate([p["mhsa"].transpose(0, 1, 3, 2), p["conv"]], axis=1)          # (B, 2+16, 384, 30)
    c1, c2, tbase, tpad = _tail_geometry(arch)
    tail = np.zeros((B, c1 * c2 * tpad), dtype=np.float16)
    flat_tail = np.concatenate([p["preproc"].reshape(B, -1), p["sub1"].reshape(B, -1), p["reduction"].reshape(B, -1)], 1)
    tail[:, : flat_tail.shape[1]] = flat_tail
    chan = np.concatenate([p["sub2"], tail.reshape(B, c1, c2, tpad)], axis=3)            # (B, 32, 8, 44+6)
    length = np.rint(p["mhsa_len"].reshape(B).astype(np.float32)).astype(np.int64)
    return np.ascontiguousarray(time), np.ascontiguousarray(chan), length


def triton_to_flat(cache_last_time, cache_last_channel, cache_last_chan_len, arch: ToneArch = DEFAULT_ARCH) -> np.ndarray:
    """Inverse of :func:`flat_to_triton`; reference: tone/scripts/export.py:293-333."""
    time = np.asarray(cache_last_time, dtype=np.float16)
    chan = np.asarray(cache_last_channel, dtype=np.float16)
    B = time.shape[0]
    n_m = arch.n_mhsa_stateful
    c1, c2, tbase, tpad = _tail_geometry(arch)
    if time.shape[1:] != (n_m + arch.n_layers, arch.d_model, arch.conv_state):
        raise ValueError(f"cache_last_time has shape {time.shape}")
    if chan.shape[1:] != (c1, c2, tbase + tpad):
        raise ValueError(f"cache_last_channel has shape {chan.shape}")
    tail = chan[:, :, :, tbase:].reshape(B, -1)
    n_pre, n_s1 = arch.pre_state, arch.sub1_rows * arch.n_mels
    n_red = arch.d_model * arch.red_state
    parts = {
        "preproc": tail[:, :n_pre],
        "mhsa": time[:, :n_m].transpose(0, 1, 3, 2),
        "conv": time[:, n_m:],
        "mhsa_len": np.asarray(cache_last_chan_len).reshape(B, 1).astype(np.float16),
        "sub1": tail[:, n_pre:n_pre + n_s1].reshape(B, 1, arch.sub1_rows, arch.n_mels),
        "sub2": chan[:, :, :, :tbase],
        "reduction": tail[:, n_pre + n_s1:n_pre + n_s1 + n_red].reshape(B, arch.d_model, arch.red_state),
    }
    return join_flat(parts, arch)
