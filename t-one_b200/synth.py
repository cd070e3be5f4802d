"""Synthetic telephony-shaped audio for parity tests and benchmarks.

8 kHz mono, int16-range samples carried as int32, as the reference's acoustic model expects
(reference: tone/onnx_wrapper.py:30-33,100-113).  Per stream: a voiced harmonic source with a
4 Hz syllabic envelope, alternating 1-3 s "speech" and 0.3-1.5 s near-silence, plus noise.
"""
from __future__ import annotations

import numpy as np


def telephony_pcm(n_streams: int, n_samples: int, seed: int = 1234, sample_rate: int = 8000) -> np.ndarray:
    """(n_streams, n_samples) int32 in [-32768, 32767]."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    out = np.empty((n_streams, n_samples), dtype=np.int32)
    for s in range(n_streams):
        f0 = rng.uniform(80.0, 300.0)
        sig = np.zeros(n_samples, dtype=np.float64)
        k = 1
        while k * f0 < 3400.0:
            sig += np.sin(2 * np.pi * k * f0 * t + rng.uniform(0, 2 * np.pi)) / k
            k += 1
        sig *= 0.5 * (1.0 + np.sin(2 * np.pi * 4.0 * t + rng.uniform(0, 2 * np.pi)))
        gate = np.zeros(n_samples, dtype=np.float64)
        pos, speech = 0, bool(rng.integers(0, 2))
        while pos < n_samples:
            dur = rng.uniform(1.0, 3.0) if speech else rng.uniform(0.3, 1.5)
            n = int(dur * sample_rate)
            gate[pos:pos + n] = 1.0 if speech else 0.02
            pos += n
            speech = not speech
        sig = sig * gate
        peak = np.max(np.abs(sig)) + 1e-9
        sig = sig / peak * 8000.0 + rng.normal(0.0, 200.0, n_samples)
        out[s] = np.clip(np.rint(sig), -32768, 32767).astype(np.int32)
    return out
