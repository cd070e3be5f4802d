"""Parameter inventory, seeded synthetic weights and frontend constants.

Names and shapes follow the reference ``Tone`` module's ``state_dict``
(reference: tone/nn/model.py:31-42, tone/nn/modules/conformer_blocks.py:364-401,452-466,
501-553,747-797,850-867, tone/nn/modules/submodules.py:181-201, tone/nn/modules/conformer.py:325-336),
so a real checkpoint (``ToneForCTC`` keys carry an extra ``tone.`` prefix,
tone/training/model_wrapper.py:146,156) loads through :func:`from_state_dict`.

There is no network in the build or GPU environment, so benchmarks and parity tests use
:func:`init_weights`: a seeded numpy draw in the reference's default-init ranges with the
norm gains, BatchNorm statistics and biases perturbed away from identity (default init
leaves BN = identity and gains = 1, which would hide folding bugs).
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict

import numpy as np

from .arch import DEFAULT_ARCH, ToneArch


def param_shapes(arch: ToneArch = DEFAULT_ARCH) -> "OrderedDict[str, tuple]":
    d, dff, dk = arch.d_model, arch.d_ff, arch.d_head
    c0, c1 = arch.sub_channels
    (k0t, k0f), (k1t, k1f) = arch.sub_kernels
    p: "OrderedDict[str, tuple]" = OrderedDict()
    pe = "encoder.pre_encode."
    p[pe + "pre_norm.weight"] = (arch.n_mels,)
    for i, (cin, cout, kt, kf) in enumerate(((1, c0, k0t, k0f), (c0, c1, k1t, k1f))):
        p[pe + f"conv.{i}.0.weight"] = (cout, cin, kt, kf)
        p[pe + f"conv.{i}.0.bias"] = (cout,)
        for s in ("weight", "bias", "running_mean", "running_var"):
            p[pe + f"conv.{i}.1.{s}"] = (cout,)
    p[pe + "out.weight"] = (d, arch.sub_out)
    p[pe + "out_norm.weight"] = (d,)
    for l in range(arch.n_layers):
        L = f"encoder.layers.{l}."
        for n in ("norm_feed_forward1", "norm_conv", "norm_self_att", "norm_feed_forward2", "norm_out"):
            p[L + n + ".weight"] = (d,)
        for ff in ("feed_forward1", "feed_forward2"):
            p[L + ff + ".linear1.weight"] = (dff, d)
            p[L + ff + ".linear1.bias"] = (dff,)
            p[L + ff + ".linear2.weight"] = (d, dff)
            p[L + ff + ".linear2.bias"] = (d,)
            p[L + ff + ".linearv.weight"] = (dff, d)
            p[L + ff + ".linearv.bias"] = (dff,)
        p[L + "conv.pointwise_conv1.weight"] = (2 * d, d, 1)
        p[L + "conv.pointwise_conv1.bias"] = (2 * d,)
        p[L + "conv.depthwise_conv.conv.weight"] = (d, 1, arch.conv_kernel)
        p[L + "conv.depthwise_conv.conv.bias"] = (d,)
        for s in ("weight", "bias", "running_mean", "running_var"):
            p[L + "conv.batch_norm." + s] = (d,)
        p[L + "conv.pointwise_conv2.weight"] = (d, d, 1)
        p[L + "conv.pointwise_conv2.bias"] = (d,)
        att = L + "self_attn."
        names = ["linear_v", "linear_out"]
        if arch.recompute_scores[l]:
            names += ["linear_q", "linear_k"]
        for n in names:
            p[att + n + ".weight"] = (d, d)
            p[att + n + ".bias"] = (d,)
        if arch.recompute_scores[l]:
            for n in ("q_ln", "k_ln"):
                p[att + n + ".weight"] = (dk,)
                p[att + n + ".bias"] = (dk,)
    r = "encoder.temportal_reduction."  # (sic) reference: tone/nn/modules/conformer.py:113
    p[r + "conv.weight"] = (4 * d, 1, arch.reduction_kernel)
    p[r + "conv.bias"] = (4 * d,)
    p[r + "conv_pw.weight"] = (d, 4 * d, 1)
    p[r + "conv_pw.bias"] = (d,)
    p["decoder.decoder_layers.0.weight"] = (arch.n_classes, d, 1)
    p["decoder.decoder_layers.0.bias"] = (arch.n_classes,)
    return p


def n_params(arch: ToneArch = DEFAULT_ARCH) -> int:
    """Trainable parameter count (BN running stats are buffers and are not counted)."""
    n = 0
    for k, s in param_shapes(arch).items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            continue
        n += int(np.prod(s))
    return n


def _fan_in(shape) -> int:
    return int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])


def init_weights(seed: int = 0, arch: ToneArch = DEFAULT_ARCH, perturb: float = 0.1,
                 decoder_gain: float = 4.0) -> "OrderedDict[str, np.ndarray]":
    """Seeded synthetic fp32 weights in state_dict naming.

    Matrices/filters: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch's default Linear/Conv range);
    gains 1 + perturb*N(0,1); biases perturb*N(0,1)-scaled; BN running_var U(0.5, 1.5).
    ``decoder_gain`` scales the decoder matrix so that frame posteriors are not near-uniform.
    """
    rng = np.random.default_rng(seed)
    w: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in param_shapes(arch).items():
        leaf = name.rsplit(".", 1)[1]
        is_norm = (".norm_" in name or "pre_norm" in name or "out_norm" in name or "_ln." in name
                   or ".batch_norm." in name or (".conv." in name and name.split(".")[-2] == "1"))
        if leaf == "running_var":
            a = rng.uniform(0.5, 1.5, shape)
        elif leaf == "running_mean":
            a = perturb * rng.standard_normal(shape)
        elif is_norm and leaf == "weight":
            a = 1.0 + perturb * rng.standard_normal(shape)
        elif is_norm and leaf == "bias":
            a = perturb * rng.standard_normal(shape)
        elif leaf == "weight":
            b = 1.0 / np.sqrt(_fan_in(shape))
            a = rng.uniform(-b, b, shape)
            if name.startswith("decoder."):
                a = a * decoder_gain
        else:  # bias of a Linear / Conv
            src = name[: -len("bias")] + "weight"
            b = 1.0 / np.sqrt(_fan_in(param_shapes(arch)[src]))
            a = rng.uniform(-b, b, shape)
        w[name] = np.ascontiguousarray(a, dtype=np.float32)
    return w


def from_state_dict(sd, arch: ToneArch = DEFAULT_ARCH) -> "OrderedDict[str, np.ndarray]":
    """Pick our parameters out of a torch/numpy state_dict (accepts the ``tone.`` prefix)."""
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in param_shapes(arch).items():
        for key in (name, "tone." + name):
            if key in sd:
                v = sd[key]
                v = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
                if tuple(v.shape) != tuple(shape):
                    raise ValueError(f"{key}: shape {tuple(v.shape)} != expected {tuple(shape)}")
                out[name] = np.ascontiguousarray(v, dtype=np.float32)
                break
        else:
            raise KeyError(f"missing parameter {name}")
    return out


def digest(weights) -> str:
    h = hashlib.sha256()
    for k in sorted(weights):
        h.update(k.encode())
        h.update(np.ascontiguousarray(weights[k], dtype=np.float32).tobytes())
    return h.hexdigest()


# --------------------------------------------------------------------------------------
# Frontend constants (reference: tone/nn/modules/feats.py:60-93).  The reference builds
# them with torch.hann_window / torch.fft / torchaudio.melscale_fbanks at module init; they
# are not in checkpoints.  Here they are restated in numpy (float64 -> float32).
# --------------------------------------------------------------------------------------
def hann_symmetric(n: int) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / (n - 1))


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(arch: ToneArch = DEFAULT_ARCH) -> np.ndarray:
    """(n_mels, n_bins) slaney-scale, slaney-normalised triangles (feats.py:82-93)."""
    n_freqs = arch.n_bins
    all_freqs = np.linspace(0.0, arch.sample_rate // 2, n_freqs)
    m_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(arch.sample_rate / 2.0), arch.n_mels + 2)
    f_pts = _mel_to_hz_slaney(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (f_pts[2:arch.n_mels + 2] - f_pts[:arch.n_mels])
    fb = fb * enorm[None, :]
    return np.ascontiguousarray(fb.T, dtype=np.float32)


def dft_basis(arch: ToneArch = DEFAULT_ARCH) -> np.ndarray:
    """(2*n_bins, win) matrix = [cos; -sin] DFT rows x symmetric Hann x Kaldi pre-emphasis,
    so that spectrum = basis @ frame (feats.py:66-80)."""
    n = arch.n_fft
    j = np.arange(n, dtype=np.float64)
    k = np.arange(arch.n_bins, dtype=np.float64)
    ang = 2.0 * np.pi * np.outer(k, j) / n
    basis = np.concatenate([np.cos(ang), -np.sin(ang)], axis=0)  # (162, 160), fft sign convention
    basis = basis * hann_symmetric(arch.win_length)[None, :]
    a = arch.preemphasis
    if a != 0:
        # y_0 = (1-a) x_0 ; y_j = x_j - a x_{j-1}  folded into the basis columns
        out = basis.copy()
        out[:, :-1] -= a * basis[:, 1:]
        out[:, 0] -= a * basis[:, 0]
        basis = out
    return np.ascontiguousarray(basis, dtype=np.float32)
