"""Generate the golden vectors that pin oracle/tone_oracle.py to the reference.

Run in the BUILD container only (needs the read-only reference tree):

    python tests/golden/make_golden.py            # writes tests/golden/step_{300,400}ms.npz

It imports the reference's own torch model (``tone.nn.model.Tone``; the exact graph that
``tone/scripts/export.py:411-431`` traces into model.onnx), bypassing ``tone/__init__.py``
(which imports onnxruntime / pyctcdecode, absent here), loads OUR seeded synthetic weights
into it (no checkpoint is cached and there is no network), and records what the reference
computes for seeded synthetic telephony-like audio.  Nothing here is copied from the
reference; the fixtures are its outputs.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("TONE_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)


def import_reference_model():
    """SURVEY.md §8c recipe (a): model only, skipping tone/__init__.py."""
    if "tone" not in sys.modules:
        pkg = types.ModuleType("tone")
        pkg.__path__ = [os.path.join(REF, "tone")]
        sys.modules["tone"] = pkg
    from tone.nn.model import Tone  # type: ignore
    from tone.training.model_wrapper import ToneConfig  # type: ignore
    return Tone, ToneConfig


def build_reference(weights):
    Tone, ToneConfig = import_reference_model()
    cfg = ToneConfig()
    model = Tone(cfg.feature_extraction_params, cfg.encoder_params, cfg.decoder_params).eval()
    sd = {k: torch.from_numpy(v) for k, v in weights.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if not m.endswith("num_batches_tracked")]
    assert not missing and not unexpected, (missing, unexpected)
    return model


def reference_stream(weights, pcm_chunks, mode="fp32"):
    """pcm_chunks: list of (B,C) int32.  Returns (list of logprobs, final 7-tuple state).
    mode 'fp32'  : fp32 states, whole graph fp32 (except the mandated fp16 waveform rounding)
    mode 'export': fp16 states under fp16 autocast, i.e. what ORT runs (export.py:411,454-455)."""
    model = build_reference(weights)   # fresh instance: the RoPE table is cached per instance
    B = pcm_chunks[0].shape[0]
    dtype = torch.float16 if mode == "export" else torch.float32
    state = model.get_initial_state(batch_size=B, dtype=dtype, len_dtype=torch.int64, device="cpu")
    outs = []
    with torch.no_grad():
        for pcm in pcm_chunks:
            x = torch.from_numpy(pcm.astype(np.int32))[:, :, None]
            if mode == "export":
                with torch.amp.autocast("cpu", dtype=torch.float16):
                    res = model.forward_for_export(x, None, *state)
                state = tuple(s.half() if s.is_floating_point() else s for s in res[1:])
            else:
                res = model.forward_for_export(x, None, *state)
                state = tuple(res[1:])
            outs.append(res[0].float().numpy())
    return outs, state


def reference_feature_stream(weights, feat_chunks):
    """The reference's feature-input mode (skip_preprocessor=True, tone/nn/model.py:151-160): feat_chunks is a list of
    (B,64,F) float arrays; `length` = F frames for every stream.  Returns the list of logprobs and the final state."""
    Tone, ToneConfig = import_reference_model()
    cfg = ToneConfig()
    model = Tone(cfg.feature_extraction_params, cfg.encoder_params, cfg.decoder_params, skip_preprocessor=True).eval()
    sd = {k: torch.from_numpy(v) for k, v in weights.items()}
    model.load_state_dict(sd, strict=False)
    # the reference casts the features to fp16 (model.py:154), so this mode only runs the way it is exported:
    # fp16 states under fp16 autocast (export.py:411,454-455)
    B = feat_chunks[0].shape[0]
    state = model.get_initial_state(batch_size=B, dtype=torch.float16, len_dtype=torch.int64, device="cpu")
    outs = []
    with torch.no_grad():
        for f in feat_chunks:
            x = torch.from_numpy(f.astype(np.float32))
            length = torch.full((B,), x.shape[2], dtype=torch.int64)
            with torch.amp.autocast("cpu", dtype=torch.float16):
                res = model.forward_for_export(x, length, *state)
            state = tuple(s.half() if s.is_floating_point() else s for s in res[1:])
            outs.append(res[0].float().numpy())
    return outs, state


def synth_pcm(B, n_samples, seed=1234):
    tone_b200 = importlib.import_module("t-one_b200")
    return tone_b200.synth.telephony_pcm(B, n_samples, seed)


def main():
    tone_b200 = importlib.import_module("t-one_b200")
    weights = tone_b200.weights.init_weights(seed=0)
    for ms, C, n_chunks in ((300, 2400, 6), (400, 3200, 5)):
        B = 3
        pcm = synth_pcm(B, C * n_chunks)
        chunks = [np.ascontiguousarray(pcm[:, i * C:(i + 1) * C]) for i in range(n_chunks)]
        outs, state = reference_stream(weights, chunks, "fp32")
        outs16, state16 = reference_stream(weights, chunks, "export")
        names = ("preproc", "mhsa", "conv", "mhsa_len", "sub1", "sub2", "reduction")
        blob = {
            "weights_digest": np.frombuffer(tone_b200.weights.digest(weights).encode(), dtype=np.uint8),
            "pcm": pcm.astype(np.int16),
            "logprobs": np.stack(outs, 0).astype(np.float32),            # (n_chunks,B,T,35)
            "logprobs_export": np.stack(outs16, 0).astype(np.float32),
        }
        # final carried state of stream 0 only (keeps the fixture small); fp16 like the wire format
        for n, s in zip(names, state):
            a = s.numpy()[:1]
            blob["state_" + n] = a.astype(np.float16) if a.dtype.kind == "f" else a.astype(np.int64)
        blob["state16_mhsa_len"] = state16[3].numpy().astype(np.int64)
        path = os.path.join(HERE, f"step_{ms}ms.npz")
        np.savez_compressed(path, **blob)
        print(path, os.path.getsize(path) // 1024, "KiB", "logprobs", blob["logprobs"].shape)


def main_features():
    """features_300ms.npz: log-mel features (made by the oracle's frontend from synthetic audio, stored fp16) and what the
    reference computes from them in feature-input mode."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tone_oracle as orc
    tone_b200 = importlib.import_module("t-one_b200")
    weights = tone_b200.weights.init_weights(seed=0)
    B, C, n_chunks = 2, 2400, 4
    pcm = synth_pcm(B, C * n_chunks, seed=4321)
    pre = torch.zeros(B, 80)
    feats = []
    for i in range(n_chunks):
        f, pre = orc.frontend(torch.from_numpy(pcm[:, i * C:(i + 1) * C].astype(np.int32)), pre)
        feats.append(f.transpose(1, 2).numpy().astype(np.float16))          # (B,64,F)
    outs, state = reference_feature_stream(weights, feats)
    blob = {
        "weights_digest": np.frombuffer(tone_b200.weights.digest(weights).encode(), dtype=np.uint8),
        "feats": np.stack(feats, 0),                                          # (n,B,64,F) fp16
        "logprobs": np.stack(outs, 0).astype(np.float32),
        "state_mhsa_len": state[3].numpy().astype(np.int64),
    }
    path = os.path.join(HERE, "features_300ms.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path) // 1024, "KiB", "logprobs", blob["logprobs"].shape)


if __name__ == "__main__":
    if "--features" in sys.argv:
        main_features()
    else:
        main()
        main_features()
