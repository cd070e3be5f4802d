"""Import the UNCHANGED reference package for live cross-checks (SURVEY 8c recipe (b)): stub the third-party modules
that are absent from this image, then import `tone` from the reference tree.  The tree is looked up at
$TONE_REFERENCE, /root/reference (build container) and baseline/_ref (the copy `__graft_entry__.build()` installs so
that it travels to the GPU box).  TEST / BASELINE INFRASTRUCTURE ONLY (same rule as tone_oracle.py): tests/, and the
reference arm / cpu_baseline leg of bench.py."""
import importlib.machinery
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    for p in (os.environ.get("TONE_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if p and os.path.isdir(os.path.join(p, "tone")):
            return p
    return None


def _stub(name):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return __import__(name, fromlist=["_"])
    except Exception:
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = m
        return m


def import_reference():
    """-> the reference's `tone` package (pipeline, splitter, decoder, nn importable), or None if no tree is present."""
    ref = reference_root()
    if ref is None:
        return None
    ort = _stub("onnxruntime")
    ort.InferenceSession = getattr(ort, "InferenceSession", object)
    ort.SessionOptions = getattr(ort, "SessionOptions", object)
    _stub("pyctcdecode")
    d = _stub("pyctcdecode.decoder")
    d.BeamSearchDecoderCTC = getattr(d, "BeamSearchDecoderCTC", object)
    d.build_ctcdecoder = getattr(d, "build_ctcdecoder", lambda *a, **k: None)
    hub = _stub("huggingface_hub")
    hub.hf_hub_download = getattr(hub, "hf_hub_download", lambda *a, **k: None)
    mod = sys.modules.get("tone")
    if mod is not None and getattr(mod, "__file__", "") and os.path.abspath(mod.__file__).startswith(os.path.abspath(ref)):
        return mod
    sys.modules.pop("tone", None)
    sys.path.insert(0, ref)
    try:
        import tone  # noqa: F401
        import tone.decoder  # noqa: F401
        import tone.logprob_splitter  # noqa: F401
        import tone.pipeline  # noqa: F401
    finally:
        sys.path.remove(ref)
    return sys.modules["tone"]


def build_reference_model(weights, skip_preprocessor: bool = False):
    """The reference's own torch model (tone.nn.model.Tone, the graph tone/scripts/export.py:411-431 traces into
    model.onnx) with OUR seeded weights loaded (state_dict names are the reference's).  None if no tree is reachable."""
    import torch
    tone = import_reference()
    if tone is None:
        return None
    from tone.nn.model import Tone  # type: ignore
    from tone.training.model_wrapper import ToneConfig  # type: ignore
    cfg = ToneConfig()
    kw = {"skip_preprocessor": True} if skip_preprocessor else {}
    model = Tone(cfg.feature_extraction_params, cfg.encoder_params, cfg.decoder_params, **kw).eval()
    sd = {k: torch.from_numpy(__import__("numpy").asarray(v)) for k, v in weights.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if not m.endswith("num_batches_tracked")]
    assert not missing and not unexpected, (missing, unexpected)
    return model


class ReferenceStreamingModel:
    """The reference torch model behind the reference's model interface (tone/onnx_wrapper.py:84-123):
    forward(audio_chunk int32 (B,2400,1), state) -> [logprobs fp32 (B,T,35), state].  State = the 7-tuple of
    Tone.get_initial_state (opaque to the pipeline).  fp32 on the host cores - the stand-in for the ORT CPU session,
    which cannot be installed offline."""
    SAMPLE_RATE, MEAN_TIME_BIAS, AUDIO_CHUNK_SAMPLES, FRAME_SIZE = 8000, 0.33, 2400, 0.03

    def __init__(self, weights):
        self.model = build_reference_model(weights)
        if self.model is None:
            raise RuntimeError("no reference tree reachable (baseline/_ref or /root/reference)")

    def forward(self, audio_chunk, state=None):
        import torch
        B = audio_chunk.shape[0]
        if state is None:
            state = self.model.get_initial_state(batch_size=B, dtype=torch.float32, len_dtype=torch.int64, device="cpu")
        with torch.no_grad():
            res = self.model.forward_for_export(torch.from_numpy(audio_chunk), None, *state)
        return [res[0].float().numpy(), tuple(res[1:])]
