"""tone-b200: the T-one batched streaming acoustic-model step, B200-native.

Host-side mirror of the reference's acoustic-model interface
(reference: tone/onnx_wrapper.py:17-123) over a C-ABI CUDA library (include/tone_b200.h).
The directory name carries a hyphen; import it as ``importlib.import_module("t-one_b200")``
or through the ``tone_b200`` alias module at the repo root.
"""
from . import arch, greedy, model, scheduler, sharding, state_formats, synth, weights  # noqa: F401
from .arch import DEFAULT_ARCH, LABELS, ToneArch  # noqa: F401
from .model import B200StreamingCTCModel, Engine, StreamSlots, load_library  # noqa: F401

__all__ = ["arch", "greedy", "model", "scheduler", "sharding", "state_formats", "synth", "weights", "DEFAULT_ARCH", "LABELS", "ToneArch",
           "B200StreamingCTCModel", "Engine", "StreamSlots", "load_library"]
