"""Frozen architecture description of the streaming acoustic model.

The numbers are the `configs/streaming_acoustic` architecture, i.e. the defaults of
the reference's ``ToneConfig`` (reference: tone/training/model_wrapper.py:27-115).
They decide every shape on the hot path; the C side holds the same numbers as compile-time constants
(t-one_b200/csrc/kernels.cuh, engine.cu) and checks the shapes of the loaded weights against them.
"""
from __future__ import annotations

from dataclasses import dataclass, field

LABELS = "абвгдеёжзийклмнопрстуфхцчшщъыьэюя "  # reference: tone/decoder.py:23


@dataclass(frozen=True)
class ToneArch:
    # frontend (reference: tone/nn/modules/feats.py:34-64)
    sample_rate: int = 8000
    win_length: int = 160
    hop_length: int = 80
    n_fft: int = 160
    n_mels: int = 64
    preemphasis: float = 0.97
    # encoder (reference: tone/nn/modules/conformer.py:66-147)
    n_layers: int = 16
    d_model: int = 384
    d_ff: int = 1536
    n_heads: int = 8
    rope_dim: int = 32
    conv_kernel: int = 31
    sub_channels: tuple = (32, 64)
    sub_kernels: tuple = ((11, 21), (11, 11))
    sub_strides: tuple = ((1, 1), (3, 1))
    mhsa_stateless_layers: int = 14
    mhsa_state_size: int = 30
    recompute_scores: tuple = (
        True, False, False, False, False, False, False,
        True, False, False, False, False, False, False,
        True, True,
    )
    reduction_position: int = 6
    upsample_position: int = 14
    reduction_factor: int = 2
    reduction_kernel: int = 3
    # decoder
    n_classes: int = 35  # 34 labels + blank (reference: tone/nn/modules/conformer.py:331-332)

    # ---- derived sizes -------------------------------------------------
    @property
    def d_head(self) -> int:
        return self.d_model // self.n_heads

    @property
    def n_bins(self) -> int:
        return self.n_fft // 2 + 1

    @property
    def pre_state(self) -> int:  # samples carried by the frontend (feats.py:57)
        return self.n_fft - self.hop_length

    @property
    def sub1_rows(self) -> int:  # time rows cached in front of conv0 (conformer_blocks.py:517)
        return self.sub_kernels[0][0] - self.sub_strides[0][0]

    @property
    def sub2_rows(self) -> int:
        return self.sub_kernels[1][0] - self.sub_strides[1][0]

    @property
    def sub_f1(self) -> int:  # feature width after conv0: 64-21+1
        return self.n_mels - self.sub_kernels[0][1] + 1

    @property
    def sub_f2(self) -> int:  # after conv1: 44-11+1
        return self.sub_f1 - self.sub_kernels[1][1] + 1

    @property
    def sub_out(self) -> int:  # 64*34 = 2176 inputs of pre_encode.out
        return self.sub_channels[1] * self.sub_f2

    @property
    def conv_state(self) -> int:
        return self.conv_kernel - 1

    @property
    def red_state(self) -> int:
        return self.reduction_kernel - self.reduction_factor

    @property
    def n_mhsa_stateful(self) -> int:
        return self.n_layers - self.mhsa_stateless_layers

    def frames(self, chunk_samples: int) -> int:
        """mel frames produced by one chunk (feats.py:98, stride = hop)."""
        return chunk_samples // self.hop_length

    def t_out(self, chunk_samples: int) -> int:
        """30 ms output frames of one chunk: conv1 on [8 cached | F new] rows, stride 3."""
        f = self.frames(chunk_samples)
        return (f + self.sub2_rows - self.sub_kernels[1][0]) // self.sub_strides[1][0] + 1

    def t_red(self, chunk_samples: int) -> int:
        """frames inside the time-reduced block (layers 7..14)."""
        t = self.t_out(chunk_samples)
        return (t + self.red_state - self.reduction_kernel) // self.reduction_factor + 1

    def layer_reduced(self, layer: int) -> bool:
        return self.reduction_position < layer <= self.upsample_position

    def state_layout(self):
        """(name, shape) of the seven state tensors per stream, in get_initial_state order
        (reference: tone/nn/model.py:259-267).  Their sizes sum to STATE_SIZE."""
        return (
            ("preproc", (self.pre_state,)),
            ("mhsa", (self.n_mhsa_stateful, self.mhsa_state_size, self.d_model)),
            ("conv", (self.n_layers, self.d_model, self.conv_state)),
            ("mhsa_len", (1,)),
            ("sub1", (1, self.sub1_rows, self.n_mels)),
            ("sub2", (self.sub_channels[0], self.sub2_rows, self.sub_f1)),
            ("reduction", (self.d_model, self.red_state)),
        )

    @property
    def state_size(self) -> int:
        n = 0
        for _, shp in self.state_layout():
            k = 1
            for s in shp:
                k *= s
            n += k
        return n


DEFAULT_ARCH = ToneArch()
assert DEFAULT_ARCH.state_size == 219729  # reference: tone/onnx_wrapper.py:34
assert DEFAULT_ARCH.t_out(2400) == 10 and DEFAULT_ARCH.t_out(3200) == 13
assert DEFAULT_ARCH.t_red(2400) == 5 and DEFAULT_ARCH.t_red(3200) == 6
