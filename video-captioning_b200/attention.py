"""Attention parameter containers (drop-in names for the reference's ``src/models/attention.py``).

The modules only *hold* parameters under the reference's ``state_dict`` keys; the arithmetic runs in
the fused CUDA attention step (``csrc/attention.cuh``).  ``forward`` evaluates one attention step
through the native library (there is no PyTorch implementation to fall back to).

Reference: BahdanauAttention attention.py:9-73, LuongAttention :76-187, MultiHeadAttention :190-275,
factory :278-296.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _native


def _init_linear(weight: torch.Tensor, bias, fan_in: int) -> None:
    bound = 1.0 / (fan_in ** 0.5)
    with torch.no_grad():
        weight.uniform_(-bound, bound)
        if bias is not None:
            bias.uniform_(-bound, bound)


class _Projection(nn.Module):
    """weight [out,in] (+ bias [out]) with nn.Linear's parameter names; never called as a layer."""

    def __init__(self, fan_in: int, fan_out: int, bias: bool = True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(fan_out, fan_in))
        if bias:
            self.bias = nn.Parameter(torch.empty(fan_out))
        else:
            self.register_parameter("bias", None)
        _init_linear(self.weight, self.bias, fan_in)


class _NativeAttention(nn.Module):
    attention_kind = "bahdanau"
    num_heads = 1

    def __init__(self, config):
        super().__init__()
        self.encoder_dim = config.model.encoder_hidden_dim
        self.decoder_dim = config.model.decoder_hidden_dim
        self.attention_dim = config.model.attention_dim
        self._standalone = None

    def forward(self, encoder_outputs, decoder_hidden, encoder_mask=None):
        """(context [R,H], weights [R,T]) for R == encoder_outputs.shape[0] rows, like the reference call."""
        _native.require_cuda(encoder_outputs, "encoder_outputs")
        from .video_captioning_model import standalone_attention_handle
        h = standalone_attention_handle(self)
        return h.attention_step(encoder_outputs, decoder_hidden, encoder_mask, 1)


class BahdanauAttention(_NativeAttention):
    """score_t = v . tanh(W_e enc_t + b_e + W_d h + b_d) + b_v   (attention.py:52-57)."""
    attention_kind = "bahdanau"

    def __init__(self, config):
        super().__init__(config)
        self.encoder_projection = _Projection(self.encoder_dim, self.attention_dim)
        self.decoder_projection = _Projection(self.decoder_dim, self.attention_dim)
        self.attention_linear = _Projection(self.attention_dim, 1)


class LuongAttention(_NativeAttention):
    """dot / general / concat scoring (attention.py:118-146)."""

    def __init__(self, config, score_function: str = "general"):
        super().__init__(config)
        if score_function not in ("dot", "general", "concat"):
            raise ValueError(f"Unknown score function: {score_function}")
        self.score_function = score_function
        self.attention_kind = f"luong_{score_function}"
        if score_function == "dot" and self.decoder_dim != self.encoder_dim:
            raise ValueError("For dot attention, decoder and encoder dimensions must match")
        if score_function == "general":
            self.linear_in = _Projection(self.decoder_dim, self.encoder_dim, bias=False)
        elif score_function == "concat":
            self.linear_query = _Projection(self.decoder_dim, self.attention_dim)
            self.linear_context = _Projection(self.encoder_dim, self.attention_dim)
            self.linear_v = _Projection(self.attention_dim, 1, bias=False)


class MultiHeadAttention(_NativeAttention):
    """n-head scaled dot-product attention with a single query (attention.py:237-275)."""
    attention_kind = "multihead"

    def __init__(self, config, num_heads: int = 8):
        super().__init__(config)
        if self.encoder_dim % num_heads != 0:
            raise AssertionError("encoder_dim must be divisible by num_heads")
        self.num_heads = num_heads
        self.head_dim = self.encoder_dim // num_heads
        self.query_linear = _Projection(self.decoder_dim, self.encoder_dim)
        self.key_linear = _Projection(self.encoder_dim, self.encoder_dim)
        self.value_linear = _Projection(self.encoder_dim, self.encoder_dim)
        self.output_linear = _Projection(self.encoder_dim, self.encoder_dim)


def create_attention_mechanism(config, attention_type: str = "bahdanau") -> nn.Module:
    """Same factory contract as attention.py:278-296 ('luong' yields the general score)."""
    kind = attention_type.lower()
    if kind == "bahdanau":
        return BahdanauAttention(config)
    if kind == "luong":
        return LuongAttention(config)
    if kind in ("luong_dot", "luong_general", "luong_concat"):
        return LuongAttention(config, kind.split("_", 1)[1])
    if kind == "multihead":
        return MultiHeadAttention(config)
    raise ValueError(f"Unsupported attention type: {attention_type}")
