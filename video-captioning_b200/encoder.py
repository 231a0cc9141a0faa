"""VideoEncoder: feature projection -> bidirectional LSTM -> shared output projection.

Drop-in parameter layout for the reference's ``src/models/encoder.py:10-98`` (keys
``feature_projection.*``, ``lstm.weight_ih_l{k}[_reverse]`` ..., ``output_projection.*``).  The module is a
parameter container; ``forward`` runs the native encoder (all-timestep tensor-core / FFMA input
projections + fused recurrent GEMM/LSTM-cell steps, ``csrc/capi.cu:run_encoder``).

The CNN feature extractors of encoder.py:101-226 are out of scope: features are precomputed inputs
(SURVEY.md section 2.1 row 6).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _native
from .attention import _Projection


class _LSTMParams(nn.Module):
    """Parameters of an ``nn.LSTM`` under the same names (weight_ih_l0, bias_hh_l1_reverse, ...)."""

    def __init__(self, input_size: int, hidden_size: int, num_layers: int, bidirectional: bool):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.num_layers, self.bidirectional = num_layers, bidirectional
        dirs = 2 if bidirectional else 1
        bound = 1.0 / (hidden_size ** 0.5)
        for layer in range(num_layers):
            inp = input_size if layer == 0 else hidden_size * dirs
            for d in range(dirs):
                sfx = f"_l{layer}" + ("_reverse" if d else "")
                for name, shape in (("weight_ih", (4 * hidden_size, inp)), ("weight_hh", (4 * hidden_size, hidden_size)),
                                    ("bias_ih", (4 * hidden_size,)), ("bias_hh", (4 * hidden_size,))):
                    p = nn.Parameter(torch.empty(*shape).uniform_(-bound, bound))
                    self.register_parameter(name + sfx, p)


class VideoEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.feature_dim = config.model.cnn_feature_dim
        self.hidden_dim = config.model.encoder_hidden_dim
        self.num_layers = config.model.encoder_num_layers
        self.feature_projection = _Projection(self.feature_dim, self.hidden_dim)
        self.lstm = _LSTMParams(self.hidden_dim, self.hidden_dim, self.num_layers, bidirectional=True)
        self.output_projection = _Projection(2 * self.hidden_dim, self.hidden_dim)
        self._owner = None   # set by VideoCaptioningModel so the shared native handle is used

    def forward(self, video_features: torch.Tensor, video_mask=None):
        """-> (encoded_features [B,T,H], final_hidden_state [B,H])  (encoder.py:52-98)."""
        _native.require_cuda(video_features, "video_features")
        if self._owner is None:
            raise RuntimeError("VideoEncoder.forward needs the encoder to belong to a VideoCaptioningModel "
                               "(the native handle is built from the whole state_dict)")
        return self._owner()._handle().encoder_forward(video_features, video_mask)
