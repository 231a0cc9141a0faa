"""CaptionDecoder: embedding + attention + L-layer LSTM step + context/vocab projections.

Drop-in parameter layout for the reference's ``src/models/decoder.py`` (keys ``embedding.weight``,
``attention.*``, ``lstm.*``, ``context_projection.*``, ``output_projection.*``).  Initialisers follow
decoder.py:64-79.  The decode loop itself (decoder.py:223-289) runs on the device
(``csrc/capi.cu:run_decode``); this module only holds the parameters.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .attention import _Projection, create_attention_mechanism
from .encoder import _LSTMParams


class _Embedding(nn.Module):
    def __init__(self, num_embeddings: int, embedding_dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(num_embeddings, embedding_dim).uniform_(-0.1, 0.1))


class CaptionDecoder(nn.Module):
    def __init__(self, config, vocabulary_size: int, attention_type: str = "bahdanau", num_heads: int = 8):
        super().__init__()
        m = config.model
        if not getattr(m, "use_attention", True):
            raise ValueError("the native path implements the attention decoder (config.model.use_attention=True)")
        if m.encoder_hidden_dim != m.decoder_hidden_dim:
            # decoder.py:97-99 creates a fresh random nn.Linear on every call in this case: not reproducible
            raise ValueError("encoder_hidden_dim != decoder_hidden_dim is not supported (irreproducible in the reference)")
        self.config = config
        self.vocabulary_size = vocabulary_size
        self.embedding_dim = m.embedding_dim
        self.hidden_dim = m.decoder_hidden_dim
        self.encoder_dim = m.encoder_hidden_dim
        self.num_layers = m.decoder_num_layers
        self.use_attention = True
        self.embedding = _Embedding(vocabulary_size, self.embedding_dim)
        if attention_type == "multihead":
            from .attention import MultiHeadAttention
            self.attention = MultiHeadAttention(config, num_heads)
        else:
            self.attention = create_attention_mechanism(config, attention_type)
        self.lstm = _LSTMParams(self.embedding_dim + self.encoder_dim, self.hidden_dim, self.num_layers, bidirectional=False)
        self.context_projection = _Projection(self.encoder_dim + self.hidden_dim + self.embedding_dim, self.hidden_dim)
        self.output_projection = _Projection(self.hidden_dim, vocabulary_size)
        self._init_weights()

    def _init_weights(self) -> None:
        for name, p in self.lstm.named_parameters():
            if "weight" in name:
                nn.init.orthogonal_(p)
            else:
                nn.init.zeros_(p)
        nn.init.xavier_uniform_(self.output_projection.weight)
        nn.init.zeros_(self.output_projection.bias)
        nn.init.xavier_uniform_(self.context_projection.weight)
        nn.init.zeros_(self.context_projection.bias)
