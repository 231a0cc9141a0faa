"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic configs, state_dicts and features.

The reference ships no checkpoints, datasets or fixtures, so every parity test, golden vector and
benchmark uses random-init weights of the reference architecture.  They are drawn from numpy's
PCG64 stream (stable across numpy/torch versions, unlike torch's init RNG) with the same
*distributions* as the reference's initialisers:

  nn.Linear / encoder nn.LSTM defaults  U(+-1/sqrt(fan_in))   (encoder.py:29-47, attention.py:26-28)
  decoder embedding                      U(-0.1, 0.1)          (decoder.py:66)
  decoder LSTM (orthogonal there)        U(+-sqrt(3/rows)): same per-entry variance as an
                                         orthogonal [4H,in] matrix            (decoder.py:68-72)
  context/vocab projections (Xavier)     U(+-sqrt(6/(in+out)))                (decoder.py:74-79)

Biases that the reference zero-initialises are given small non-zero values so the bias paths of the
kernels are exercised.  Key names and shapes are exactly the reference ``state_dict`` layout
 so the same dict loads into the reference module, the oracle and the product.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

ATTENTION_TYPES = ("bahdanau", "luong_general", "luong_dot", "luong_concat", "multihead")

# Named shapes: "tiny"/"small" are the golden-fixture shapes, c1..c5 are BASELINE.json configs.
SHAPES = {
    "tiny": dict(F=256, H=128, E=128, A=128, V=1000, T=16, Le=2, Ld=2),
    "small": dict(F=512, H=256, E=192, A=128, V=2500, T=24, Le=2, Ld=2),
    "msvd": dict(F=4096, H=512, E=512, A=512, V=10000, T=80, Le=2, Ld=2),       # configs 1, 2
    "c3": dict(F=4096, H=1024, E=512, A=512, V=10000, T=80, Le=2, Ld=2),        # config 3
    "c4": dict(F=2048, H=512, E=512, A=512, V=10000, T=40, Le=2, Ld=2),         # config 4
    "c5": dict(F=4096, H=512, E=512, A=512, V=30000, T=80, Le=2, Ld=2),         # config 5
}


def make_config(shape="msvd", **over):
    """Attribute bag with the ``config.model.*`` names the hot path reads (config.py:13-31)."""
    d = dict(SHAPES[shape]) if isinstance(shape, str) else dict(shape)
    d.update(over)
    model = SimpleNamespace(
        cnn_feature_dim=d["F"], encoder_hidden_dim=d["H"], encoder_num_layers=d["Le"],
        encoder_dropout=0.3, decoder_hidden_dim=d["H"], decoder_num_layers=d["Ld"],
        decoder_dropout=0.3, vocab_size=d["V"], embedding_dim=d["E"], attention_dim=d["A"],
        use_attention=True, max_sequence_length=20, video_sequence_length=d["T"])
    data = SimpleNamespace(pad_token="<PAD>", start_token="<START>", end_token="<END>",
                           unk_token="<UNK>", vocab_threshold=5, max_vocab_size=d["V"])
    return SimpleNamespace(model=model, data=data)


def _u(rng, shape, bound):
    return ((rng.random(shape, dtype=np.float32) * 2.0 - 1.0) * np.float32(bound)).astype(np.float32)


def make_state_dict(cfg, vocab_size, attention="bahdanau", seed=0, logit_gain=1.0,
                    end_token_id=None, end_bias=0.0):
    """Reference-layout state_dict of float32 numpy arrays.

    ``logit_gain`` scales the vocab projection (wider top1-top2 margins than the nearly flat
    Xavier logits); ``end_bias`` raises ``output_projection.bias[END]``
    so videos finish at staggered steps.
    """
    m = cfg.model
    F, H, E, A, V = m.cnn_feature_dim, m.encoder_hidden_dim, m.embedding_dim, m.attention_dim, vocab_size
    Le, Ld = m.encoder_num_layers, m.decoder_num_layers
    assert m.decoder_hidden_dim == H, "encoder_hidden_dim != decoder_hidden_dim is rejected (decoder.py:97-99)"
    rng = np.random.default_rng(seed)
    sd = {}
    b = 1.0 / np.sqrt(F)
    sd["encoder.feature_projection.weight"] = _u(rng, (H, F), b)
    sd["encoder.feature_projection.bias"] = _u(rng, (H,), b)
    b = 1.0 / np.sqrt(H)
    for layer in range(Le):
        for sfx in ("", "_reverse"):
            inp = H if layer == 0 else 2 * H
            sd[f"encoder.lstm.weight_ih_l{layer}{sfx}"] = _u(rng, (4 * H, inp), b)
            sd[f"encoder.lstm.weight_hh_l{layer}{sfx}"] = _u(rng, (4 * H, H), b)
            sd[f"encoder.lstm.bias_ih_l{layer}{sfx}"] = _u(rng, (4 * H,), b)
            sd[f"encoder.lstm.bias_hh_l{layer}{sfx}"] = _u(rng, (4 * H,), b)
    b = 1.0 / np.sqrt(2 * H)
    sd["encoder.output_projection.weight"] = _u(rng, (H, 2 * H), b)
    sd["encoder.output_projection.bias"] = _u(rng, (H,), b)
    sd["decoder.embedding.weight"] = _u(rng, (V, E), 0.1)
    if attention == "bahdanau":
        sd["decoder.attention.encoder_projection.weight"] = _u(rng, (A, H), 1 / np.sqrt(H))
        sd["decoder.attention.encoder_projection.bias"] = _u(rng, (A,), 1 / np.sqrt(H))
        sd["decoder.attention.decoder_projection.weight"] = _u(rng, (A, H), 1 / np.sqrt(H))
        sd["decoder.attention.decoder_projection.bias"] = _u(rng, (A,), 1 / np.sqrt(H))
        sd["decoder.attention.attention_linear.weight"] = _u(rng, (1, A), 1 / np.sqrt(A))
        sd["decoder.attention.attention_linear.bias"] = _u(rng, (1,), 1 / np.sqrt(A))
    elif attention == "luong_general":
        sd["decoder.attention.linear_in.weight"] = _u(rng, (H, H), 1 / np.sqrt(H))
    elif attention == "luong_dot":
        pass
    elif attention == "luong_concat":
        sd["decoder.attention.linear_query.weight"] = _u(rng, (A, H), 1 / np.sqrt(H))
        sd["decoder.attention.linear_query.bias"] = _u(rng, (A,), 1 / np.sqrt(H))
        sd["decoder.attention.linear_context.weight"] = _u(rng, (A, H), 1 / np.sqrt(H))
        sd["decoder.attention.linear_context.bias"] = _u(rng, (A,), 1 / np.sqrt(H))
        sd["decoder.attention.linear_v.weight"] = _u(rng, (1, A), 1 / np.sqrt(A))
    elif attention == "multihead":
        for nm in ("query", "key", "value", "output"):
            sd[f"decoder.attention.{nm}_linear.weight"] = _u(rng, (H, H), 1 / np.sqrt(H))
            sd[f"decoder.attention.{nm}_linear.bias"] = _u(rng, (H,), 1 / np.sqrt(H))
    else:
        raise ValueError(attention)
    for layer in range(Ld):
        inp = (E + H) if layer == 0 else H
        sd[f"decoder.lstm.weight_ih_l{layer}"] = _u(rng, (4 * H, inp), np.sqrt(3.0 / (4 * H)))
        sd[f"decoder.lstm.weight_hh_l{layer}"] = _u(rng, (4 * H, H), np.sqrt(3.0 / (4 * H)))
        sd[f"decoder.lstm.bias_ih_l{layer}"] = _u(rng, (4 * H,), 0.05)
        sd[f"decoder.lstm.bias_hh_l{layer}"] = _u(rng, (4 * H,), 0.05)
    sd["decoder.context_projection.weight"] = _u(rng, (H, 2 * H + E), np.sqrt(6.0 / (3 * H + E)))
    sd["decoder.context_projection.bias"] = _u(rng, (H,), 0.05)
    sd["decoder.output_projection.weight"] = _u(rng, (V, H), np.sqrt(6.0 / (H + V))) * np.float32(logit_gain)
    sd["decoder.output_projection.bias"] = _u(rng, (V,), 0.05)
    if end_token_id is not None and end_bias != 0.0:
        sd["decoder.output_projection.bias"][end_token_id] += np.float32(end_bias)
    return sd


def make_features(B, T, F, seed=1, kind="randn"):
    """[B,T,F] float32.  'randn' ~ N(0,1); 'relu' = max(randn,0) like post-ReLU fc7 features;
    'ragged' = relu features with a per-video scale and a zero-padded tail of frames (short videos
    are zero-padded to T by predictor.py:312-315 and encoded as real frames), which makes the
    videos of a batch behave differently (staggered END steps)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, F), dtype=np.float32)
    if kind in ("relu", "ragged"):
        x = np.maximum(x, 0.0)
    if kind == "ragged":
        for b in range(B):
            x[b] *= np.float32(0.25 + 0.75 * ((b * 7) % 5))
            x[b, T - ((b * 3) % (T // 2)):] = 0.0
    return x
