"""Shared helpers for the test-suite (oracle = checker only)."""
import glob
import json
import os

import numpy as np
import torch

from oracle import synth
from oracle.caption_oracle import CaptionOracle

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
START, END = 1, 2


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    g = {k: z[k] for k in z.files}
    g["recipe"] = json.loads(bytes(g["recipe"]).decode())
    return g


def diverse_golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "diverse", "*.npz")))


def load_diverse_golden(name):
    return load_golden(os.path.join("diverse", name))


def build_inputs(rc):
    cfg = synth.make_config(rc["shape"])
    V = cfg.model.vocab_size
    sd = synth.make_state_dict(cfg, V, rc["attention"], seed=rc["wseed"], logit_gain=rc["logit_gain"],
                               end_token_id=END, end_bias=rc["end_bias"])
    feats = synth.make_features(rc["B"], cfg.model.video_sequence_length, cfg.model.cnn_feature_dim,
                                seed=rc["fseed"], kind=rc["feat_kind"])
    return cfg, V, sd, feats


def make_oracle(sd, **kw):
    return CaptionOracle(sd, **kw)


def make_native_model(cfg, V, sd, attention, precision="fp32", device="cuda"):
    import video_captioning_b200 as vc
    m = vc.VideoCaptioningModel(cfg, V, attention_type=attention, precision=precision)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return m.to(device).eval()


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
