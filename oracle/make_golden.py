"""TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  ``python -m oracle.make_golden``

Each fixture stores only *outputs* of the reference modules (tokens, attention weights, logits,
encoder outputs) plus the recipe (shape name, attention, seeds, gains); inputs and weights are
re-derived from ``oracle.synth`` (numpy PCG64, deterministic), so fixtures stay small.
The reference's beam search is only well defined for batch size 1, so beam
goldens are produced by one reference call per video, as predictor.py:102 does.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

from . import ref_shim, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> recipe.  START=1, END=2 (vocabulary.py:36-37).
CASES = {}
for _att in synth.ATTENTION_TYPES:
    CASES[f"tiny_{_att}"] = dict(shape="tiny", attention=_att, wseed=3, fseed=5, B=4, S=12, K=5,
                                 logit_gain=4.0, end_bias=0.0, feat_kind="randn")
    # END made likely so videos stop at staggered steps
    CASES[f"tiny_{_att}_end"] = dict(shape="tiny", attention=_att, wseed=7, fseed=9, B=6, S=12, K=3,
                                     logit_gain=4.0, end_bias="auto", feat_kind="ragged")
CASES["small_bahdanau"] = dict(shape="small", attention="bahdanau", wseed=11, fseed=12, B=3, S=10, K=5,
                               logit_gain=6.0, end_bias=0.0, feat_kind="randn")
CASES["msvd_bahdanau"] = dict(shape="msvd", attention="bahdanau", wseed=0, fseed=1, B=4, S=20, K=5,
                              logit_gain=1.0, end_bias=0.0, feat_kind="randn")
CASES["msvd_bahdanau_gain"] = dict(shape="msvd", attention="bahdanau", wseed=0, fseed=1, B=4, S=20, K=5,
                                   logit_gain=8.0, end_bias="auto", feat_kind="ragged")
CASES["c4_multihead"] = dict(shape="c4", attention="multihead", wseed=21, fseed=22, B=2, S=20, K=3,
                             logit_gain=8.0, end_bias=0.0, feat_kind="randn")

START, END = 1, 2


def auto_end_bias(rc):
    """Pick an END bias that makes the videos of the batch stop at staggered steps: the median of the
    per-video minimum (top logit - END logit) over the END-free greedy run, kept >= 2e-3 away from
    every observed margin so the choice is not a near-tie."""
    from .caption_oracle import CaptionOracle
    rc0 = dict(rc, end_bias=0.0)
    cfg, V, sd, feats = build_inputs(rc0)
    lg = CaptionOracle(sd).greedy(feats, START, END, max_length=rc["S"], return_logits=True)["logits"]
    m = (lg.max(-1).values - lg[..., END]).numpy()
    mins = np.sort(m.min(axis=1))
    bias = float(0.5 * (mins[len(mins) // 2 - 1] + mins[len(mins) // 2]))
    while np.abs(m - bias).min() < 2e-3:
        bias += 1e-3
    return round(bias, 6)


def build_inputs(rc):
    cfg = synth.make_config(rc["shape"])
    V = cfg.model.vocab_size
    sd = synth.make_state_dict(cfg, V, rc["attention"], seed=rc["wseed"], logit_gain=rc["logit_gain"],
                               end_token_id=END, end_bias=rc["end_bias"])
    feats = synth.make_features(rc["B"], cfg.model.video_sequence_length, cfg.model.cnn_feature_dim,
                                seed=rc["fseed"], kind=rc["feat_kind"])
    return cfg, V, sd, feats


def run_reference(rc):
    cfg, V, sd, feats = build_inputs(rc)
    model = ref_shim.build_reference_model(cfg, V, rc["attention"], state_dict=sd)
    x = torch.from_numpy(feats)
    out = {}
    with torch.no_grad():
        enc_out, final = model.encoder(x)
        out["enc_out"] = enc_out.numpy()
        out["enc_final"] = final.numpy()
        g = model.generate(x, START, END, max_length=rc["S"], method="greedy")
        out["greedy_tokens"] = g["generated_tokens"].numpy()
        out["greedy_attention"] = g["attention_weights"].numpy()
        # teacher-forced logits on the reference's own greedy tokens (the bf16 parity harness)
        inp = torch.cat([torch.full((rc["B"], 1), START, dtype=torch.long), g["generated_tokens"][:, :-1]], dim=1)
        f = model(x, inp, g["generated_tokens"])
        lg = f["logits"].numpy()
        out["tf_input_tokens"] = inp.numpy()
        if lg.size <= 400_000:
            out["tf_logits"] = lg
        else:  # keep fixtures small: first 512 vocab columns + per-row top-8
            out["tf_logits_head"] = lg[..., :512].copy()
            top = np.argsort(-lg, axis=-1)[..., :8]
            out["tf_top8_idx"] = top.astype(np.int32)
            out["tf_top8_val"] = np.take_along_axis(lg, top, axis=-1)
        srt = np.sort(lg, axis=-1)
        out["tf_margin"] = (srt[..., -1] - srt[..., -2]).astype(np.float32)      # top1-top2 audit
        out["tf_logsumexp"] = torch.logsumexp(f["logits"], dim=-1).numpy()
        rows = []
        for b in range(rc["B"]):                                                 # B=1 calls, predictor.py:102
            t = model.generate(x[b:b + 1], START, END, max_length=rc["S"], method="beam",
                               beam_size=rc["K"], length_penalty=1.0)["generated_tokens"][0]
            rows.append(t.numpy())
        L = max(len(r) for r in rows)
        bt = np.full((rc["B"], L), START, dtype=np.int64)
        bl = np.zeros(rc["B"], dtype=np.int64)
        for b, r in enumerate(rows):
            bt[b, : len(r)] = r
            bl[b] = len(r)
        out["beam_tokens"] = bt
        out["beam_lengths"] = bl
    return out


def main(argv=None):
    names = (argv or sys.argv[1:]) or list(CASES)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name in names:
        rc = dict(CASES[name])
        if rc["end_bias"] == "auto":
            rc["end_bias"] = auto_end_bias(rc)
        out = run_reference(rc)
        out["recipe"] = np.frombuffer(json.dumps(rc).encode(), dtype=np.uint8)
        out["torch_version"] = np.frombuffer(torch.__version__.encode(), dtype=np.uint8)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: greedy[0]={out['greedy_tokens'][0][:8].tolist()} beam_len={out['beam_lengths'].tolist()} "
              f"min margin={out['tf_margin'].min():.2e} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
