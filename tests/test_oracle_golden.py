"""CPU: the oracle restatement against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  This is what pins the oracle (the reference has no tests of its own)."""
import numpy as np
import pytest
import torch

from _util import END, START, build_inputs, golden_names, load_golden, make_oracle

NAMES = golden_names()
FAST = [n for n in NAMES if not n.startswith(("msvd", "c4"))]


def test_fixtures_present():
    assert len(NAMES) >= 10


@pytest.mark.parametrize("name", FAST + ["msvd_bahdanau_gain"])
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    rc = g["recipe"]
    cfg, V, sd, feats = build_inputs(rc)
    o = make_oracle(sd)
    enc, fin = o.encode(torch.from_numpy(feats))
    np.testing.assert_allclose(enc.numpy(), g["enc_out"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(fin.numpy(), g["enc_final"], rtol=0, atol=2e-6)
    gr = o.greedy(feats, START, END, max_length=rc["S"])
    assert np.array_equal(gr["generated_tokens"].numpy(), g["greedy_tokens"])
    np.testing.assert_allclose(gr["attention_weights"].numpy(), g["greedy_attention"], rtol=0, atol=2e-6)
    tf = o.forward_teacher(feats, g["tf_input_tokens"])
    lg = tf["logits"].numpy()
    if "tf_logits" in g:
        np.testing.assert_allclose(lg, g["tf_logits"], rtol=0, atol=5e-6)
    else:
        np.testing.assert_allclose(lg[..., :512], g["tf_logits_head"], rtol=0, atol=5e-6)
        np.testing.assert_allclose(np.take_along_axis(lg, g["tf_top8_idx"].astype(np.int64), -1), g["tf_top8_val"],
                                   rtol=0, atol=5e-6)
    bm = o.beam(feats, START, END, max_length=rc["S"], beam_size=rc["K"])
    assert np.array_equal(bm["generated_tokens"].numpy(), g["beam_tokens"])
    assert np.array_equal(bm["lengths"].numpy(), g["beam_lengths"])


@pytest.mark.parametrize("name", FAST)
def test_beam_equals_greedy_known_answer(name):
    """reference beam-K == [START] + greedy tokens, truncated after the first END."""
    g = load_golden(name)
    for b in range(g["recipe"]["B"]):
        row = g["greedy_tokens"][b].tolist()
        if END in row:
            row = row[: row.index(END) + 1]
        # the batched greedy run may have been cut short by the all-rows-END rule
        n = min(len(row) + 1, int(g["beam_lengths"][b]))
        assert g["beam_tokens"][b, :n].tolist() == ([START] + row)[:n]


def test_fp64_oracle_close_to_fp32():
    g = load_golden("tiny_bahdanau")
    cfg, V, sd, feats = build_inputs(g["recipe"])
    o64 = make_oracle(sd, dtype=torch.float64)
    tf = o64.forward_teacher(torch.from_numpy(feats).double(), g["tf_input_tokens"])
    np.testing.assert_allclose(tf["logits"].numpy(), g["tf_logits"], rtol=0, atol=2e-5)
