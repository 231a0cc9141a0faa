"""CPU: the oracle restatement against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  This is what pins the oracle (the reference has no tests of its own)."""
import numpy as np
import pytest
import torch

from _util import (END, START, build_inputs, diverse_golden_names, golden_names, load_diverse_golden, load_golden,
                   make_oracle)

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


# ------------------------------------------------------------------ real ("diverse") beam search, SURVEY 8f rank 3
DIVERSE = diverse_golden_names()


@pytest.mark.parametrize("name", [n for n in DIVERSE if not n.startswith(("msvd", "c4"))])
def test_oracle_diverse_beam_matches_repaired_reference(name):
    """oracle.beam(diverse=True) against the unmodified reference loop run with scores[1:] = -inf
    (oracle/make_golden_diverse.py, ref_shim.reference_diverse_beam): same best hypothesis per video."""
    g = load_diverse_golden(name)
    rc = g["recipe"]
    cfg, V, sd, feats = build_inputs(rc)
    bm = make_oracle(sd).beam(feats, START, END, max_length=rc["S"], beam_size=rc["K"], length_penalty=rc["lp"],
                              diverse=True, num_return=rc["K"])
    assert np.array_equal(bm["lengths"].numpy(), g["beam_lengths"])
    assert np.array_equal(bm["generated_tokens"].numpy(), g["beam_tokens"])
    # n-best list: entry 0 is the returned hypothesis; completed hypotheses come sorted by normalised score
    nt, nl, ns = bm["nbest_tokens"].numpy(), bm["nbest_lengths"].numpy(), bm["nbest_scores"].numpy()
    for b in range(rc["B"]):
        n0 = int(nl[b, 0])
        assert nt[b, 0, :n0].tolist() == g["beam_tokens"][b, :n0].tolist()
        done = [j for j in range(rc["K"]) if nl[b, j] > 0 and nt[b, j, nl[b, j] - 1] == END]
        assert done == list(range(len(done))), "completed hypotheses precede live ones"
        assert all(ns[b, j] >= ns[b, j + 1] for j in range(len(done) - 1))
        rows = {tuple(nt[b, j, : nl[b, j]].tolist()) for j in range(rc["K"]) if nl[b, j] > 0}
        assert len(rows) == int((nl[b] > 0).sum()), "hypotheses are distinct"


def test_oracle_sequence_logprob_equals_beam_score():
    """The score the beam reports for its best hypothesis is the (length-normalised) sum of the log-probabilities
    of its tokens: sequence_logprob re-derives it teacher-forced."""
    g = load_diverse_golden("tiny_bahdanau_k5")
    rc = g["recipe"]
    cfg, V, sd, feats = build_inputs(rc)
    o = make_oracle(sd)
    bm = o.beam(feats, START, END, max_length=rc["S"], beam_size=rc["K"], length_penalty=rc["lp"], diverse=True)
    lp = o.sequence_logprob(feats, bm["generated_tokens"], bm["lengths"])
    toks, lens = bm["generated_tokens"], bm["lengths"]
    for b in range(rc["B"]):
        n = int(lens[b]) - 1
        ended = int(toks[b, n]) == END
        exp = float(lp[b]) / (n ** rc["lp"]) if ended else float(lp[b])
        assert abs(exp - float(bm["scores"][b])) < 1e-4 * max(1.0, abs(exp))
