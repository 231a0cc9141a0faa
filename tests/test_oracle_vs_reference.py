"""CPU, build container only: live comparison oracle <-> unmodified reference (skipped where
/root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ref_shim, synth
from oracle.caption_oracle import CaptionOracle, decode_caption, resize_features

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference sources not mounted")


@pytest.mark.parametrize("att", synth.ATTENTION_TYPES)
def test_live_greedy_beam_masked(att):
    cfg = synth.make_config("tiny")
    V = cfg.model.vocab_size
    sd = synth.make_state_dict(cfg, V, att, seed=31, logit_gain=4.0)
    ref = ref_shim.build_reference_model(cfg, V, att, state_dict=sd)
    o = CaptionOracle(sd)
    x = torch.from_numpy(synth.make_features(3, 16, 256, seed=32))
    with torch.no_grad():
        r = ref.generate(x, 1, 2, max_length=8, method="greedy")
    q = o.greedy(x, 1, 2, max_length=8)
    assert torch.equal(r["generated_tokens"], q["generated_tokens"])
    mask = torch.ones(3, 16)
    mask[1, 11:] = 0
    mask[2, 7:] = 0
    with torch.no_grad():
        e1, f1 = ref.encoder(x, mask)
    e2, f2 = o.encode(x, mask)
    assert torch.allclose(e1, e2, atol=2e-6) and torch.allclose(f1, f2, atol=2e-6)
    with torch.no_grad():
        rb = ref.generate(x[:1], 1, 2, max_length=8, method="beam", beam_size=3)["generated_tokens"][0]
    qb = o.beam(x[:1], 1, 2, max_length=8, beam_size=3)["generated_tokens"][0]
    assert torch.equal(rb, qb)


def test_live_vocabulary_and_resize():
    vm = ref_shim.load_reference_vocabulary()
    cfg = synth.make_config("tiny")
    voc = vm.Vocabulary(cfg)
    for w in ["a", "man", "is", "running"]:
        voc.word2idx[w] = len(voc.word2idx)
        voc.idx2word[voc.word2idx[w]] = w
    toks = [1, 4, 5, 2, 6, 7, 0, 99]
    for rm in (True, False):
        assert voc.decode_caption(toks, rm) == decode_caption(toks, voc.idx2word, remove_special_tokens=rm)
    x = np.arange(200 * 3, dtype=np.float32).reshape(200, 3)
    idx = torch.linspace(0, 199, 80, dtype=torch.long)
    assert np.array_equal(resize_features(x, 80), x[idx.numpy()])
    assert resize_features(x[:10], 16).shape == (16, 3) and resize_features(x[:10], 16)[10:].sum() == 0


@pytest.mark.parametrize("att", synth.ATTENTION_TYPES)
def test_live_diverse_beam(att):
    """Real beam search: oracle.beam(diverse=True) vs the unmodified reference loop with scores[1:] = -inf."""
    cfg = synth.make_config("tiny")
    V = cfg.model.vocab_size
    sd = synth.make_state_dict(cfg, V, att, seed=33, logit_gain=4.0, end_token_id=2, end_bias=0.3)
    ref = ref_shim.build_reference_model(cfg, V, att, state_dict=sd)
    o = CaptionOracle(sd)
    x = torch.from_numpy(synth.make_features(3, 16, 256, seed=34, kind="ragged"))
    q = o.beam(x, 1, 2, max_length=10, beam_size=4, length_penalty=1.2, diverse=True)
    for b in range(3):
        r = ref_shim.reference_diverse_beam(ref, x[b:b + 1], 1, 2, 10, 4, 1.2)
        assert r.tolist() == q["generated_tokens"][b, : int(q["lengths"][b])].tolist()
