import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import synth
from _util import make_native_model
START, END = 1, 2
cfg = synth.make_config("msvd"); V = cfg.model.vocab_size
sd = synth.make_state_dict(cfg, V, "bahdanau", seed=0, logit_gain=8.0, end_token_id=END, end_bias=0.45)
m = make_native_model(cfg, V, sd, "bahdanau", "bf16")
x = torch.from_numpy(synth.make_features(64, 80, 4096, seed=3, kind="ragged")).cuda()
h = m._handle()
# 1. encoder determinism
e1, f1 = h.encoder_forward(x); e2, f2 = h.encoder_forward(x)
print("encoder deterministic:", torch.equal(e1, e2), torch.equal(f1, f2))
# 2. teacher forced logits determinism
tok = torch.randint(4, V, (64, 12), device="cuda")
l1, a1, _ = h.forward_teacher(x, tok); l2, a2, _ = h.forward_teacher(x, tok)
print("teacher logits deterministic:", torch.equal(l1, l2), (l1 - l2).abs().max().item())
# 3. batch-position independence: rows 0..7 alone vs in batch
l3, _, _ = h.forward_teacher(x[:8], tok[:8])
print("batch independence:", torch.equal(l1[:8], l3), (l1[:8] - l3).abs().max().item())
e3, f3 = h.encoder_forward(x[:8])
print("encoder batch independence:", torch.equal(e1[:8], e3), (e1[:8]-e3).abs().max().item(), torch.equal(f1[:8], f3))
# 4. greedy vs beam
g1 = m.generate(x, START, END, max_length=20)["generated_tokens"].cpu()
g2 = m.generate(x, START, END, max_length=20)["generated_tokens"].cpu()
print("greedy deterministic:", torch.equal(g1, g2))
for K in (1, 3, 5):
    b = m.generate(x, START, END, max_length=20, method="beam", beam_size=K)
    t, l = b["generated_tokens"].cpu(), b["lengths"].cpu()
    bad = 0
    for i in range(64):
        row = g1[i].tolist()
        if END in row: row = row[:row.index(END) + 1]
        n = min(len(row) + 1, int(l[i]))
        if t[i, :n].tolist() != ([START] + row)[:n]:
            bad += 1
            if bad <= 2: print("  K", K, "row", i, t[i, :n].tolist(), ([START] + row)[:n])
    print("beam K=%d mismatching rows: %d" % (K, bad))
