"""One small invocation of the path for compute-sanitizer (scripts/sanitize.sh): `python scripts/sanitize_case.py CASE`.
CASE = tiny (fp32 + bf16 greedy / beam / diverse beam on the tiny golden shape, B=6) or b300 (bf16 beam, B=300, "small"
shape: persistent attention kernel, persistent GEMMs, persistent encoder recurrence, fused selection)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import video_captioning_b200 as vc
from oracle import synth

case = sys.argv[1] if len(sys.argv) > 1 else "tiny"
os.environ.setdefault("VC_CUDA_GRAPHS", "0")
if case == "tiny":
    cfg = synth.make_config("tiny")
    B, S, precs = 6, 6, ("fp32", "bf16")
elif case == "b300":
    cfg = synth.make_config("small")
    B, S, precs = 300, 4, ("bf16",)
else:
    cfg = synth.make_config("msvd")
    B, S, precs = int(case), 3, ("bf16",)
V = cfg.model.vocab_size
sd = synth.make_state_dict(cfg, V, "bahdanau", seed=7, logit_gain=4.0, end_token_id=2, end_bias=0.3)
x = torch.from_numpy(synth.make_features(B, cfg.model.video_sequence_length, cfg.model.cnn_feature_dim, seed=9, kind="ragged")).cuda()
for prec in precs:
    m = vc.VideoCaptioningModel(cfg, V, precision=prec)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.cuda().eval()
    g = m.generate(x, 1, 2, max_length=S)["generated_tokens"]
    b = m.generate(x, 1, 2, max_length=S, method="beam", beam_size=5)
    d = m.generate(x, 1, 2, max_length=S, method="beam", beam_size=5, diverse_beams=True, num_return_sequences=5)
    torch.cuda.synchronize()
    print(case, prec, "ok", g.shape, b["lengths"][:4].tolist(), d["nbest_lengths"][0].tolist())
