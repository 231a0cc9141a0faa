"""Two generate() calls of a bench.py workload with nothing else around them -- the command to put under ncu (the whole bench.py
under ncu costs minutes of intercepted launches):  python scripts/ncu_case.py [workload] [calls]
  ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip <launches of call 1> -c 900 --csv --log-file ... python scripts/ncu_case.py
CUDA graphs are off (every launch is a kernel launch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("VC_CUDA_GRAPHS", "0")
import torch

import video_captioning_b200 as vc
from bench import WORKLOADS
from oracle import synth

name = sys.argv[1] if len(sys.argv) > 1 else "c2_beam5_msvd_bf16"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = WORKLOADS[name]
cfg = synth.make_config(w["shape"])
cm = cfg.model
sd = synth.make_state_dict(cfg, cm.vocab_size, w["attention"], seed=0)
model = vc.VideoCaptioningModel(cfg, cm.vocab_size, attention_type=w["attention"], precision=w["precision"], chunk_size=min(w["B"], 2048))
model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
model = model.cuda().eval()
x = torch.randn(w["B"], cm.video_sequence_length, cm.cnn_feature_dim, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
kw = dict(beam_size=w["K"], length_penalty=1.0) if w["method"] == "beam" else {}
from video_captioning_b200 import _native
n0 = _native.launch_count()
for i in range(calls):
    out = model.generate(x, 1, 2, max_length=w["S"], method=w["method"], **kw)
    torch.cuda.synchronize()
    n1 = _native.launch_count()
    print(f"call {i}: tokens {tuple(out['generated_tokens'].shape)}, library launches so far {n1 - n0}", flush=True)
