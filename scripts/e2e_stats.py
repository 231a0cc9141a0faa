"""End-to-end step (pinned host fp32 features -> generate -> tokens on the host) under different ingest settings:
   python scripts/e2e_stats.py  -- prints ms per step and the pipeline's own counters (video_captioning_model.host_stats)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import video_captioning_b200 as vc
from oracle import synth

cfg = synth.make_config("msvd")
cm = cfg.model
sd = synth.make_state_dict(cfg, cm.vocab_size, "bahdanau", seed=0)
x = torch.randn(1024, 80, 4096).pin_memory()


def run(tag, **attrs):
    m = vc.VideoCaptioningModel(cfg, cm.vocab_size, precision="bf16", chunk_size=1024)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.cuda().eval()
    for k, v in attrs.items():
        setattr(m, k, v)
    ts = []
    for i in range(7):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = m.generate(x, 1, 2, max_length=20, method="beam", beam_size=5)
        out["generated_tokens"].cpu()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    hs = m.host_stats
    print(f"{tag}: {min(ts[3:]) * 1e3:.1f} ms best of 4 ({1024 / min(ts[3:]):.0f} captions/s)  packed {hs.get('packed')} raw {hs.get('raw')} "
          f"pack {hs.get('pack_s', 0) * 1e3:.1f} ms loop {hs.get('loop_s', 0) * 1e3:.1f} ms sync {hs.get('sync_s', 0) * 1e3:.1f} ms", flush=True)


run("default")
# settings swept in round 2 (run-to-run spread of the box: +-2 ms, larger than any of their effects):
#   host_chunk_fractions (0.5,0.85,1.0) (0.6,0.9,1.0) (0.45,0.75,0.94,1.0) (0.75,1.0) (0.56,0.88,1.0); host_inflight 1 2 4 6;
#   host_piece_size 32 128
for infl in (2, 4):
    run(f"inflight {infl}", host_inflight=infl)
for piece in (32, 128):
    run(f"piece {piece}", host_piece_size=piece)
