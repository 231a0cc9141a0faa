import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import synth
from _util import make_native_model
cfg = synth.make_config("msvd"); V = cfg.model.vocab_size
sd = synth.make_state_dict(cfg, V, "bahdanau", seed=0)
m = make_native_model(cfg, V, sd, "bahdanau", "bf16")
h = m._handle()
for B in [int(a) for a in sys.argv[1:]]:
    x = torch.randn(B, 80, 4096, device="cuda")
    try:
        e, f = h.encoder_forward(x)
        torch.cuda.synchronize()
        print("B", B, "ok", float(e.abs().mean()), float(f.abs().mean()), flush=True)
        if os.environ.get("SAVE"):
            torch.save({"e": e.cpu(), "f": f.cpu()}, f"gpurun_out/enc_{os.environ['SAVE']}_{B}.pt")
    except Exception as ex:
        print("B", B, "FAILED", str(ex)[:200], flush=True)
        break
