import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
import video_captioning_b200 as vc
from oracle import synth
cfg = synth.make_config("msvd"); cm = cfg.model; V = cm.vocab_size
sd = synth.make_state_dict(cfg, V, "bahdanau", seed=0)
m = vc.VideoCaptioningModel(cfg, V, precision="bf16", chunk_size=1024)
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.cuda().eval()
voc = vc.Vocabulary.from_words([f"w{i}" for i in range(V - 4)])
pred = vc.VideoCaptionPredictor.from_model(m, voc, config=cfg)
host = torch.randn(1024, 80, 4096).pin_memory()
vids = [host.numpy()[i] for i in range(1024)]
import cProfile, pstats
for i in range(3): pred.predict_batch(vids, method="beam", max_length=20, beam_size=5)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(3): pred.predict_batch(vids, method="beam", max_length=20, beam_size=5)
torch.cuda.synchronize()
print("predict_batch: %.1f ms per 1024" % ((time.perf_counter() - t0) / 3 * 1e3))
t0 = time.perf_counter()
for i in range(3):
    out = m.generate(host, voc.start_idx, voc.end_idx, max_length=20, method="beam", beam_size=5); out["generated_tokens"].cpu()
torch.cuda.synchronize()
print("generate(host): %.1f ms per 1024" % ((time.perf_counter() - t0) / 3 * 1e3))
pr = cProfile.Profile(); pr.enable()
pred.predict_batch(vids, method="beam", max_length=20, beam_size=5)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
print(m.host_stats)
