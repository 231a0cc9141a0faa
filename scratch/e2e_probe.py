"""Timeline of the packed host ingest (bf16 mode) on the GPU box: how many pieces go raw / packed, where the loop waits."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_captioning_b200 as vc
from oracle import synth
cfg = synth.make_config("msvd"); V = cfg.model.vocab_size
sd = synth.make_state_dict(cfg, V, "bahdanau", seed=0)
m = vc.VideoCaptioningModel(cfg, V, precision="bf16", chunk_size=1024)
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.cuda().eval()
host = torch.randn(1024, 80, 4096).pin_memory()
for label, setup in (("default", {}), ("c3", dict(host_chunk_fractions=(0.375, 0.75, 1.0))), ("c4", dict(host_chunk_fractions=(0.25, 0.5, 0.75, 1.0))),
                     ("c3b", dict(host_chunk_fractions=(0.5, 0.8, 1.0))), ("c4b", dict(host_chunk_fractions=(0.4, 0.7, 0.9, 1.0))),
                     ("c2b", dict(host_chunk_fractions=(0.75, 1.0))), ("default2", {}), ("nopack", dict(host_pack=False))):
    m.host_pack_threads = 16; m.host_inflight = 3; m.host_chunk_fractions = (0.625, 1.0); m.host_pack = True; m._ingest_trials = {}
    for k, v in setup.items(): setattr(m, k, v)
    for _ in range(2):
        o = m.generate(host, 1, 2, max_length=20, method="beam", beam_size=5); o["generated_tokens"].cpu()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 8
    for _ in range(n):
        o = m.generate(host, 1, 2, max_length=20, method="beam", beam_size=5); o["generated_tokens"].cpu()
    dt = (time.perf_counter() - t0) / n
    st = getattr(m, "host_stats", {})
    print(f"{label:10s} {dt*1e3:6.1f} ms/1024 -> {1024/dt:7.0f} cap/s  stats", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()})
