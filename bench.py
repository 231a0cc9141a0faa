#!/usr/bin/env python
"""bench.py -- captions/sec of the batched caption-generation path on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (CPU arm: the reference algorithm on host cores)

A "step" = one pass of the whole hot path (bi-LSTM encoder -> hoisted attention projections -> S decode
steps with beam search) over one batch of synthetic videos.  Default workload = BASELINE.json configs[1]:
beam-5, 1024 MSVD-shape videos (80 x 4096 features, H=E=A=512, V=10k, max_len 20), Bahdanau, bf16.
`value`  : captions/s with the features already resident in HBM (CUDA events, max over ranks).
`e2e`    : captions/s through the public API VideoCaptioningModel.generate called with HOST (pinned) features:
           the H2D copies (chunked, overlapped with compute), the whole path and the D2H of tokens/lengths
           are inside the timed region.
`roofline`: for the kernel class with the largest share of device time, measured with CUDA events on
           the launching stream in an instrumented pass of the same workload (vc_profile_begin/end).
`cpu_baseline`: the reference itself (oracle/_ref: the unmodified model files, vendored by __graft_entry__.build();
           `kind` "reference") -- or, where that copy is absent, the oracle port (`kind` "port") -- timed on the host
           cores on a bounded sample, one B=1 call per video as predict.py batch does.
`sustained`: the same device-resident step repeated back to back for >= 2 s (clocks sampled under load).
`other_configs`: short runs of BASELINE.json configs[0], [2], [3], [4] at their stated sizes (N=1 only).
Weights are random-init (oracle.synth, reference layout/initialiser distributions); data synthetic.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

START, END = 1, 2

WORKLOADS = {
    # name: shape, attention, method, K, S, per-GPU batch, precision
    "c2_beam5_msvd_bf16": dict(shape="msvd", attention="bahdanau", method="beam", K=5, S=20, B=1024, precision="bf16"),
    "c1_greedy_msvd_fp32": dict(shape="msvd", attention="bahdanau", method="greedy", K=1, S=20, B=32, precision="fp32"),
    "c2_beam5_msvd_fp32": dict(shape="msvd", attention="bahdanau", method="beam", K=5, S=20, B=256, precision="fp32"),
    "c3_luong_general_h1024": dict(shape="c3", attention="luong_general", method="beam", K=5, S=20, B=1024, precision="bf16"),
    "c3_luong_dot_h1024": dict(shape="c3", attention="luong_dot", method="beam", K=5, S=20, B=1024, precision="bf16"),
    "c4_multihead_resnet": dict(shape="c4", attention="multihead", method="beam", K=3, S=20, B=1024, precision="bf16"),
    "c5_vocab30k_len30": dict(shape="c5", attention="bahdanau", method="beam", K=5, S=30, B=1024, precision="bf16"),
    # BASELINE configs[4] as written: 65 536 videos in total, sharded over the ranks (strong scaling)
    "c5_sharded_65536": dict(shape="c5", attention="bahdanau", method="beam", K=5, S=30, B=65536, precision="bf16",
                             total=True),
}
DEFAULT_WORKLOAD = "c2_beam5_msvd_bf16"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, source="fallback")


def load_traffic(workload, B):
    """Per-launch DRAM bytes of the kernel classes from the newest committed ncu full capture (profiles/*_traffic.json);
    only valid for the workload and batch it was captured on (the default one)."""
    import glob
    if workload != DEFAULT_WORKLOAD or B != WORKLOADS[DEFAULT_WORKLOAD]["B"]:
        return None
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None
    d = json.load(open(files[-1]))
    d["file"] = os.path.relpath(files[-1], ROOT)
    return d


# ---------------------------------------------------------------------------- algorithmic work per kernel class
def class_work(wl, cfgm, B):
    """Algorithmic FLOPs and HBM bytes per STEP (whole batch) for each kernel class.
    bytes are the unavoidable traffic at the class's operand types (b = bytes/element of activations)."""
    H, E, A, F, V, T = (cfgm.encoder_hidden_dim, cfgm.embedding_dim, cfgm.attention_dim, cfgm.cnn_feature_dim,
                        cfgm.vocab_size, cfgm.video_sequence_length)
    K, S = wl["K"], wl["S"]
    Le, Ld = cfgm.encoder_num_layers, cfgm.decoder_num_layers
    R = B * K
    b = 2 if wl["precision"] == "bf16" else 4
    att = wl["attention"]
    w = {}
    w["convert"] = dict(flops=0, bytes=B * T * F * (4 + 2))
    w["enc_feature_proj"] = dict(flops=2 * B * T * F * H, bytes=B * T * F * b + H * F * b + B * T * H * b)
    fl = by = 0
    for l in range(Le):
        inp = H if l == 0 else 2 * H
        fl += 2 * B * T * inp * 8 * H
        by += B * T * inp * b + 8 * H * inp * b + B * T * 8 * H * b
    w["enc_input_proj"] = dict(flops=fl, bytes=by)
    # per layer, per direction, per timestep: h[B,H] . W_hh[4H,H]^T, + xproj read, h write, c r/w
    w["enc_recurrent"] = dict(flops=Le * 2 * T * 2 * B * H * 4 * H,
                              bytes=Le * 2 * T * (B * H * b + 4 * H * H * b + B * 4 * H * b + B * H * b + 2 * B * H * 4))
    w["enc_output_proj"] = dict(flops=2 * (B * T + B) * 2 * H * H, bytes=B * T * 2 * H * b + B * T * H * b)
    if att in ("bahdanau", "luong_concat"):
        w["attn_precompute"] = dict(flops=2 * B * T * H * A, bytes=B * T * H * b + B * T * A * b)
        w["attn_query_proj"] = dict(flops=S * 2 * R * H * A, bytes=S * (R * H * b + A * H * b + R * A * 4))
        # keys + enc_out read once per VIDEO-step, q read, ctx write
        w["attn_step"] = dict(flops=S * R * (2 * T * A + 2 * T * H), tanh=S * R * T * A,
                              bytes=S * (B * T * (A + H) * b + R * A * 4 + R * H * b))
    elif att == "multihead":
        w["attn_precompute"] = dict(flops=4 * B * T * H * H, bytes=B * T * H * b * 3)
        w["attn_query_proj"] = dict(flops=S * 2 * R * H * H, bytes=S * (R * H * b + H * H * b + R * H * 4))
        w["attn_step"] = dict(flops=S * R * 4 * T * H, bytes=S * (B * T * 2 * H * b + R * H * 4 + R * H * b))
        w["attn_output_proj"] = dict(flops=S * 2 * R * H * H, bytes=S * (2 * R * H * b + H * H * b))
    else:
        w["attn_precompute"] = dict(flops=0, bytes=0)
        if att == "luong_general":
            w["attn_query_proj"] = dict(flops=S * 2 * R * H * H, bytes=S * (R * H * b + H * H * b + R * H * 4))
        w["attn_step"] = dict(flops=S * R * 4 * T * H, bytes=S * (B * T * H * b + R * H * 4 + R * H * b))
    fl = by = 0
    for l in range(Ld):
        kin = (E + H if l == 0 else H) + H
        fl += 2 * R * kin * 4 * H
        by += R * kin * b + 4 * H * kin * b + 2 * R * H * b + 2 * R * H * 4
    w["dec_lstm"] = dict(flops=S * fl, bytes=S * by)
    w["dec_context_proj"] = dict(flops=S * 2 * R * (2 * H + E) * H, bytes=S * (R * (2 * H + E) * b + (2 * H + E) * H * b + R * H * b))
    tn = (V + 255) // 256
    if wl["precision"] == "bf16" and V >= 256 and 8 * tn <= 1024:
        # fused selection: the GEMM emits per-row chunk maxima (8 per 256-column tile) and
        # log-sum-exp partials (2 float2 per tile) and only the >= K chunks of 128 bytes that can reach the top-K;
        # the selection reads the statistics and K chunks per row
        stats = R * (8 * tn + 2 * 2 * tn) * 4
        w["dec_vocab"] = dict(flops=S * 2 * R * H * V, bytes=S * (R * H * b + V * H * b + stats + R * K * 128))
        w["select"] = dict(flops=0, bytes=S * (stats + R * K * 128))
    else:
        # logits materialised in fp32 (written by the GEMM, read by the streaming selection)
        w["dec_vocab"] = dict(flops=S * 2 * R * H * V, bytes=S * (R * H * b + V * H * b + R * V * 4))
        w["select"] = dict(flops=0, bytes=S * R * V * 4)
    w["reorder_embed"] = dict(flops=0, bytes=(S - 1) * R * (Ld * (2 * H * b + 2 * H * 4) + 2 * E * b))
    w["misc"] = dict(flops=0, bytes=0)
    return w


TENSOR_CLASSES = {"enc_feature_proj", "enc_input_proj", "enc_recurrent", "enc_output_proj", "attn_precompute",
                  "attn_query_proj", "attn_output_proj", "dec_lstm", "dec_context_proj", "dec_vocab"}


# ---------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_ready(self, timeout=4.0):
        """nvidia-smi needs up to a second to produce its first row on a fresh box: do not start the (short) timed region
        before the sampler is actually sampling."""
        t = time.time()
        while self.proc is not None and not self.rows and time.time() - t < timeout:
            time.sleep(0.02)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        return self.window(t0, t1)

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                smax = float(f[1])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[0]))
                    for n, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(n)
            except ValueError:
                continue
        note = None
        if not sm:
            # no row fell inside the timed window (it is tens of milliseconds long): use the rows closest to it
            near = sorted(((abs(ts - 0.5 * (t0 + t1)), line) for ts, line in self.rows), key=lambda x: x[0])[:2]
            for _, line in near:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0]))
                    for n, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(n)
                except (ValueError, IndexError):
                    continue
            note = "no sample inside the timed window; nearest samples used"
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
               "samples": len(sm)}
        if note:
            out["note"] = note
        return out


# ---------------------------------------------------------------------------- CPU arm (the reference on the host cores)
def cpu_reference_run(wl, n_videos, threads=None):
    """Times the reference's own CPU path on the host cores: the UNMODIFIED reference modules (oracle/_ref, or
    /root/reference where mounted) when present -- `VideoCaptioningModel.generate` called once per video with B=1, which
    is what predict.py batch does (predictor.py:102,217,464; a batched beam call raises on staggered END) --
    else the oracle port run the same way.  Returns (captions/s, seconds, threads, kind)."""
    from oracle import ref_shim, synth
    torch.set_num_threads(threads or os.cpu_count() or 1)
    cfg = synth.make_config(wl["shape"])
    V = cfg.model.vocab_size
    sd = synth.make_state_dict(cfg, V, wl["attention"], seed=0)
    feats = synth.make_features(n_videos, cfg.model.video_sequence_length, cfg.model.cnn_feature_dim, seed=1)
    kw = dict(beam_size=wl["K"], length_penalty=1.0) if wl["method"] == "beam" else {}
    if ref_shim.available():
        kind = "reference"
        model = ref_shim.build_reference_model(cfg, V, wl["attention"], state_dict=sd)
        x = torch.from_numpy(feats)
        t0 = time.perf_counter()
        with torch.no_grad():
            for b in range(n_videos):       # predictor.py:102: unsqueeze(0) -> B = 1
                model.generate(x[b:b + 1], START, END, max_length=wl["S"], method=wl["method"], **kw)
        dt = time.perf_counter() - t0
    else:
        kind = "port"
        from oracle.caption_oracle import CaptionOracle
        o = CaptionOracle(sd)
        t0 = time.perf_counter()
        with torch.no_grad():
            for b in range(n_videos):
                if wl["method"] == "beam":
                    o.beam(feats[b:b + 1], START, END, max_length=wl["S"], beam_size=wl["K"])
                else:
                    o.greedy(feats[b:b + 1], START, END, max_length=wl["S"])
        dt = time.perf_counter() - t0
    return n_videos / dt, dt, torch.get_num_threads(), kind


def cpu_sample_text(n, dt, kind, wl):
    from oracle import ref_shim
    where = "oracle/_ref, the vendored copy" if ref_shim.is_vendored_copy() else ref_shim.REF_ROOT
    what = (f"the unmodified reference ({where}), VideoCaptioningModel.generate" if kind == "reference"
            else "oracle port of the reference")
    return f"{n} videos ({dt:.1f} s), {what}, one B=1 {wl['method']} call per video as predict.py batch does"


def run_reference_arm(args, wl, wl_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_videos
    vals = []
    n_warm = 1 if args.warmup else 0
    for _ in range(n_warm):
        cpu_reference_run(wl, max(2, n // 8))
    for _ in range(max(1, min(args.steps, 3))):
        vals.append(cpu_reference_run(wl, n))
    v = float(np.median([x[0] for x in vals]))
    dt = float(np.median([x[1] for x in vals]))
    th, kind = vals[0][2], vals[0][3]
    line = {"impl": "reference", "metric": "captions/sec", "value": v, "unit": "captions/s", "n_gpus": args.gpus,
            "steps": len(vals), "warmup": n_warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl_name, "sample_videos": n, **{k: wl[k] for k in ("method", "K", "S", "attention")}},
            "cpu_baseline": {"value": v, "unit": "captions/s", "cores": th, "kind": kind,
                             "sample": cpu_sample_text(n, dt, kind, wl) + " (per step)"},
            "e2e": {"value": v, "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="videos per GPU per step (default: workload's)")
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16"])
    ap.add_argument("--cpu-videos", type=int, default=384, help="bounded CPU-baseline sample (videos), both arms: ~12 s at the ~30 captions/s of a 16-core host")
    ap.add_argument("--no-sweep", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--sweep-json", default=None, help="also write the other-configs sweep here")
    ap.add_argument("--min-sustained-s", type=float, default=2.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-json", default=None, help="write the per-class device-time breakdown here")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["B"] = args.batch
    if args.precision:
        wl["precision"] = args.precision

    if args.impl == "reference":
        run_reference_arm(args, wl, args.workload)
        return

    import torch.distributed as dist
    import video_captioning_b200 as vc
    from oracle import synth   # synthetic weights/features only (shared recipe); never on the timed path
    from video_captioning_b200 import _native
    from video_captioning_b200.affinity import bind_to_gpu_numa
    from video_captioning_b200.sharding import gather_captions_equal

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # host side of the ingest (pinned buffers, packer threads) on the GPU's NUMA node / this rank's share of the cores
    affinity = bind_to_gpu_numa(local_rank, local_world, local_rank) if world > 1 else {"bound": False}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    strong = bool(wl.get("total"))
    if strong:       # a fixed total split over the ranks (BASELINE configs[4])
        assert wl["B"] % world == 0
        wl["B"] = wl["B"] // world

    def setup(w, seed):
        cfg = synth.make_config(w["shape"])
        cm = cfg.model
        sd = synth.make_state_dict(cfg, cm.vocab_size, w["attention"], seed=0)
        model = vc.VideoCaptioningModel(cfg, cm.vocab_size, attention_type=w["attention"], precision=w["precision"],
                                        chunk_size=min(w["B"], 2048))
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        model = model.to(dev).eval()
        # synthetic features generated on the device (per-rank seed), in slabs: resident in HBM before any timed region
        g = torch.Generator(device=dev).manual_seed(seed)
        feats = torch.empty(w["B"], cm.video_sequence_length, cm.cnn_feature_dim, device=dev, dtype=torch.float32)
        for lo in range(0, w["B"], 1024):
            feats[lo:lo + 1024].normal_(generator=g)
        return cfg, model, feats

    cfg, model, feats = setup(wl, 1234 + rank)
    cm = cfg.model
    V, T, F = cm.vocab_size, cm.video_sequence_length, cm.cnn_feature_dim
    B, K, S = wl["B"], wl["K"], wl["S"]
    kw = dict(beam_size=K, length_penalty=1.0) if wl["method"] == "beam" else {}

    def step(x, m=model, w=wl, k=kw):
        return m.generate(x, START, END, max_length=w["S"], method=w["method"], **k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(out):
        if world > 1:   # the path's only collective: final caption gather (latency-bound, fixed [B, S+2] shape)
            tk = out["generated_tokens"]
            width = S + 1 if wl["method"] == "beam" else S
            if tk.shape[1] < width:
                tk = torch.nn.functional.pad(tk, (0, width - tk.shape[1]), value=START)
            ln = out["lengths"] if "lengths" in out else torch.full((B,), tk.shape[1], device=dev)
            gather_captions_equal(tk, ln)

    def timed(n_steps):
        """n_steps back-to-back steps bracketed by barrier + synchronize; CUDA events; max over ranks -> (ms, t0, t1)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.time()
        e0.record()
        for _ in range(n_steps):
            gather(step(feats))
        e1.record()
        barrier()
        t1 = time.time()
        tms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        return float(tms.item()), t0, t1

    for _ in range(max(args.warmup, 3)):
        out = step(feats)
    barrier()

    # ---- sustained pass: the step repeated back to back for >= min_sustained_s.  It is a measurement of its own (clocks
    # sampled under seconds of load) and it brings the GPU to its steady state before the K timed steps, which would
    # otherwise be ~0.15 s of burst clocks on an idle box.
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.wait_ready()
    probe_ms, _, _ = timed(2)
    n_sus = max(args.steps, int(np.ceil(args.min_sustained_s * 1e3 / max(probe_ms / 2, 1e-3))))
    if world > 1:
        t = torch.tensor([n_sus], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_sus = int(t.item())
    sus_ms, st0, st1 = timed(n_sus)
    sus_clocks = sampler.window(st0, st1) if sampler else None

    # ---- timed region: EXACTLY K steps, device-resident inputs (B*T*F*4 bytes, far larger than the 126 MB L2)
    l0 = _native.launch_count()
    torch.cuda.nvtx.range_push("timed")     # ncu --nvtx --nvtx-include "timed/" profiles exactly these steps
    ms, t0, t1 = timed(args.steps)
    torch.cuda.nvtx.range_pop()
    launches = _native.launch_count() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    value = world * B * args.steps / (ms * 1e-3)
    sustained = {"value": world * B * n_sus / (sus_ms * 1e-3), "unit": "captions/s", "steps_run": n_sus,
                 "seconds": sus_ms * 1e-3, "clocks": sus_clocks}

    # ---- e2e through the public API with HOST (pinned) fp32 features: H2D inside the timed region, D2H of the results
    Be = min(B, 2048)
    host = torch.empty(Be, T, F, dtype=torch.float32).pin_memory()
    host.copy_(feats[:Be].cpu())
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        o = step(host)          # pinned host tensor: generate() streams it in pieces overlapped with compute
        _ = o["generated_tokens"].cpu()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    d2h = 0
    for _ in range(e2e_steps):
        o = step(host)
        tk = o["generated_tokens"].cpu()
        d2h = tk.numel() * tk.element_size()
        if "lengths" in o:
            ln = o["lengths"].cpu()
            d2h += ln.numel() * ln.element_size()
    e1.record()
    barrier()
    ems = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = world * Be * e2e_steps / (float(ems.item()) * 1e-3)
    # bf16 mode: part of the batch is rounded to bf16 on the host cores and crosses the link at half the size
    # (VideoCaptioningModel._generate_from_host_packed); the bytes are those of the last step's actual copies
    h2d = int(getattr(model, "host_stats", {}).get("h2d_bytes", 0)) or int(host.numel() * 4)
    packing = bool(getattr(model, "host_pack", False)) and wl["precision"] == "bf16"
    # the ingest's own roofline: H2D rate of one pinned 1 GiB copy per rank, all ranks at once, and the host cores' packing
    # rate (fp32 bytes consumed per second), measured here; ceiling = the captions/s the faster of "all fp32 over the link"
    # and "the best raw/packed split" could reach if nothing else took time
    probe = torch.empty(1 << 28, dtype=torch.float32).pin_memory()
    dprobe = torch.empty_like(probe, device=dev)
    dprobe.copy_(probe, non_blocking=True)
    barrier()
    e0.record()
    for _ in range(2):
        dprobe.copy_(probe, non_blocking=True)
    e1.record()
    barrier()
    lms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(lms, op=dist.ReduceOp.MAX)
    link_gbs = 2 * probe.numel() * 4 / (float(lms.item()) * 1e-3) / 1e9
    pack_gbs = None
    if packing:
        dst16 = torch.empty(probe.numel() // 4, dtype=torch.bfloat16).pin_memory()
        src32 = probe[: probe.numel() // 4]
        _native.host_pack_bf16(src32, dst16, model.host_pack_threads)
        tp = time.perf_counter()
        _native.host_pack_bf16(src32, dst16, model.host_pack_threads)
        pack_gbs = src32.numel() * 4 / (time.perf_counter() - tp) / 1e9
        del dst16
    del probe, dprobe
    per_cap = T * F * 4.0                          # fp32 bytes of one video
    ceil_plain = world * link_gbs * 1e9 / per_cap
    ceiling = ceil_plain
    if pack_gbs:
        # a fraction r of the videos crosses raw (4 B/elem), 1-r packed (2 B/elem over the link, 4 B/elem through the cores):
        # time per video = max(link: (r + (1-r)/2) * per_cap / L, cores: (1-r) * per_cap / P); best r equalises the two
        L_, P_ = link_gbs * 1e9, pack_gbs * 1e9
        r = max(0.0, min(1.0, (2 * L_ - P_) / (2 * L_ + P_)))
        ceiling = world / max((r + (1 - r) / 2) * per_cap / L_, (1 - r) * per_cap / P_)
    e2e = {"value": e2e_value, "unit": "captions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h),
           "host_bytes_per_step": int(host.numel() * 4), "videos_per_gpu_per_step": Be,
           "link_gbs": link_gbs, "host_pack_gbs": pack_gbs, "ceiling": ceiling, "ceiling_fp32_link_only": ceil_plain,
           "frac_of_ceiling": e2e_value / ceiling, "affinity": affinity,
           "ingest": ("fp32 host features; pieces go raw (fp32 H2D + device rounding) or host-packed to bf16 (host cores, "
                      f"{getattr(model, 'host_pack_threads', 0)} threads), whichever route is free") if h2d != host.numel() * 4 else "fp32 H2D"}
    # the same through the Predictor (predict.py batch): a list of per-video numpy arrays -> one native staging pass
    # (resize + bf16 rounding into a reused pinned buffer) -> generate()'s host pipeline -> token rows -> caption strings
    if world == 1 and not strong:
        voc = vc.Vocabulary.from_words([f"w{i}" for i in range(V - 4)])
        pred = vc.VideoCaptionPredictor.from_model(model, voc, config=cfg)
        host_np = host.numpy()
        vids = [host_np[i] for i in range(Be)]
        pk = dict(method=wl["method"], max_length=S, **({"beam_size": K} if wl["method"] == "beam" else {}))
        pred.predict_batch(vids, **pk)
        torch.cuda.synchronize()
        tp = time.perf_counter()
        n_pred = 3
        for _ in range(n_pred):
            pred.predict_batch(vids, **pk)
        torch.cuda.synchronize()
        e2e["predictor"] = {"value": Be * n_pred / (time.perf_counter() - tp), "unit": "captions/s",
                            "what": "VideoCaptionPredictor.predict_batch on a list of per-video fp32 numpy arrays, "
                                    "captions decoded to strings (wall clock)"}
        del pred, vids, host_np
    del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- instrumented pass: per-class device time with CUDA events on the launching stream
    prof_steps = 2
    _native.profile_begin()
    for _ in range(prof_steps):
        step(feats)
    prof = _native.profile_end()
    peaks = load_peaks()
    work = class_work(wl, cm, B)
    if prof.get("reorder_embed", {"scopes": 0})["scopes"] == 0 and "attn_step" in work:
        # the reorder / embedding gather rode along inside the attention kernel (attention.cuh: RowGather): its bytes are
        # that kernel's algorithmic bytes now
        work["attn_step"]["bytes"] += work["reorder_embed"]["bytes"]
        work["attn_step"]["includes"] = "reorder_embed row gather"
    total_ms = sum(v["ms"] for v in prof.values()) / prof_steps
    breakdown = {}
    for cls, v in prof.items():
        if v["scopes"] == 0:
            continue
        cms = v["ms"] / prof_steps
        wk = work.get(cls, dict(flops=0, bytes=0))
        ent = {"ms_per_step": cms, "share": cms / total_ms, "launches_per_step": v["scopes"] / prof_steps,
               "tflops": wk["flops"] / (cms * 1e-3) / 1e12 if cms > 0 else 0.0,
               "gbs": wk["bytes"] / (cms * 1e-3) / 1e9 if cms > 0 else 0.0}
        breakdown[cls] = ent
    dom = max(breakdown, key=lambda c: breakdown[c]["ms_per_step"])
    d = breakdown[dom]
    per_launch = d["launches_per_step"]
    if dom in TENSOR_CLASSES and wl["precision"] == "bf16":
        # kernels timed inside a long step -> sustained cuBLAS figure
        roof = {"kernel": dom, "bound": "tensor", "achieved": d["tflops"], "peak": peaks["tf_sus"], "unit": "TFLOP/s",
                "frac": d["tflops"] / peaks["tf_sus"], "traffic": None,
                "avg_launch_ms": d["ms_per_step"] / per_launch, "flops_per_launch": work[dom]["flops"] / per_launch}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": d["gbs"], "peak": peaks["hbm"], "unit": "GB/s",
                "frac": d["gbs"] / peaks["hbm"], "traffic": None,
                "avg_launch_ms": d["ms_per_step"] / per_launch, "bytes_per_launch": work[dom]["bytes"] / per_launch}
    roof["peak_source"] = peaks["source"]
    roof["share_of_step"] = d["share"]
    # DRAM traffic of that kernel per launch, from the committed ncu --set full capture of this workload (profiles/)
    tr = load_traffic(args.workload, B)
    if tr is not None and dom in tr["classes"]:
        roof["traffic"] = tr["classes"][dom]["dram_bytes_per_launch"]
        roof["traffic_source"] = tr["file"]
    if "tanh" in work.get(dom, {}):
        # Bahdanau scoring is bound by the special-function pipe (16 tanh / clk / SM, measured with scripts/mufu_bench.cu),
        # not by HBM: report that roofline as well
        sm_hz = ((clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0) * 1e6
        xu_peak = 16.0 * torch.cuda.get_device_properties(dev).multi_processor_count * sm_hz
        ach = work[dom]["tanh"] / (d["ms_per_step"] * 1e-3)
        roof["xu"] = {"bound": "mufu", "achieved": ach, "peak": xu_peak, "unit": "tanh/s", "frac": ach / xu_peak,
                      "note": "peak = 16 MUFU results/clk/SM x SMs x sampled SM clock"}
    # every class against its own bound (tensor classes: sustained bf16 peak; the others: HBM)
    for cls, ent in breakdown.items():
        if cls in TENSOR_CLASSES and wl["precision"] == "bf16":
            ent["frac"] = ent["tflops"] / peaks["tf_sus"]
            ent["bound"] = "tensor"
        else:
            ent["frac"] = ent["gbs"] / peaks["hbm"]
            ent["bound"] = "hbm"

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, dt, th, kind = cpu_reference_run(wl, args.cpu_videos)
        cpu = {"value": v, "unit": "captions/s", "cores": th, "kind": kind,
               "sample": cpu_sample_text(args.cpu_videos, dt, kind, wl)}

    # ---- the other BASELINE configs at their stated sizes, short runs (N=1, default workload only)
    other = None
    if world == 1 and args.workload == DEFAULT_WORKLOAD and not args.no_sweep and not args.batch:
        del feats, model
        torch.cuda.empty_cache()
        other = {}
        sweep = [("configs[0] greedy B=32 fp32", dict(WORKLOADS["c1_greedy_msvd_fp32"])),
                 ("configs[2] Luong general H=1024 beam-5 B=4096", dict(WORKLOADS["c3_luong_general_h1024"], B=4096)),
                 ("configs[2] Luong dot H=1024 beam-5 B=4096", dict(WORKLOADS["c3_luong_dot_h1024"], B=4096)),
                 ("configs[3] multi-head(8) ResNet 2048-d T=40 beam-3, 3 captions per video B=1024",
                  dict(WORKLOADS["c4_multihead_resnet"], multiple=3)),
                 ("configs[4] V=30k max_len 30 beam-5, one 8192-video shard (the per-GPU share of 65536 at N=8)",
                  dict(WORKLOADS["c5_vocab30k_len30"], B=8192))]
        for name, w in sweep:
            try:
                _, m2, f2 = setup(w, 99)
                k2 = dict(beam_size=w["K"], length_penalty=1.0) if w["method"] == "beam" else {}
                if w.get("multiple"):      # predict.py multiple: n-best from one real beam search (opt-in mode)
                    k2.update(diverse_beams=True, num_return_sequences=w["multiple"])
                for _ in range(3):
                    step(f2, m2, w, k2)
                torch.cuda.synchronize()
                e0.record()
                n2 = 5
                for _ in range(n2):
                    step(f2, m2, w, k2)
                e1.record()
                torch.cuda.synchronize()
                ms2 = e0.elapsed_time(e1) / n2
                other[name] = {"value": w["B"] / (ms2 * 1e-3), "unit": "captions/s", "ms_per_step": ms2, "videos": w["B"],
                               "dtype": w["precision"], "steps": n2}
                del m2, f2
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001 -- a sweep entry must not take the headline line down
                other[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        if args.sweep_json:
            with open(args.sweep_json, "w") as f:
                json.dump(other, f, indent=1)

    line = {"metric": "captions/sec", "value": value, "unit": "captions/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": wl["precision"], "data": "synthetic",
            "config": {"workload": args.workload, "videos_per_gpu_per_step": B, "frames": T, "feature_dim": F,
                       "hidden": cm.encoder_hidden_dim, "vocab": V, "method": wl["method"], "beam": K, "max_len": S,
                       "attention": wl["attention"], "l2": "inputs (B*T*F*4 bytes) larger than L2, no flush needed",
                       "sharding": f"dp{world}: videos split over ranks, final NCCL all_gather of tokens",
                       "preload": f"{n_sus} untimed back-to-back steps ({sus_ms * 1e-3:.2f} s, reported as `sustained`) right before the timed steps"},
            "clocks": clocks, "sustained": sustained, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
            "cpu_baseline": cpu, "breakdown": breakdown, "other_configs": other}
    print(json.dumps(line), flush=True)
    if args.profile_json:
        with open(args.profile_json, "w") as f:
            json.dump({"workload": args.workload, "config": line["config"], "ms_per_step": line["ms_per_step"],
                       "breakdown": breakdown, "roofline": roof, "peaks": peaks}, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
