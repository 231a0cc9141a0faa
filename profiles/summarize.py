#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.
  python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python profiles/summarize.py raw gpurun_out/prof_r1_decode.ncu-rep profiles/r1_decode_kernels.md
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu (MUFU) pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global ld sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global ld requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global st sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global st requests"),
    ("lts__t_bytes.sum", "L2 bytes"),
]


def short(name):
    return name.replace("vc::", "").replace("tc::", "").replace("(int)", "").replace("(bool)", "")[:110]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(short(r[ik]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` window of "
                f"{sum(v[0] for v in agg.values())} consecutive launches, {tot / 1e3:.1f} us total (cold-cache, serialised: compare shares).\n\n")
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n}` | {c} | {t / 1e3:.1f} | {100 * t / tot:.1f}% | {t / c / 1e3:.1f} |\n")


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\nOne capture per kernel (`--clock-control none`, ~40 replays: durations are not bench values).\n")
        for r in rows[2:]:
            f.write(f"\n## `{short(r[idx['Kernel Name']])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m, label in METRICS:
                if m in idx:
                    f.write(f"| {label} (`{m}`) | {r[idx[m]]} | {units[idx[m]]} |\n")
            stalls = [(h, r[i]) for h, i in idx.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
            top = sorted(((h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(v or 0)) for h, v in stalls), key=lambda x: -x[1])[:5]
            f.write("| top stall reasons (samples) | " + ", ".join(f"{h} {int(v)}" for h, v in top) + " | |\n")


# kernel name fragment -> bench.py kernel class (VC_CLS_*), benchmark workload
CLASS_OF = [
    ("gemm_tc_persistent_kernel<4, 0, __nv_bfloat16, 0, 0, 1, 1, 0>", "enc_feature_proj"),
    ("lstm_layer_persistent_kernel", "enc_recurrent"), ("lstm_layer_pair_kernel", "enc_recurrent"),
    ("gemm_tc_kernel<128, 3, 2, 0, __half", "attn_query_proj"),
    ("attn_additive_ws_kernel", "attn_step"),
    ("gemm_tc_persistent_kernel<5, 1,", "dec_lstm"),
    ("gemm_tc_kernel<128, 3, 2, 0, __nv_bfloat16, 1", "dec_context_proj"),
    ("gemm_tc_persistent_kernel<3, 0, float, 0, 1", "dec_vocab"),
    ("select_fused_kernel", "select"),
    ("reorder_embed_kernel", "reorder_embed"),
]


def traffic(src, dst):
    """Per-launch DRAM bytes of the kernel classes from a --set full capture -> the JSON bench.py reads (roofline.traffic)."""
    import json
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    acc = collections.OrderedDict()
    for r in rows[2:]:
        name = short(r[idx["Kernel Name"]])
        cls = next((c for frag, c in CLASS_OF if frag in name), None)
        if cls is None:
            continue
        b = sum(float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t = float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * tscale[units[idx["gpu__time_duration.sum"]]]
        a = acc.setdefault(cls, [0.0, 0.0, 0])
        a[0] += b; a[1] += t; a[2] += 1
    d = {"source": f"{src} (ncu --set full --clock-control none, scripts/profile_round.sh): dram__bytes_read.sum + dram__bytes_write.sum "
                   "per launch, workload c2_beam5_msvd_bf16 B=1024",
         "classes": {c: {"dram_bytes_per_launch": a[0] / a[2], "ncu_duration_us": a[1] / a[2], "launches_captured": a[2]}
                     for c, a in acc.items()}}
    json.dump(d, open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
