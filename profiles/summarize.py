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


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
