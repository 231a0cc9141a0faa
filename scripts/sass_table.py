#!/usr/bin/env python
"""profiles/<tag>_sass_opcodes.md: per-kernel counts of the SASS mnemonics that show which hardware paths the in-tree library
uses (B200_PROFILING.md: UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld, UBLKCP = bulk copy,
HMMA = mma.sync, MUFU.TANH, LDGSTS = cp.async, REDUX = warp reduce).   python scripts/sass_table.py profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "video-captioning_b200", "libvc_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "HMMA", "MUFU.TANH", "MUFU.EX2",
       "LDGSTS", "REDUX", "FFMA", "HFMA2", "LDSM"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main(dst):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["_total"] += 1
        for o in OPS:
            if o == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    cur[o] += 1
            elif op == o or op.startswith(o + "."):
                cur[o] += 1
    names = demangle(list(counts))
    def short(n):
        n = names.get(n, n)
        n = re.sub(r"\(.*", "", n).replace("void ", "").replace("vc::tc::", "").replace("vc::", "")
        return n[:88]
    rows = [(short(k), v) for k, v in counts.items() if v["_total"] > 0]
    rows.sort(key=lambda kv: -(kv[1]["UTCHMMA"] * 1000 + kv[1]["HMMA"] * 10 + kv[1]["MUFU.TANH"] + kv[1]["UBLKCP"]))
    tot = collections.Counter()
    for _, v in rows:
        tot.update(v)
    with open(dst, "w") as f:
        f.write("# SASS opcode table of `video-captioning_b200/libvc_b200.so` (`cuobjdump -sass`, sm_100a)\n\n"
                "Static instruction counts per kernel (`python scripts/sass_table.py`).  UTCHMMA = `tcgen05.mma` (`.2CTA`: CTA pairs), "
                "UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = `cp.async.bulk`, LDTM = `tcgen05.ld`, UTCBAR = `tcgen05.commit`, "
                "SYNCS = mbarrier ops, HMMA = `mma.sync`, LDSM = `ldmatrix`, LDGSTS = `cp.async`.\n\n")
        f.write("| kernel | SASS | " + " | ".join(OPS) + " |\n|---|---:|" + "---:|" * len(OPS) + "\n")
        f.write("| **all kernels (%d)** | %d | " % (len(rows), tot["_total"]) + " | ".join(str(tot[o]) for o in OPS) + " |\n")
        for n, v in rows:
            if sum(v[o] for o in OPS if o != "FFMA") == 0 and v["FFMA"] < 64:
                continue
            f.write("| `%s` | %d | " % (n, v["_total"]) + " | ".join(str(v[o]) if v[o] else "" for o in OPS) + " |\n")


if __name__ == "__main__":
    main(sys.argv[1])
