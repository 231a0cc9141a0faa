#!/bin/bash
# ncu evidence of the current build for profiles/: (1) launch list of two generate() calls of the benchmark workload,
# (2) one --set full capture of every kernel of the second call's encoder and first two decode steps.
#   scripts/profile_round.sh <tag>        (under gpurun; outputs in gpurun_out/)
tag=${1:-rX}
wl=c2_beam5_msvd_bf16
mkdir -p gpurun_out
python scripts/ncu_case.py $wl 2 > gpurun_out/ncu_case_$tag.log 2>&1 || { tail -5 gpurun_out/ncu_case_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$tag.csv \
    python scripts/ncu_case.py $wl 2 > gpurun_out/ncu_launches_$tag.log 2>&1
# index of the second call's first library kernel (the feature projection: the tf32 persistent GEMM)
skip=$(python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/launches_$tag.csv")) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iid = hdr.index("ID")
hits = [int(r[iid]) for r in rows[1:] if "0, 0, 1, 1, 0>" in r[ik] and "gemm_tc_persistent_kernel<4" in r[ik]]
print(hits[1] if len(hits) > 1 else 0)
PY
)
echo "second call starts at launch $skip"
ncu --set full --clock-control none --launch-skip $skip -c 22 -f -o gpurun_out/prof_$tag \
    python scripts/ncu_case.py $wl 2 > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep
# summaries are made here (the report itself can exceed what gpurun_out/ carries back)
python profiles/summarize.py launches gpurun_out/launches_$tag.csv gpurun_out/${tag}_launches.md
python profiles/summarize.py raw gpurun_out/prof_$tag.ncu-rep gpurun_out/${tag}_kernels.md
python profiles/summarize.py traffic gpurun_out/prof_$tag.ncu-rep gpurun_out/${tag}_traffic.json
if [ $(stat -c %s gpurun_out/prof_$tag.ncu-rep) -gt 45000000 ]; then rm gpurun_out/prof_$tag.ncu-rep; echo "(report removed: too large to carry back)"; fi
