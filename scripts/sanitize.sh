#!/bin/bash
# compute-sanitizer over small invocations of the path; logs -> gpurun_out/sanitizer_<tool>_<case>.log
# usage: scripts/sanitize.sh [tools...]   (default: memcheck racecheck synccheck)
tools=${@:-memcheck racecheck synccheck}
mkdir -p gpurun_out
for tool in $tools; do
  for case in tiny b300; do
    log=gpurun_out/sanitizer_${tool}_${case}.log
    t0=$(date +%s)
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python scripts/sanitize_case.py $case > $log 2>&1
    rc=$?
    echo "== $tool $case rc=$rc $(( $(date +%s) - t0 ))s: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
  done
done
