#!/bin/bash
# usage: scripts/bench_classes.sh <tag> [bench args...]  -- runs bench.py and prints the per-class device times
tag=$1; shift
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { echo "bench $tag failed"; tail -5 gpurun_out/bench_$tag.err; }
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{tag}.json").read().strip().splitlines()[-1])
except Exception as e:
    print(tag, "no json", e); sys.exit(0)
print(f"[{tag}] value={d['value']:.0f} ms/step={d['ms_per_step']:.3f} e2e={d['e2e']['value']:.0f} roof={d['roofline']['kernel']}:{d['roofline']['frac']:.3f}")
print("   " + " ".join(f"{k}={v['ms_per_step']:.3f}" for k, v in d["breakdown"].items()))
PY
