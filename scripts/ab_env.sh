#!/bin/bash
# A/B of an environment switch on the default bench workload: scripts/ab_env.sh VAR [values...] (alternating runs)
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py ${WORKLOAD:+--workload $WORKLOAD} --no-cpu-baseline --no-sweep > gpurun_out/ab_${var}_$v.json 2> gpurun_out/ab_${var}_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${var}_$v.json").read().strip().splitlines()[-1])
print("$var=$v", round(d["value"]), round(d["ms_per_step"],3), "sustained", round(d["sustained"]["value"]), "e2e", round(d["e2e"]["value"]), {k:round(x["ms_per_step"],3) for k,x in d["breakdown"].items() if k.startswith("dec_") or k.startswith("attn_s")})
PY
done
