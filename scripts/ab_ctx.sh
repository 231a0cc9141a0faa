for v in 0 1 2 0 1 2; do
  VC_CTX_PERSISTENT=$v python bench.py --no-cpu-baseline --no-sweep > gpurun_out/s3_ctx$v.json 2> gpurun_out/s3_ctx$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s3_ctx$v.json").read().strip().splitlines()[-1])
print("ctx_persistent=$v", round(d["value"]), round(d["ms_per_step"],3), round(d["sustained"]["value"]), {k:round(x["ms_per_step"],3) for k,x in d["breakdown"].items() if k in ("dec_context_proj","dec_vocab","dec_lstm","attn_step")})
PY
done
