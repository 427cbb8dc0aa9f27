for ord in owner window owner window; do
  GSP_ITEM_ORDER=$ord python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-approx-er > gpurun_out/ab.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/ab.log") if x.startswith("{")]
d=json.loads(l[-1]); print("$ord", round(d["ms_per_step"],2), {k:round(v["ms"],2) for k,v in d["per_method"].items() if k in ("jaccard","adamic_adar")})
PY
done
