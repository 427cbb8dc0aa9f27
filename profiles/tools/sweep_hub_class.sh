for cfg in "16384,1024,640,2" "16384,1024,512,2" "16384,1024,768,2" "32768,1024,1024,1" "8192,1024,512,3" "8192,1024,640,2" "32768,1024,768,1"; do
  GSP_HUB_CLASS=$cfg python bench.py --scale 22 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/sw.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/sw.log") if x.startswith("{")]
d=json.loads(l[-1]); print("$cfg", round(d["ms_per_step"],2), {k:round(v["ms"],2) for k,v in d["per_method"].items() if k in ("jaccard","adamic_adar")})
PY
done
