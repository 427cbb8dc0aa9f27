set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo rc=$?
for cfg in "16384,1024,768,2" "32768,1024,1024,1" "16384,1024,1024,1"; do
  GSP_HUB_CLASS=$cfg python bench.py --no-cpu-baseline --no-approx-er --no-e2e --steps 4 --warmup 3 > gpurun_out/sw.json 2> gpurun_out/sw.err || { tail -3 gpurun_out/sw.err; continue; }
  python - "$cfg" <<'PY'
import json, sys
d = json.load(open("gpurun_out/sw.json"))
print(sys.argv[1], round(d["ms_per_step"], 2), {k: round(v["ms"], 2) for k, v in d["roofline"]["per_method"].items()})
PY
done
