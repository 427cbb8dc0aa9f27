set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for hf in 0 1; do
  GSP_ITEM_HEAVY_FIRST=$hf python bench.py --no-cpu-baseline --no-approx-er --no-e2e --steps 4 --warmup 3 > gpurun_out/sw.json 2> gpurun_out/sw.err || { tail -3 gpurun_out/sw.err; continue; }
  python - "$hf" <<'PY'
import json, sys
d = json.load(open("gpurun_out/sw.json"))
print("heavy_first", sys.argv[1], round(d["ms_per_step"], 2), {k: round(v["ms"], 2) for k, v in d["roofline"]["per_method"].items()})
PY
done
PYTHONPATH=. GSP_ITEM_HEAVY_FIRST=0 timeout 300 python profiles/tools/emulated_rank_pass.py 8 3 2>&1 | tail -1
PYTHONPATH=. timeout 300 python profiles/tools/emulated_rank_pass.py 8 3 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo rc=$?
