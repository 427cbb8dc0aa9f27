if [ "$1" == "pytest" ]; then timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_parity.py -m gpu -x -q -k "world2 or owner" 2>&1 | tail -15; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline --no-approx-er --no-e2e > gpurun_out/ab_cost.json 2> gpurun_out/ab_cost.err || tail -5 gpurun_out/ab_cost.err
python - <<'PY'
import json, sys
t = open("gpurun_out/ab_cost.json").read()
d = json.loads([l for l in t.splitlines() if l.startswith('{"')][-1])
print(round(d["ms_per_step"], 2), d["roofline"]["rank_spread"]["ms_per_rank"], {k: round(v["ms"], 2) for k, v in d["roofline"]["per_method"].items()})
PY
