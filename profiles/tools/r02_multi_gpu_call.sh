#!/bin/bash
# usage (inside gpurun --gpus N): bash profiles/tools/r02_multi_gpu_call.sh N [pytest]
N=$1
if [ "$2" == "pytest" ]; then
  timeout 1200 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_parity.py -m gpu -x -q -k "world2 or sharded or select or emulated or owner" 2>&1 | tail -5
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
  > gpurun_out/r02_bench_s24_n$N.json 2> gpurun_out/r02_bench_s24_n$N.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_s24_n$N.json | head -c 400; tail -3 gpurun_out/r02_bench_s24_n$N.err
python - "$N" <<'PY'
import json, sys
t = open(f"gpurun_out/r02_bench_s24_n{sys.argv[1]}.json").read()
d = json.loads([l for l in t.splitlines() if l.startswith('{"')][-1])
r = d["roofline"]
print("ms_per_step", round(d["ms_per_step"], 2), "e2e", d["e2e"].get("ms_per_step"), "spread", r.get("rank_spread", {}).get("ms_per_rank"))
print({k: round(v["ms"], 2) for k, v in r["per_method"].items()})
print("approx_er", r.get("approx_er", {}).get("sparsify_ms"))
PY
