for mode in smi nvml smi nvml; do
GSP_BENCH_CLOCKS=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline --no-approx-er --no-e2e > gpurun_out/ab_clk.json 2> gpurun_out/ab_clk.err || tail -5 gpurun_out/ab_clk.err
python - "$mode" <<'PY'
import json, sys
t = open("gpurun_out/ab_clk.json").read()
d = json.loads([l for l in t.splitlines() if l.startswith('{"')][-1])
print(sys.argv[1], round(d["ms_per_step"], 2), d["roofline"]["rank_spread"]["ms_per_rank"], d["clocks"])
PY
done
