#!/bin/bash
# A/B of library builds inside ONE gpurun call (box-to-box variance is ~5 %): variants are prebuilt copies
# gnn-sparsification-research_b200/csrc/_build/libgsp_<tag>.so; usage: bash profiles/tools/ab_libs.sh A B C [-- extra bench args]
P=gnn-sparsification-research_b200
cp $P/libgsp.so /tmp/libgsp_orig.so
tags=(); extra=()
while [ $# -gt 0 ]; do if [ "$1" == "--" ]; then shift; extra=("$@"); break; fi; tags+=("$1"); shift; done
for round in 1 2; do
for t in "${tags[@]}"; do
  cp $P/csrc/_build/libgsp_$t.so $P/libgsp.so
  python bench.py --no-cpu-baseline --no-approx-er --no-e2e --steps 4 --warmup 3 "${extra[@]}" > gpurun_out/ab_$t.json 2> gpurun_out/ab_$t.err || { echo "$t failed"; tail -3 gpurun_out/ab_$t.err; continue; }
  python - "$t" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
print(sys.argv[1], round(d["ms_per_step"], 2), {k: round(v["ms"], 2) for k, v in d["roofline"]["per_method"].items()})
PY
done
done
cp /tmp/libgsp_orig.so $P/libgsp.so
