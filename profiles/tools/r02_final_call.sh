# The 1-GPU gpurun command behind the final evidence: all GPU tests, then the default bench line.
# (The call of record also A/B-ed a "heaviest hub items first" ordering, GSP_ITEM_HEAVY_FIRST=0/1, with this script's
# quick bench and profiles/tools/emulated_rank_pass.py: 182.4 -> 186.1 ms on one GPU, 23.8 -> 24.2 ms on one rank of
# eight; the variant was removed — DESIGN.md section 7 — and the default line re-measured without it.)
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo rc=$?
