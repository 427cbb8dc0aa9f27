#!/bin/bash
# ncu --set full capture of the fused Jaccard + Adamic-Adar hub launch at the bench workload (run after the same command
# exited 0 without ncu). Usage: bash profiles/tools/ncu_fused_hub.sh  -> gpurun_out/prof_fused_hub_s24.ncu-rep
set -e
CMD="python bench.py --no-e2e --no-cpu-baseline --no-approx-er --steps 1 --warmup 1"
$CMD > gpurun_out/plain_fused.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:cta_owner_kernel<\(int\)2' --launch-skip 0 --launch-count 1 \
    -o gpurun_out/prof_fused_hub_s24 -f $CMD > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log
