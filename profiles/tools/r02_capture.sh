#!/bin/bash
# Round-2 ncu evidence at the bench workload (R-MAT scale 24), one gpurun call:
#   1. launch list with device times of two bench steps            -> gpurun_out/r02_launches.csv
#   2. --set full of the fused Jaccard + Adamic-Adar hub launch     -> gpurun_out/r02_fused_hub.ncu-rep
#   3. --set full of the select kernels (histogram, tally, emit)    -> gpurun_out/r02_select.ncu-rep
# Each ncu run repeats a command that has just exited 0 without ncu.
set -e
mkdir -p gpurun_out
CMD="python bench.py --no-e2e --no-cpu-baseline --no-approx-er --no-small-configs --steps 2 --warmup 1"
$CMD > gpurun_out/r02_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:cta_owner_kernel<\(int\)2' --launch-skip 0 --launch-count 1 \
    -o gpurun_out/r02_fused_hub -f $CMD > gpurun_out/r02_ncu_hub.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:histogram_kernel|fused_tally_kernel|fused_emit_kernel' --launch-skip 0 --launch-count 9 \
    -o gpurun_out/r02_select -f $CMD > gpurun_out/r02_ncu_select.log 2>&1
tail -2 gpurun_out/r02_ncu_select.log
