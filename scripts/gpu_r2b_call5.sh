#!/bin/bash
# call 5: ncu capture of the class-sorted refresh kernel and the apply kernel after the counter changes
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T="timeout -k 5"
B="python bench.py --no-cpu-baseline --no-e2e --no-api-e2e"
$T 400 ncu --set full --clock-control none --import-source on \
    -k regex:"rates_refresh|sweep_apply" -s 6 -c 2 \
    -o gpurun_out/c5_prof -f $B --steps 2 > gpurun_out/c5_ncu.log 2>&1
tail -n 3 gpurun_out/c5_ncu.log | cut -c1-300
