#!/bin/bash
# call 2: source-level ncu captures of the sparse sweep kernels (stream / pick / apply / refresh), one launch each
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T="timeout -k 5"
B="python bench.py --no-cpu-baseline --no-e2e --no-api-e2e"
$T 400 ncu --set full --clock-control none --import-source on \
    -k regex:"dirty_eval_compact|dirty_scan|sweep_stream|sweep_pick|sweep_apply" -s 15 -c 5 \
    -o gpurun_out/c2_prof_sweep -f $B --steps 2 > gpurun_out/c2_ncu.log 2>&1
tail -n 3 gpurun_out/c2_ncu.log
