#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun), in two calls so that the bench line quotes
# the DRAM traffic of the SAME round's capture:
#   bash scripts/profile_round2.sh capture   launch list + ncu --set full captures  (then, locally:
#                                            scripts/make_profile_summary.py -> profiles/r02_ncu_summary.md, traffic.json)
#   bash scripts/profile_round2.sh bench     bench lines (N=1 default, laser variant, reference arm), smoke(), pytest -m gpu
# Every command is bounded by its own timeout.  A number printed under ncu is never used as a bench value.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T="timeout -k 5"
B="python bench.py --no-cpu-baseline --no-e2e --no-api-e2e"
if [ "$1" = "capture" ]; then
  $T 200 $B > gpurun_out/r02_bench_plain.log 2>&1 && \
    $T 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv \
        $B > gpurun_out/r02_ncu_l.log 2>&1
  $T 300 ncu --set full --clock-control none --import-source on -k regex:"thermal_kernel|rates_dense|tile_pairop" -c 3 \
    -o gpurun_out/r02_prof_thermal_rates -f $B --steps 2 > gpurun_out/r02_ncu_a.log 2>&1
  $T 400 ncu --set full --clock-control none --import-source on \
    -k regex:"rates_refresh|dirty_scan|sweep_stream|sweep_pick|sweep_apply|sweep_plane_reduce|sweep_finalize" -s 14 -c 14 \
    -o gpurun_out/r02_prof_sweep -f $B --steps 2 > gpurun_out/r02_ncu_b.log 2>&1
  # the measured alternatives, same workload: pair-compacting refresh (262144), dense kernel without the pair-count
  # sort (131072), dense gather kernel on the compact state (65536), shared-memory tile kernel for refresh + rebuild
  # staged by TMA (32), gather refresh + rebuild of the first design (2)
  for f in 262144 131072 65536 32 2; do
    $T 200 $B --debug-flags $f > gpurun_out/r02_bench_flags$f.json 2> gpurun_out/r02_bench_flags$f.err
  done
  $T 300 ncu --set full --clock-control none --import-source on -k regex:"rates_tile3d" -s 2 -c 1 \
    -o gpurun_out/r02_prof_tile3d_tma -f $B --steps 2 --debug-flags 32 > gpurun_out/r02_ncu_c.log 2>&1
else
  $T 400 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
  $T 300 python bench.py --thermal laser --no-cpu-baseline > gpurun_out/r02_bench_laser.json 2> gpurun_out/r02_bench_laser.err
  $T 400 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
  $T 120 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1
  $T 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1
  tail -n 2 gpurun_out/r02_smoke.log gpurun_out/r02_pytest_gpu.log
fi
