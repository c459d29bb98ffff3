#!/bin/bash
# Round measurement pass on one B200 (run under gpurun): bench lines, launch list, ncu captures.
# Every command is bounded by its own timeout.
T="timeout -k 5"
$T 200 python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
$T 200 python bench.py --impl reference > gpurun_out/r01_bench_reference.json 2>> gpurun_out/r01_bench.err
$T 120 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_plain.log 2>&1 && \
  $T 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bench.csv \
      python bench.py --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1
$T 200 ncu --set full --clock-control none --import-source on -k regex:"thermal_kernel|rates_tile" -c 2 \
  -o gpurun_out/r01_prof_thermal_rates -f python bench.py --steps 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_a.log 2>&1
$T 300 ncu --set full --clock-control none --import-source on -k regex:"dirty_eval|dirty_scan|sweep_stream|sweep_pick|sweep_apply" -s 10 -c 10 \
  -o gpurun_out/r01_prof_sweep -f python bench.py --steps 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b.log 2>&1
$T 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
$T 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -n 2 gpurun_out/smoke.log gpurun_out/pytest_gpu.log
