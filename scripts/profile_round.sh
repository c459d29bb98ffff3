set -x
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
python bench.py --impl reference > gpurun_out/r01_bench_reference.json 2>> gpurun_out/r01_bench.err
python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"thermal_kernel|rates_tile" -c 2 -o gpurun_out/r01_prof_thermal_rates -f python bench.py --steps 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"dirty_eval|dirty_scan|sweep_stream|sweep_pick|sweep_apply" -s 10 -c 10 -o gpurun_out/r01_prof_sweep -f python bench.py --steps 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b.log 2>&1
python scripts/time_run_kmc.py > gpurun_out/time_run_kmc.log 2>&1
python scripts/probe_perf.py 512 > gpurun_out/probe_final.log 2>&1
tail -n 3 gpurun_out/time_run_kmc.log
