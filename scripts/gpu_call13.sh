cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sweep_stream|sweep_pick|sweep_apply|dirty_scan|dirty_eval_compact' -c 10 --launch-skip 30 -o gpurun_out/r13_sweep -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-api-e2e > gpurun_out/r13_ncu.log 2>&1
