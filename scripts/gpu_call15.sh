cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dirty_eval_compact" -s 4 -c 2 -o gpurun_out/r15_refresh -f python bench.py --no-cpu-baseline --no-e2e --no-api-e2e --steps 2 > gpurun_out/r15_ncu.log 2>&1
