cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --debug-flags 16 > gpurun_out/r6_bench_f16.json 2> gpurun_out/r6_bench_f16.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rates_tile3d -c 3 -o gpurun_out/r6_tile3d -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --debug-flags 16 > gpurun_out/r6_ncu.log 2>&1
