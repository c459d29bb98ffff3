cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r5_bench_g32.json 2> gpurun_out/r5_bench_g32.err
CETKMC_L2_FETCH_DEFAULT=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r5_bench_g64.json 2> gpurun_out/r5_bench_g64.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --debug-flags 2 > gpurun_out/r5_bench_g32_gather.json 2> gpurun_out/r5_bench_g32_gather.err
