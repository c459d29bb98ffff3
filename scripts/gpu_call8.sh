cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err
timeout 600 python bench.py --thermal laser --no-cpu-baseline > gpurun_out/r8_bench_laser.json 2> gpurun_out/r8_bench_laser.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r8_bench_ref.json 2> gpurun_out/r8_bench_ref.err
