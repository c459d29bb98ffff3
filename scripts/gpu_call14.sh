cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in 65552; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-api-e2e --debug-flags $f > gpurun_out/r14_bench_f$f.json 2> gpurun_out/r14_bench_f$f.err
done
