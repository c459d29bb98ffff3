cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sweep.py -q --timeout 180 -x -k "variants or rebuild" > gpurun_out/r5_sweep_tests.log 2>&1
echo "sweep tests rc=$?" >> gpurun_out/r5_sweep_tests.log
for f in 16 20; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --debug-flags $f > gpurun_out/r5_bench_f$f.json 2> gpurun_out/r5_bench_f$f.err
done
