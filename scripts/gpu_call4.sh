cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep.py -q --timeout 180 -x > gpurun_out/r4_sweep_tests.log 2>&1
echo "sweep tests rc=$?" >> gpurun_out/r4_sweep_tests.log
for f in 0 4 8 12 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --debug-flags $f > gpurun_out/r4_bench_f$f.json 2> gpurun_out/r4_bench_f$f.err
done
