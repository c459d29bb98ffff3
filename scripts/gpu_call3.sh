cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 0 5 6; do timeout 60 ./scripts/tma_probe $v; echo "variant $v rc=$?"; done > gpurun_out/r3_tma_probe.log 2>&1
timeout 900 python -m pytest tests/test_gpu_sweep.py -q --timeout 180 -x > gpurun_out/r3_sweep_tests.log 2>&1
echo "sweep tests rc=$?" >> gpurun_out/r3_sweep_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r3_bench_fused.json 2> gpurun_out/r3_bench_fused.err
