cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sweep.py -q --timeout 400 -x -k "not level3" > gpurun_out/r5_sweep_tests.log 2>&1
echo "sweep tests rc=$?" >> gpurun_out/r5_sweep_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-api-e2e > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err
