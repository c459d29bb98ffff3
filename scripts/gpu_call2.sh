cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_gpu_sweep.py::test_sweep_is_deterministic_and_consistent[64-0]" -x -q --timeout 500 > gpurun_out/r2_sanitizer.log 2>&1
echo "rc=$?" >> gpurun_out/r2_sanitizer.log
