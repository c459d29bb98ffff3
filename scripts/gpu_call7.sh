cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_sweep.py -m gpu -q --timeout 900 -k level3 -s > gpurun_out/r7_level3.log 2>&1
echo "rc=$?" >> gpurun_out/r7_level3.log
