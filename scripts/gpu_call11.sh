cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 600 > gpurun_out/r11_multi_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r11_multi_tests.log
timeout 600 python -m pytest tests/test_gpu_grains.py tests/test_gpu_campaign.py -q --timeout 600 > gpurun_out/r11_grains_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r11_grains_tests.log
