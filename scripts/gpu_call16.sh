cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 600 > gpurun_out/r16_multi_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r16_multi_tests.log
P=29911
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 > gpurun_out/r16_strong512_2.json 2> gpurun_out/r16_strong512_2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --scaling weak > gpurun_out/r16_weak512_2.json 2> gpurun_out/r16_weak512_2.err
