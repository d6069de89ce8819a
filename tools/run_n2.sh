# two GPUs: the real-GPU multi-rank parity test, then the default bench line at N = 2
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/n2_tests.log 2>&1; tail -2 gpurun_out/n2_tests.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
tail -c 300 gpurun_out/n2_bench.json
