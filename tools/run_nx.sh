# usage: run_nx.sh N : the default bench line at N GPUs (our arm), as the driver launches it
N=$1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
tail -c 300 gpurun_out/n${N}_bench.json
