#!/bin/bash
# usage: tools/gpu_multi_n.sh N   -- bench.py on N GPUs of one box (torchrun, NCCL), JSON line -> gpurun_out/bench_nN.json
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2>> gpurun_out/multi_n$N.log
echo "rc=$?" >> gpurun_out/multi_n$N.log
tail -5 gpurun_out/multi_n$N.log
