#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi.log
echo "== 2-GPU nccl test" >> gpurun_out/multi.log
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q >> gpurun_out/multi.log 2>&1
echo "rc=$?" >> gpurun_out/multi.log
for n in 1 2; do
  echo "== bench N=$n" >> gpurun_out/multi.log
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2>> gpurun_out/multi.log
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n$n.json 2>> gpurun_out/multi.log
  fi
  echo "rc=$?" >> gpurun_out/multi.log
done
echo "== reference arm" >> gpurun_out/multi.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/multi.log
tail -30 gpurun_out/multi.log
