#!/bin/bash
# N-GPU session: NCCL tests (class shard, captured all-gather), then bench.py at N (main line cfg2 + cfg3 / cfg4 sub-lines)
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/r2_multi_n$N.log
nvidia-smi -L > $L
echo "== nccl test" >> $L
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q >> $L 2>&1
echo "rc=$?" >> $L
echo "== bench N=$N" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.json 2>> $L
echo "rc=$?" >> $L
tail -25 $L | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_n$N.json"))
    print("N=$N cfg2 value %.0f ms %.4f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["e2e"].get("h2d_GBps_per_rank"), d["e2e"].get("h2d_only_GBps_per_rank"))
    for k,v in d["configs"].items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","launch","allgather_ms","class_shard_check","error","batch_per_gpu")})
except Exception as e: print("no json", e)
PY
