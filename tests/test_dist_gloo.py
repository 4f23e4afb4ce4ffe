"""CPU, world_size 2, gloo: the multi-process plumbing of the sharded head (batch shards, class-feature all-gather,
max-vertices all-reduce) -- the N > 1 path of SURVEY.md section 8e without GPUs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, K, D):
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(here, "schemanet-pytorch_b200"))
    from schemanet_b200 import dist as shdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(K * D, dtype=torch.float32).reshape(K, D)
        per = (K + world - 1) // world
        local = torch.zeros(per, D)
        lo, hi = min(rank * per, K), min((rank + 1) * per, K)
        local[:hi - lo] = full[lo:hi]
        got = shdist.gather_class_features(local, K)
        assert torch.equal(got, full)
        mv = torch.tensor([100 + 7 * rank], dtype=torch.int32)
        shdist.global_max_vertices(mv)
        assert int(mv) == 100 + 7 * (world - 1)
        batch = torch.arange(11)
        mine = shdist.shard_batch(batch, rank, world)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine.numel()]))
        assert sum(int(s) for s in sizes) == 11
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 5, 3), nprocs=2, join=True)
