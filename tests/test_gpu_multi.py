"""GPU, world_size 2, NCCL: class-sharded head (one all-gather of the class embeddings) and
batch sharding.  Needs 2 GPUs: skipped on a single-GPU box (the gloo version of the plumbing runs on CPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "schemanet-pytorch_b200"), os.path.join(root, "oracle"), os.path.join(root, "tests")):
        sys.path.insert(0, p)
    import head_oracle as ho
    from schemanet_b200.head import SchemaHead
    from schemanet_b200 import dist as shdist
    from test_gpu_parity import build_modules
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        B, d, M, K, Vc, D = 10, 64, 128, 7, 128, 256          # K = 7 does not divide by 2: ragged class shards
        vocab, mid, attn, attn_cls = ho.synth_inputs(B, d, M, seed=900)
        schema = ho.synth_schema(M, K, Vc, seed=901)
        gnn = ho.synth_gnn(M, D, seed=902)
        sn, m = build_modules(schema=schema, gnn=gnn, M=M, K=K, Vc=Vc, D=D, dev=dev)
        lo, hi = shdist.shard_range(B, rank, world)          # batch shard of this rank
        mid_s, attn_s, cls_s = mid[:, lo:hi].contiguous().to(dev), attn[lo:hi].to(dev), attn_cls[lo:hi].to(dev)
        # every check is recorded and reduced over the ranks before anything is raised: a rank that raised alone would leave its
        # peer waiting in the next collective
        problems = []

        def rel(a, b):
            return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

        def mark(msg):
            print(f"[rank {rank}] {msg}", file=sys.stderr, flush=True)

        dist.barrier()                                       # (communicator up before the first collective on a side stream)
        full = SchemaHead(vocab.to(dev), sn, m)(mid_s, attn_s, cls_s)
        torch.cuda.synchronize(); mark("full head done")
        shard = SchemaHead(vocab.to(dev), sn, m, class_shard=(rank, world))(mid_s, attn_s, cls_s)
        # (not bit-identical: the GEMM operand scales are powers of two taken from the maxima of the graphs a call processes, and
        # a different grouping of the class graphs moves which tiny entries fall into fp16 subnormals -- ~1e-7 relative)
        if rel(shard["feat_class"], full["feat_class"]) > 1e-6:
            problems.append(f"class embeddings differ under class sharding: {rel(shard['feat_class'], full['feat_class']):.2e}")
        if rel(shard["pred"], full["pred"]) > 1e-6:
            problems.append(f"logits differ under class sharding: {rel(shard['pred'], full['pred']):.2e}")
        torch.cuda.synchronize(); mark("sharded head done")
        ref = ho.head_forward(mid[:, lo:hi].contiguous(), attn[lo:hi], attn_cls[lo:hi], vocab, schema, gnn, ho.HEAD_CFG)
        err = (shard["pred"].cpu() - ref["pred"]).abs().max() / ref["pred"].abs().max()
        if not err < 1e-5:
            problems.append(f"rank {rank}: logits rel err {err:.2e}")
        # the class-sharded step as ONE CUDA graph (the NCCL all-gather is captured on the class-side stream); replay == eager
        from schemanet_b200.head import GraphedHead
        head_g = SchemaHead(vocab.to(dev), sn, m, class_shard=(rank, world))
        gh = GraphedHead(head_g, mid_s.clone(), attn_s.clone(), cls_s.clone())
        mark("graph captured")
        for _ in range(3):
            out_g = gh.replay()
        torch.cuda.synchronize()
        if not torch.equal(out_g["pred"], shard["pred"]):
            problems.append("graph replay of the class-sharded head differs from the eager result")
        del gh, out_g, head_g            # a live CUDA graph holds NCCL kernels: release it before the communicator goes away
        torch.cuda.synchronize()
        mv = shard["graphs"].max_vertices.clone()
        shdist.global_max_vertices(mv)
        sizes = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(world)]
        dist.all_gather(sizes, shard["graphs"].max_vertices)
        if int(mv) != max(int(s) for s in sizes):
            problems.append("global_max_vertices")
        mark(f"checks done: {problems}")
        bad = torch.tensor([len(problems)], device=dev)
        dist.all_reduce(bad)
        assert int(bad) == 0, f"rank {rank}: {problems or 'a peer rank failed'}"
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_class_sharded_head_world2_nccl():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
