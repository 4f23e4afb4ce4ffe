#!/usr/bin/env python
"""bench.py -- schema-head images/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA head
    python bench.py --impl reference --gpus N ...            # the reference's CPU head on the box's host cores

A "step" is one pass of the whole head (discretize -> instance graphs -> class atlas -> class-side GNN ->
instance-side GNN -> logits) over one batch of synthetic tensors.  The line's own `value` is the DeiT-Small / CIFAR-100
shape (BASELINE.json configs[1]; B=256 per GPU, d=384, M=1024, K=100, Vc=1024, D=256), class side recomputed every step as
the reference does; N > 1: one process per GPU, the batch is sharded (weak scaling: 256 images per GPU), no data-path
collective.  The same line carries, under "configs", the two multi-GPU configurations BASELINE.json
names, measured in the same run at the same N:
    cfg3  DeiT-Base / Caltech-101: the 512-image batch SHARDED over the N GPUs (strong scaling) + classes sharded
    cfg4  DeiT-Base / ImageNet-1k: 1024 images per GPU, the 1000 class schemas sharded by class, one NCCL all-gather of
          the [K, D] class embeddings per step (captured inside the CUDA graph, overlapping the instance side)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "schemanet-pytorch_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOAD = "cfg2"
L = 196
N_INPUT_SETS = 3      # distinct input batches rotated between steps (plus the whole class edge tensor streamed per step)
TARGET_FRAC = 0.70    # north star: every stage at >= 70 % of its HBM or tensor roofline
PARITY_BAR = 1e-5


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, from the newest committed `ncu --set full` capture
    (profiles/rNN_ncu_traffic.json, written by tools/ncu_summary.py traffic <.ncu-rep>)."""
    prof = os.path.join(ROOT, "profiles")
    try:
        files = sorted(f for f in os.listdir(prof) if f.endswith("_ncu_traffic.json"))
        if not files:
            return None, None
        with open(os.path.join(prof, files[-1])) as f:
            return json.load(f), "profiles/" + files[-1]
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.thread, self.gpu = [], None, None, str(gpu_index)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", self.gpu], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(local_rank):
    """Pin this process (and therefore the pinned host buffers it first-touches) to the CPUs nearest its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus near gpu {local_rank}"
    except Exception as e:
        return f"not bound ({type(e).__name__})"
    return "not bound"


def make_problem(cfg, seed, device):
    """Seeded synthetic tensors of SURVEY.md section 8d, built on the CPU (identical for the GPU and CPU arms)."""
    import head_oracle as ho
    c = ho.CONFIGS[cfg]
    vocab = None
    sets = []
    for i in range(N_INPUT_SETS):
        vocab, mid, attn, attn_cls = ho.synth_inputs(c["B"], c["d"], c["M"], seed + 10 * i, L, "easy", vocab)
        sets.append((mid, attn, attn_cls))
    schema = ho.synth_schema(c["M"], c["K"], c["Vc"], seed + 1)
    gnn = ho.synth_gnn(c["M"], c["D"], seed + 2)
    return c, vocab, sets, schema, gnn


def build_head(c, vocab, schema, gnn, device, class_shard=None):
    from schema_inference.graph import SchemaNet, Matcher
    from schemanet_b200.head import SchemaHead
    import head_oracle as ho
    sn = SchemaNet(c["M"], c["K"], class_max_vertices=c["Vc"], clamp_vertex_attn=ho.HEAD_CFG["clamp_vertex_attn"],
                   clamp_edge_attn=ho.HEAD_CFG["clamp_edge_attn"], prune_node_threshold=ho.HEAD_CFG["prune_node_threshold"])
    sn.vertex_weights.copy_(schema["vertex_weights"])
    sn.edge_weights.copy_(schema["edge_weights"])
    sn.vertex_attribute_weights.copy_(schema["w_v"])
    sn.edge_attribute_weights.copy_(schema["w_e"])
    sn.register_class_vertices(schema["class_ingredients"])
    m = Matcher("inner_product", c["M"], dict(embed_dim=c["D"], num_layers=2, identity_proj=False, activation="relu"))
    m.gnn.load_state_dict(gnn)
    sn.to(device).eval()
    m.to(device).eval()
    return SchemaHead(vocab.to(device), sn, m, class_shard=class_shard)


def cpu_head_time(c, vocab, sets, schema, gnn, iters, batch=None, budget_s=None):
    """The reference head on host cores: the reference's own C++ (oracle/_ref) for the native loops when it was built,
    the oracle's restatement (same ATen CPU ops the reference calls) for the Python parts.  Returns the last output too
    (bench.py's parity check).  budget_s: stop early once that much wall time has been spent (>= 1 step is always timed)."""
    import head_oracle as ho
    import build_ref
    ext = build_ref.load()
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    mid, attn, attn_cls = sets[0]
    B = batch or c["B"]
    mid, attn, attn_cls = mid[:, :B].contiguous(), attn[:B].contiguous(), attn_cls[:B].contiguous()
    times, out, t_start = [], None, time.perf_counter()
    for _ in range(iters):
        t0 = time.perf_counter()
        out = ho.head_forward(mid, attn, attn_cls, vocab, schema, gnn, ho.HEAD_CFG, ext=ext)
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_start > budget_s:
            break
    total = sum(times)
    times.sort()
    med = times[len(times) // 2]
    return {"ips": B / med, "median_s": med, "total_s": total, "steps": len(times), "cores": cores,
            "kind": "reference" if ext is not None else "port", "out": out}


def kind_detail(kind):
    return ("native loops: the reference's own C++ compiled in place (oracle/_ref); Python glue: restated op by op in "
            "oracle/head_oracle.py on the ATen CPU ops the reference calls (its classes need /root/reference, absent on the box)"
            if kind == "reference" else "oracle restatement (C + ATen CPU ops); the reference C++ was not built")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c, vocab, sets, schema, gnn = make_problem(WORKLOAD, 1234, "cpu")
    B = c["B"]      # bounded sample: the full 256-image batch per step (~1-3 s each); every requested step is timed
    t0 = time.perf_counter()
    n_warm = max(1, min(args.warmup, 2))
    for _ in range(n_warm):
        cpu_head_time(c, vocab, sets, schema, gnn, 1, batch=B)
    r = cpu_head_time(c, vocab, sets, schema, gnn, max(1, args.steps), batch=B, budget_s=150.0)
    sample = (f"full {B}-image {WORKLOAD} batch per step, K={c['K']} class side recomputed per step, "
              f"{n_warm} warm-up + {r['steps']} timed steps (of {args.steps} requested; 150 s budget); "
              f"{kind_detail(r['kind'])}; ATen ops with {r['cores']} threads; value = batch / median step")
    line = {"impl": "reference", "metric": "schema_head_images_per_sec", "value": r["ips"], "unit": "images/s",
            "n_gpus": args.gpus, "steps": r["steps"], "steps_requested": args.steps, "warmup": n_warm,
            "ms_per_step": r["total_s"] / r["steps"] * 1e3, "ms_per_step_median": r["median_s"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(c, args.gpus),
            "cpu_baseline": {"value": r["ips"], "unit": "images/s", "cores": r["cores"], "kind": r["kind"],
                             "kind_detail": kind_detail(r["kind"]), "sample": sample},
            "e2e": {"value": r["ips"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    emit(line)


WORKLOAD_NAMES = {"cfg1": "DeiT-Tiny SchemaNet head, CIFAR-10 shape (BASELINE.json configs[0])",
                  "cfg2": "DeiT-Small SchemaNet head, CIFAR-100 shape (BASELINE.json configs[1])",
                  "cfg3": "DeiT-Base SchemaNet head, Caltech-101 shape (BASELINE.json configs[2])",
                  "cfg4": "DeiT-Base SchemaNet head, ImageNet-1k shape (BASELINE.json configs[3])"}


def workload_config(c, n_gpus):
    return {"workload": WORKLOAD_NAMES[WORKLOAD], "batch_per_gpu": c["B"],
            "global_batch": c["B"] * n_gpus, "tokens": L, "d": c["d"], "vocab_M": c["M"], "classes_K": c["K"],
            "class_vertices_Vc": c["Vc"], "gnn_dim_D": c["D"], "parallelism": f"batch-shard dp{n_gpus}",
            "class_side": "recomputed every step (reference semantics), on a second CUDA stream overlapping the instance side",
            "launch": "device-resident arm: one CUDA graph replay per step (captured per resident input set; --no-graph launches each kernel); e2e arm: eager launches",
            "cache_policy": "3 rotating input sets + the whole class edge tensor streamed per step: larger than the "
                            "126 MB L2 (cfg2: 3 x 116 MB + 419 MB)"}


# ----------------------------------------------------------------------------------------------------------------
# algorithmic work and the work gemm3x_kernel actually visits
# ----------------------------------------------------------------------------------------------------------------
def stage_bytes_flops(c, n_bar):
    B, d, M, K, Vc = c["B"], c["d"], c["M"], c["K"], c["Vc"]
    return {
        "discretize": {"flops": 2.0 * L * B * d * M, "bytes": B * (L * d * 4 + L * 8) + M * d * 4},
        "graph_build": {"bytes": B * (L * L * 4 + L * 4 + L * 8 + 4 * n_bar * n_bar + 12 * n_bar + 8) + L * L * 4},
        # one read of the edge parameter; the normalised [K, Vc, Vc] tensor is not materialised on the hot path
        "atlas": {"bytes": 1.0 * K * Vc * Vc * 4 + 3.0 * K * Vc * 4},
    }


def adj_gemm_kblocks(sizes, rows_per_graph, unit, identity_tail, skip_past, n_tiles):
    """k-blocks (32 wide) of all work units one adjacency-GEMM launch visits: the TILE_LOOP rule of gemm3x_kernel
    (csrc/gnn_tc.cu) restated.  A unit = `unit` rows x 256 columns; units wholly past their graph's size are skipped when
    the epilogue allows it (skip_past) and there is no identity tail; with an identity tail they visit their own diagonal."""
    total = 0
    ub_per = -(-rows_per_graph // unit)
    kmax = -(-rows_per_graph // 32)
    for n_g in sizes:
        for ub in range(ub_per):
            past = ub * unit >= n_g
            if skip_past and not identity_tail and past:
                continue
            ka = 0 if (identity_tail and past) else max(1, -(-n_g // 32))
            k2 = 0
            if identity_tail and (ub + 1) * unit > n_g:
                k2s = max(ka, ub * (unit // 32))
                k2 = max(0, min((ub + 1) * (unit // 32), kmax) - k2s)
            total += (ka + k2) * n_tiles
    return total


def linear_gemm_kblocks(sizes, rows_per_graph, unit, D, skip_masked):
    """Same for the linear GEMM over the flattened [G * rows_per_graph, D] rows."""
    total = 0
    rows = len(sizes) * rows_per_graph
    for ub in range(-(-rows // unit)):
        r0 = ub * unit
        gi = r0 // rows_per_graph
        if skip_masked and (r0 + unit - 1) // rows_per_graph == gi and r0 - gi * rows_per_graph >= sizes[gi]:
            continue
        total += (D // 32) * (D // 256)
    return total


def visited_gemm_flops(c, n_act, n_inst, unit=256):
    """fp32 multiply-adds (x2) of the tiles the tensor-core GEMMs of one step visit, per family.  Mirrors run_layers_tc:
    embed_dim 256: class graphs reduced to their un-pruned vertices (skip_masked, no identity tail), layer 0 is one
    adjacency GEMM (the first Linear is applied to the (M+1)-row table); wider: every class vertex stays, identity tail."""
    D, Vc = c["D"], c["Vc"]
    nt = D // 256
    per_kb = unit * 32 * 256 * 2.0
    fused = D == 256
    layers = 2
    adj = lin = 0.0
    # class side
    adj += layers * adj_gemm_kblocks(n_act, Vc, unit, 0 if fused else 1, True, nt) * per_kb
    lin_layers = layers - 1 if fused else layers
    lin += lin_layers * linear_gemm_kblocks(n_act if fused else [Vc] * len(n_act), Vc, unit, D, fused) * per_kb
    # instance side (one unit per graph: rows_per_graph = 196 <= unit)
    adj += layers * adj_gemm_kblocks(n_inst, L, unit, 0, False, nt) * per_kb
    inst_fused = D == 256 and (c["M"] + 1) <= len(n_inst) * L
    lin += (layers - 1 if inst_fused else layers) * linear_gemm_kblocks(n_inst, L, unit, D, False) * per_kb
    return {"adjacency_gemm": adj, "linear_gemm": lin}


FAMILY = {"gnn_adj_gemm_tc": "adjacency_gemm", "gnn_adj_ln_tc": "adjacency_gemm", "gnn_adj_gemm": "adjacency_gemm",
          "gnn_linear_ln_tc": "linear_gemm", "gnn_linear_tc": "linear_gemm", "gnn_linear_gemm": "linear_gemm",
          "discretize_tc_kernel": "discretize", "discretize_tc_f16_kernel": "discretize",
          "discretize_exact_kernel": "discretize", "instance_graph_kernel": "graph_build",
          "class_edges_kernel": "atlas"}
STAGE_OF = {"rows_to_half_kernel": "1 discretize", "codebook_norms_kernel": "1 discretize", "row_sqnorm_kernel": "1 discretize",
            "discretize_tc_kernel": "1 discretize", "discretize_tc_f16_kernel": "1 discretize",
            "discretize_exact_kernel": "1 discretize", "discretize_recheck_kernel": "1 discretize",
            "gather_codewords_kernel": "1 discretize", "instance_graph_kernel": "2 graph build",
            "class_vertices_kernel": "3a class atlas", "class_edges_kernel": "3a class atlas"}
STAGE_3B = "3b match (class + instance GNN, logits)"


def timed_region(fn, steps, warmup, barrier, world, dev, native):
    import torch.distributed as dist
    for i in range(warmup):
        fn(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = native.launch_count()
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = native.launch_count() - n0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms, launches


def device_problem(c, B_local, seed, dev):
    """Synthetic tensors of the same distributions as head_oracle.synth_* generated ON the device (the DeiT-Base shapes are
    0.8-1.6 GB per input set; the CPU generator would dominate the run).  Used for the extra configs only -- their parity
    is covered by tests/test_gpu_parity.py on slices the CPU oracle can finish."""
    g = torch.Generator(device=dev).manual_seed(seed)
    d, M, K, Vc, D = c["d"], c["M"], c["K"], c["Vc"], c["D"]
    gs = torch.Generator(device=dev).manual_seed(4321)          # codebook, schema and GNN: the SAME on every rank
    vocab = torch.rand(M, d, device=dev, generator=gs)
    sets = []
    for _ in range(2):
        pick = torch.randint(0, M, (L + 1, B_local), device=dev, generator=g)
        mid = vocab[pick] + 0.3 * torch.randn(L + 1, B_local, d, device=dev, generator=g)
        attn = 0.5 * torch.randn(B_local, L, L, device=dev, generator=g)
        attn_cls = 0.5 * torch.randn(B_local, L, device=dev, generator=g)
        sets.append((mid.contiguous(), attn, attn_cls))
        del pick
    vw = torch.empty(K, Vc, device=dev).normal_(0.5, 1 / 6, generator=gs).clamp_(0, 1)
    ew = torch.empty(K, Vc, Vc, device=dev).normal_(0.5, 1 / 6, generator=gs).clamp_(0, 1)
    vw = (vw / vw.sum(-1, keepdim=True)).nan_to_num(0)
    ew = (ew / ew.sum(-1, keepdim=True)).nan_to_num(0)
    ci = torch.stack([torch.randperm(M, device=dev, generator=gs)[:Vc] for _ in range(K)])
    schema = dict(vertex_weights=vw, edge_weights=ew, class_ingredients=ci,
                  w_v=torch.full((2, 1), 0.5, device=dev), w_e=torch.full((2, 1), 0.5, device=dev))
    import head_oracle as ho
    gnn = {k: v.to(dev) for k, v in ho.synth_gnn(M, D, 4322).items()}
    return vocab, sets, schema, gnn


def run_extra_config(name, args, rank, world, dev, barrier, native):
    """cfg3 / cfg4 at this N (see the module docstring).  Returns the sub-line (rank 0) or None."""
    import head_oracle as ho
    import torch.distributed as dist
    from schemanet_b200.head import GraphedHead
    c = dict(ho.CONFIGS[name])
    if name == "cfg3":
        B_local, scaling = c["B"] // world, "strong"            # the 512-image batch sharded over the GPUs
    else:
        B_local, scaling = c["B"], "weak"                       # 1024 images per GPU
    shard = (rank, world) if world > 1 else None
    vocab, sets, schema, gnn = device_problem(c, B_local, 777 + rank, dev)
    head = build_head(c, vocab, schema, gnn, dev, class_shard=shard)
    note = "one CUDA graph replay per step (NCCL all-gather captured)" if shard else "one CUDA graph replay per step"
    graphed = None
    if not args.no_graph:
        try:
            graphed = [GraphedHead(head, *s) for s in sets]
        except Exception as e:
            graphed, note = None, f"eager launches (CUDA graph capture failed: {type(e).__name__}: {e})"[:240]
            torch.cuda.synchronize()
    else:
        note = "eager launches (--no-graph)"

    def step(i):
        if graphed is not None:
            return graphed[i % len(sets)].replay()
        return head(*sets[i % len(sets)])

    steps = max(3, min(args.steps, 10 if name == "cfg4" else 20))
    ms, _ = timed_region(step, steps, 3, barrier, world, dev, native)
    out = step(0)
    torch.cuda.synchronize()
    finite = bool(torch.isfinite(out["pred"]).all())
    sub = {"workload": WORKLOAD_NAMES[name], "batch_per_gpu": B_local, "global_batch": B_local * world,
           "classes_K": c["K"], "class_vertices_Vc": c["Vc"], "gnn_dim_D": c["D"], "vocab_M": c["M"], "d": c["d"],
           "scaling": scaling, "steps": steps, "ms_per_step": ms / steps, "value": B_local * world * steps / (ms * 1e-3),
           "unit": "images/s", "launch": note, "logits_finite": finite,
           "parallelism": (f"batch-shard dp{world} + class-shard {world} (K/{world} class graphs per GPU, NCCL all-gather of "
                           f"the [K, D] class embeddings)") if shard else "single GPU: all classes local"}
    if shard:
        # the all-gather alone (same tensor sizes), and a sharding check: classes owned by ANOTHER rank, recomputed locally
        per = (c["K"] + world - 1) // world
        local = torch.zeros(per, c["D"], device=dev)
        full = torch.empty(world * per, c["D"], device=dev)
        for _ in range(3):
            dist.all_gather_into_tensor(full, local)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dist.all_gather_into_tensor(full, local)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sub["allgather_ms"] = float(t)
        sub["allgather_bytes"] = world * per * c["D"] * 4
        other = (rank + 1) % world
        k0 = other * per
        k1 = min(k0 + 2, c["K"])
        if k1 > k0:
            sn, gm = head.schema_net, head.matcher.gnn
            _, _, f = native.class_side(gm.param_pack(), sn.vertex_weights.tensor[k0:k1], sn.edge_weights.tensor[k0:k1],
                                        sn.class_ingredients.tensor[k0:k1].contiguous(), sn.prune_node_threshold, True,
                                        sn.remove_self_loop, want_edges=not native.gnn_tensor_path(c["D"], c["Vc"]))
            g_rows = out["feat_class"][k0:k1]
            ok = torch.tensor([float((f - g_rows).abs().max() / g_rows.abs().max().clamp_min(1e-30))], device=dev)
        else:
            ok = torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MAX)
        # (a call on 2 classes may take another kernel path than the shard's -- e.g. no layer-0 table shortcut when there are
        # fewer node slots than codes -- and operand scales depend on the graphs of the call: compared to 1e-5, not bit for bit)
        sub["class_shard_check"] = {"what": "rows gathered from the next rank vs the same classes recomputed locally, worst rank",
                                    "max_rel": float(ok), "ok": bool(float(ok) <= 1e-5)}
    del graphed, head, sets, schema
    torch.cuda.empty_cache()
    return sub if rank == 0 else None


def run_gpu_arm(args):
    import torch.distributed as dist
    from schemanet_b200 import native
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_note = bind_to_gpu_numa(local_rank)                    # before the first pin_memory()
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    native.lib()
    c, vocab, sets, schema, gnn = make_problem(WORKLOAD, 1234 + rank, dev)
    head = build_head(c, vocab, schema, gnn, dev)
    dev_sets = [tuple(t.to(dev) for t in s) for s in sets]
    pinned = [tuple(t.pin_memory() for t in s) for s in sets]
    from schemanet_b200.head import HostPipeline
    from schemanet_b200.head import GraphedHead
    use_graphs = not args.no_graph
    pipe = HostPipeline(head, dev)     # PCIe-bound (117 MB H2D per step): graph replay measured no gain there, kept eager
    # one CUDA graph per resident input set (the same kernels, launched with one call per step instead of ~40)
    graphed, graph_note = None, None
    if use_graphs:
        try:
            graphed = [GraphedHead(head, *s) for s in dev_sets]
        except Exception as e:      # capture is an optimisation of the launch path, never a reason to lose the measurement
            graphed, graph_note = None, f"eager launches (CUDA graph capture failed: {type(e).__name__}: {e})"[:300]
            torch.cuda.synchronize()
            print("bench: " + graph_note, file=sys.stderr)
    else:
        graph_note = "eager launches (--no-graph)"

    def step_eager(i):
        mid, attn, attn_cls = dev_sets[i % N_INPUT_SETS]
        return head(mid, attn, attn_cls)

    def step(i):
        if graphed is not None:
            return graphed[i % N_INPUT_SETS].replay()
        return step_eager(i)

    def step_e2e(i):
        # public host-buffer API: pinned host tensors in, logits in pinned host memory out; the H2D copies of this
        # step's inputs and the D2H read of its logits are inside the timed region (double-buffered against compute)
        return pipe.submit(*pinned[i % N_INPUT_SETS])

    def step_cached(i):
        mid, attn, attn_cls = dev_sets[i % N_INPUT_SETS]
        return head(mid, attn, attn_cls, cache_class=True)

    h2d_scratch = tuple(torch.empty_like(t) for t in dev_sets[0])

    def step_h2d_only(i):
        # the copies of one step and nothing else: the host-side ceiling of the e2e arm when all ranks copy at once
        for dst, src in zip(h2d_scratch, pinned[i % N_INPUT_SETS]):
            dst.copy_(src, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        return timed_region(fn, steps, warmup, barrier, world, dev, native)

    W = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches = timed(step, args.steps, W)
    clocks = sampler.stop() if rank == 0 else None
    if graphed is not None:
        # a replayed graph launches the kernels that were captured: count them on one eager step
        n0 = native.launch_count()
        step_eager(0)
        launches = (native.launch_count() - n0) * args.steps
    ms_e2e, _ = timed(step_e2e, args.steps, W)
    last_logits = pipe.result(pipe.ticket - 1)
    assert bool(torch.isfinite(last_logits).all())
    ms_cached, _ = timed(step_cached, args.steps, W)
    ms_h2d, _ = timed(step_h2d_only, args.steps, W)

    images = c["B"] * world * args.steps
    value = images / (ms * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in sets[0])
    d2h = c["B"] * c["K"] * 4

    line = None
    parity_fail = None
    if rank == 0:
        out = step(0)                                              # the timed path (graph replay when it was captured)
        torch.cuda.synchronize()
        gpu_pred0 = out["pred"].detach().cpu().clone()
        gpu_ing0 = out["ingredients"].detach().cpu().clone()
        n_inst = out["graphs"].num_vertices.tolist()
        n_bar = float(sum(n_inst)) / len(n_inst)
        # per-kernel CUDA-event timings over a further K steps of the same workload (events on the launching stream), taken
        # with the class-side stream serialised behind the main stream so that a kernel's time is its own (in the headline run
        # above the two streams overlap)
        head.overlap_class_side = False
        native.profile_enable(True)
        for i in range(args.steps):
            step_eager(i)
        prof = native.profile_collect()
        native.profile_enable(False)
        head.overlap_class_side = True
        peaks = load_peaks()
        traffic, traffic_src = load_ncu_traffic()
        alg = stage_bytes_flops(c, n_bar)
        kern = {k: {"launches": v[0], "ms_total": v[1], "ms_per_launch": v[1] / max(v[0], 1)} for k, v in prof.items()}
        total_prof = sum(v[1] for v in prof.values())
        for k in kern:
            kern[k]["share"] = kern[k]["ms_total"] / total_prof if total_prof else None
        fam_ms = {}
        for k, v in kern.items():
            f = FAMILY.get(k, k)
            fam_ms[f] = fam_ms.get(f, 0.0) + v["ms_total"]
        dom = max(fam_ms.items(), key=lambda kv: kv[1])[0]
        t_dom = fam_ms[dom] * 1e-3 / args.steps          # seconds per step spent in the family
        fam_launches = sum(v["launches"] for k, v in kern.items() if FAMILY.get(k, k) == dom) / args.steps
        tensor_path = any(k.endswith("_tc") for k in kern)
        # the GNN GEMMs issue kind::f16 MMAs (three per fp32 product): their pipe peak is the measured 16-bit dense rate
        f16_sus, f16_burst = peaks["bf16_tflops_sustained"], peaks["bf16_tflops"]
        n_act = (head.atlas["class_vertices"] > 0.001).sum(1).tolist()
        visited = visited_gemm_flops(c, n_act, n_inst) if tensor_path else None

        def ncu_of(*names):
            if not traffic:
                return None
            for n in names:
                if n in traffic:
                    return traffic[n]
            return None

        if dom in ("adjacency_gemm", "linear_gemm") and visited is not None:
            # `achieved` = fp32 multiply-adds (x2) of the tiles the kernel VISITS under the flags it is launched with (same
            # rule as TILE_LOOP in gnn_tc.cu, restated in adj_gemm_kblocks) / the family's summed duration.  Each product
            # costs the tensor pipe 3 fp16 MMAs, so tensor-pipe use = 3 x frac.
            ach = visited[dom] / t_dom / 1e12
            t_ncu = ncu_of("gemm3x_kernel")
            roof = {"kernel": "gemm3x_kernel (%s: %s)" % (dom, ", ".join(k for k in kern if FAMILY.get(k) == dom)),
                    "bound": "tensor", "achieved": ach, "peak": f16_sus, "unit": "TFLOP/s", "frac": ach / f16_sus,
                    "frac_vs_burst": ach / f16_burst, "launches_per_step": fam_launches,
                    "algorithmic_flops_per_step": visited[dom],
                    "traffic": t_ncu["dram_bytes_per_launch"] if t_ncu else None, "traffic_source": traffic_src,
                    "peak_source": peaks["source"] + " bf16 sustained (kind::f16 MMAs run at the 16-bit rate); frac_vs_burst uses bf16 burst",
                    "tensor_pipe_frac": 3 * ach / f16_sus,
                    "note": "fp32-accurate GEMM as 3 fp16 MMAs per product on tcgen05: `achieved` counts each fp32 multiply-add of the "
                            "visited tiles once; the tensor cores execute 3 MMAs per product (tensor_pipe_frac = 3 x frac), so "
                            "frac <= 1/3 by construction"}
        elif dom == "discretize":
            ach = alg["discretize"]["flops"] / t_dom / 1e12
            peak = peaks["bf16_tflops_sustained"]
            t_ncu = ncu_of("discretize_tc_kernel")
            roof = {"kernel": "discretize_tc_kernel", "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach / peak, "frac_vs_burst": ach / peaks["bf16_tflops"],
                    "traffic": t_ncu["dram_bytes_per_launch"] if t_ncu else None, "traffic_source": traffic_src,
                    "peak_source": peaks["source"] + " bf16 sustained"}
        elif dom in ("graph_build", "atlas"):
            ach = alg[dom]["bytes"] / t_dom / 1e9
            t_ncu = ncu_of("instance_graph_kernel" if dom == "graph_build" else "class_edges_fast_kernel")
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "traffic": t_ncu["dram_bytes_per_launch"] if t_ncu else None,
                    "traffic_source": traffic_src, "peak_source": peaks["source"]}
        else:
            roof = {"kernel": dom, "bound": "hbm", "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": None,
                    "traffic": None, "peak_source": peaks["source"], "note": "no algorithmic model for this helper kernel"}

        # per-stage table against the >= 70 % target: ALL kernels of a stage (pre-passes, re-checks, operand preparation)
        stage_ms = {}
        for k, v in kern.items():
            s = STAGE_OF.get(k, STAGE_3B)
            stage_ms[s] = stage_ms.get(s, 0.0) + v["ms_total"] / args.steps
        hbm = peaks["hbm_gbs"]
        stages = {}
        if "1 discretize" in stage_ms:
            t = stage_ms["1 discretize"] * 1e-3
            tf = alg["discretize"]["flops"] / t / 1e12
            stages["1 discretize"] = {"ms": t * 1e3, "bound": "tensor (f16 operands)", "achieved": tf, "unit": "TFLOP/s",
                                      "frac": tf / peaks["bf16_tflops_sustained"]}
        if "2 graph build" in stage_ms:
            t = stage_ms["2 graph build"] * 1e-3
            gb = alg["graph_build"]["bytes"] / t / 1e9
            stages["2 graph build"] = {"ms": t * 1e3, "bound": "hbm", "achieved": gb, "unit": "GB/s", "frac": gb / hbm}
        if "3a class atlas" in stage_ms:
            t = stage_ms["3a class atlas"] * 1e-3
            gb = alg["atlas"]["bytes"] / t / 1e9
            stages["3a class atlas"] = {"ms": t * 1e3, "bound": "hbm", "achieved": gb, "unit": "GB/s", "frac": gb / hbm}
        if STAGE_3B in stage_ms and visited is not None:
            t = stage_ms[STAGE_3B] * 1e-3
            fl = visited["adjacency_gemm"] + visited["linear_gemm"]
            tf = fl / t / 1e12
            stages[STAGE_3B] = {"ms": t * 1e3, "bound": "tensor (3 fp16 MMAs per fp32 product)", "achieved": tf, "unit": "TFLOP/s", "frac": tf / f16_sus,
                                "note": "visited-tile fp32 flops of the GEMMs / time of every stage-3b kernel incl. operand "
                                        "preparation; 3 MMAs per product cap this at 1/3"}
        for v in stages.values():
            v["meets_target_0.70"] = bool(v["frac"] >= TARGET_FRAC)
        # CPU baseline on a bounded sample (rank 0, N = 1 only) + parity of the timed GPU path against it
        cpu = None
        parity = None
        if world == 1 and not args.no_cpu_baseline and WORKLOAD in ("cfg1", "cfg2"):
            Bs = c["B"]
            cpu_head_time(c, vocab, sets, schema, gnn, 1, batch=Bs)
            r = cpu_head_time(c, vocab, sets, schema, gnn, 3, batch=Bs)
            cpu = {"value": r["ips"], "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "kind_detail": kind_detail(r["kind"]),
                   "sample": f"the full {Bs}-image {WORKLOAD} batch, K={c['K']} class side recomputed, 1 warm-up + median of 3 steps "
                             f"({r['median_s']:.2f} s/step)"}
            ref = r["out"]
            rel = float((gpu_pred0 - ref["pred"]).abs().max() / ref["pred"].abs().max())
            idx_equal = bool(torch.equal(gpu_ing0, ref["ingredients"]))
            parity = {"logits_max_rel": rel, "bar": PARITY_BAR, "codeword_indices_equal": idx_equal,
                      "what": "input set 0 through the TIMED path (graph replay) vs the CPU head above: max|d logits| / max|logits|"}
            if not (rel <= PARITY_BAR and idx_equal):
                parity_fail = f"parity check failed: logits rel err {rel:.3e} (bar {PARITY_BAR}), indices equal: {idx_equal}"
        line = {"metric": "schema_head_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(c, world), **({"launch": graph_note} if graph_note else {})), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps,
                        "h2d_GBps_per_rank": h2d / (ms_e2e / args.steps * 1e-3) / 1e9,
                        "h2d_only_GBps_per_rank": h2d / (ms_h2d / args.steps * 1e-3) / 1e9,
                        "note": "h2d_only = the same pinned-host -> device copies with no kernels, all ranks at once: the host-side "
                                "ceiling of this arm; cpu binding: " + str(numa_note)},
                "class_side_cached": {"value": images / (ms_cached * 1e-3), "unit": "images/s",
                                      "note": "eval-mode variant: class embeddings reused while the atlas is unchanged; "
                                              "NOT the headline (the reference recomputes them every forward)"},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "parity": parity,
                "parity_max_rel": parity["logits_max_rel"] if parity else None, "stages": stages, "kernels": kern,
                "mean_vertices_per_image": n_bar, "mean_unpruned_class_vertices": float(sum(n_act)) / len(n_act)}
    # the multi-GPU configurations BASELINE.json names, at this N (every rank takes part)
    del graphed, pipe, pinned, dev_sets, h2d_scratch
    torch.cuda.empty_cache()
    extra = {}
    if not args.no_extra and WORKLOAD == "cfg2":
        for name in ("cfg3", "cfg4"):
            try:
                sub = run_extra_config(name, args, rank, world, dev, barrier, native)
            except Exception as e:          # never lose the headline line to an extra configuration
                sub = {"error": f"{type(e).__name__}: {e}"[:300]}
                print(f"bench: {name}: {sub['error']}", file=sys.stderr)
                try:
                    torch.cuda.synchronize()
                except Exception:
                    pass
            if rank == 0:
                extra[name] = sub
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        line["configs"] = extra
        emit(line)
        if parity_fail:
            raise SystemExit(parity_fail)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything else a library prints (e.g. NCCL's version
    banner) has been re-routed to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                    # fd 1 -> stderr for the whole run (native libraries print to stdout)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel individually instead of replaying CUDA graphs")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg3 / cfg4 sub-lines")
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4"],
                    help="BASELINE.json shape; the default cfg2 (configs[1]) is the headline")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.config
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
