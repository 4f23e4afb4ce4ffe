"""The fused, sync-free tensor path of the schema-inference head (what bench.py times and what the drop-in modules
are built from):

    mid_feat [1+L, bs, d], attention  ->  discretize  ->  instance graphs  ->  class atlas  ->  GNN x2  ->  logits

Every stage is one or a few launches of libschemahead kernels on the current stream; nothing returns to the host
(the list-returning module API needs one D2H copy of the graph sizes -- this path does not).  Reference call stack
replaced: SchemaNetPredictor.forward (schema_inference/graph/__init__.py:37-57) below the backbone.
"""
from typing import Dict, Optional

import torch

from . import native


class HeadWorkspace:
    """Per-(batch, L) output buffers reused across steps (no allocation inside the timed region)."""

    def __init__(self):
        self.key = None
        self.graphs: Optional[native.PackedGraphs] = None
        self.ingredients = None

    def get(self, bs, L, device):
        key = (bs, L, str(device))
        if key != self.key:
            self.graphs = native.PackedGraphs(bs, L, device)
            self.ingredients = torch.empty(bs, L, dtype=torch.int64, device=device)
            self.key = key
        return self.graphs, self.ingredients


class SchemaHead:
    """Functional head over a `SchemaNet` and a `Matcher` (the drop-in modules) and a codebook.

    class_shard = (rank, world): this process embeds only its slice of the K class graphs and the [K, D] class
    embeddings are all-gathered over NCCL (placement (i)); the batch is sharded by the caller.
    """

    def __init__(self, vocab: torch.Tensor, schema_net, matcher, class_shard=None, disc_mode=native.DISC_AUTO):
        self.vocab = vocab
        self.schema_net = schema_net
        self.matcher = matcher
        self.class_shard = class_shard
        self.disc_mode = disc_mode
        self.ws = HeadWorkspace()
        self._class_cache = None
        self._class_cache_key = None
        # The class side (atlas + class-graph GNN: mostly HBM-bound passes over [K, Vc, Vc]) does not depend on the
        # batch, so it runs on its own stream and overlaps the tensor-core-bound instance side; joined before the logits.
        self.overlap_class_side = True
        # True: every forward also writes the full [K, Vc, Vc] class_edges tensor of get_atlas() into self.atlas (what the
        # reference's forward materialises).  False: only the logits need it, and the class side feeds the normalised
        # edges of the un-pruned vertices straight into the GNN operand (0.4 GB less HBM traffic per step at cfg2).
        self.materialize_atlas = False
        self._class_stream = None
        self._class_out = None
        self._fork = None
        self._join = None
        self._shard_out = None
        self._shard_ce = None

    # -- stage 1 -------------------------------------------------------------------------------------------------
    def discretize(self, mid_feat: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        T, bs, d = mid_feat.shape
        L = T - 1
        tokens = mid_feat[1:].reshape(L * bs, d)            # drop the cls token: a view, rows are token-major
        native.discretize(tokens, self.vocab, out_idx=out, idx_rows=bs, idx_row_stride=L, idx_col_stride=1,
                          mode=self.disc_mode)
        return out

    def _class_key(self):
        sn, gnn = self.schema_net, self.matcher.gnn
        ts = [sn.vertex_weights.tensor, sn.edge_weights.tensor, sn.class_ingredients.tensor] + list(gnn.parameters())
        return tuple((id(t), t.data_ptr(), t._version) for t in ts) + (sn.prune_node_threshold, sn.remove_self_loop)

    def invalidate_class_cache(self):
        self._class_cache, self._class_cache_key = None, None

    # -- stage 3, class side ---------------------------------------------------------------------------------------
    def class_features(self) -> torch.Tensor:
        sn, gnn = self.schema_net, self.matcher.gnn
        vw, ew, ci = sn.vertex_weights.tensor, sn.edge_weights.tensor, sn.class_ingredients.tensor
        K = vw.shape[0]
        if self.class_shard is None:
            gnn._check_inference()
            want_edges = self.materialize_atlas or not native.gnn_tensor_path(gnn.embed_dim, vw.shape[1])
            if (self._class_out is None or self._class_out[0].shape != vw.shape or self._class_out[0].device != vw.device
                    or (self._class_out[1] is not None) != want_edges):
                Vc = vw.shape[1]
                self._class_out = (torch.empty(K, Vc, dtype=torch.float32, device=vw.device),
                                   torch.empty(K, Vc, Vc, dtype=torch.float32, device=vw.device) if want_edges else None,
                                   torch.empty(K, gnn.embed_dim, dtype=torch.float32, device=vw.device))
            # persistent outputs: nothing is allocated per step (and nothing needs cross-stream allocator bookkeeping)
            cv, ce, f_kg = native.class_side(gnn.param_pack(), vw, ew, ci, sn.prune_node_threshold, True, sn.remove_self_loop,
                                             out=self._class_out, want_edges=want_edges)
            self.atlas = {"class_vertices": cv, "class_edges": ce, "class_ingredients": ci}
            return f_kg
        import torch.distributed as dist
        rank, world = self.class_shard
        per = (K + world - 1) // world
        k0, k1 = min(rank * per, K), min((rank + 1) * per, K)
        D = gnn.embed_dim
        if self._shard_out is None or self._shard_out[0].shape != (per, D) or self._shard_out[0].device != vw.device:
            # persistent buffers: the sharded step allocates nothing, so it can be captured into a CUDA graph
            self._shard_out = (torch.zeros(per, D, dtype=torch.float32, device=vw.device),
                               torch.empty(world * per, D, dtype=torch.float32, device=vw.device),
                               torch.empty(max(k1 - k0, 1), vw.shape[1], dtype=torch.float32, device=vw.device))
        local, full, cv = self._shard_out
        if k1 > k0:
            gnn._check_inference()
            want_edges = not native.gnn_tensor_path(D, vw.shape[1])
            if want_edges and (self._shard_ce is None or self._shard_ce.shape[0] != k1 - k0):
                self._shard_ce = torch.empty(k1 - k0, vw.shape[1], vw.shape[1], dtype=torch.float32, device=vw.device)
            native.class_side(gnn.param_pack(), vw[k0:k1], ew[k0:k1], ci[k0:k1], sn.prune_node_threshold, True,
                              sn.remove_self_loop, out=(cv, self._shard_ce if want_edges else None, local[:k1 - k0]),
                              want_edges=want_edges)
            self.atlas = {"class_vertices": cv, "class_edges": None, "class_ingredients": ci[k0:k1], "class_range": (k0, k1)}
        dist.all_gather_into_tensor(full, local)              # the path's only collective: K*D*4 bytes over NVLink
        return full[:K]

    # -- whole head ----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, mid_feat: torch.Tensor, attn: torch.Tensor = None, attn_cls: torch.Tensor = None,
                extracted: torch.Tensor = None, cache_class: bool = False) -> Dict[str, torch.Tensor]:
        """attn/attn_cls: raw head-averaged logits [bs, L, L] / [bs, L]; or `extracted` [bs*H, L+1, L+1] (the
        backbone tap), in which case the head mean is fused into the graph-build read."""
        sn = self.schema_net
        if sn.remove_self_loop:     # the reference cannot build instance edges with this flag (large_scale_feat_to_e.cpp:136-139)
            raise RuntimeError("diagonal dimensions cannot be identical 1, 1")
        T, bs, _ = mid_feat.shape
        L = T - 1
        graphs, ingredients = self.ws.get(bs, L, mid_feat.device)
        self.discretize(mid_feat, ingredients)
        geo = sn._geo(mid_feat.device)
        heads = 0
        if extracted is not None:
            heads = extracted.shape[0] // bs
            attn, attn_cls = extracted, None
        native.instance_graphs(ingredients, attn, attn_cls, geo, sn.vertex_attribute_weights.tensor,
                               sn.edge_attribute_weights.tensor, sn.clamp_vertex_attn, sn.clamp_edge_attn,
                               raw_logits=True, heads=heads, mean=True, out=graphs)
        side = None
        if cache_class:
            # the class embeddings are a function of the schema + GNN parameters only; they are reused
            # while every one of those tensors is the same object at the same version (optimizer steps, load_state_dict
            # and any other in-place torch op bump the version; writes through .data or raw pointers do not -- call
            # invalidate_class_cache() after those)
            key = self._class_key()
            if key != self._class_cache_key:
                self._class_cache, self._class_cache_key = None, key
        if cache_class and self._class_cache is not None:
            f_kg = self._class_cache
        elif self.overlap_class_side:
            main = torch.cuda.current_stream(mid_feat.device)
            if self._class_stream is None:
                self._class_stream = torch.cuda.Stream(device=mid_feat.device)
                self._fork, self._join = torch.cuda.Event(), torch.cuda.Event()
            side = self._class_stream
            self._fork.record(main)             # parameters written earlier on the main stream are visible to the side stream
            side.wait_event(self._fork)
            with torch.cuda.stream(side):
                f_kg = self.class_features()
                self._join.record(side)
            self._class_cache = f_kg if cache_class else None
        else:
            f_kg = self.class_features()
            self._class_cache = f_kg if cache_class else None
        f_inst = self.matcher.gnn.forward_packed(graphs)
        if side is not None:
            torch.cuda.current_stream(mid_feat.device).wait_event(self._join)
        pred = native.similarity(f_inst, f_kg, self.matcher.similarity_name)
        return {"pred": pred, "ingredients": ingredients, "graphs": graphs, "feat_instance": f_inst, "feat_class": f_kg}

    __call__ = forward


class GraphedHead:
    """The whole head for one fixed set of input buffers as a CUDA graph: `replay()` relaunches every kernel of both
    streams (instance side + class side) with a single call, which removes the per-launch gaps of ~40 back-to-back small
    kernels.  The inputs are read from the tensors given here (refill them in place between replays); the outputs are
    the same tensors after every replay.  Parameters are read at replay time (the graph holds pointers, not values), so
    an optimiser step between replays is honoured."""

    def __init__(self, head: "SchemaHead", mid_feat: torch.Tensor, attn: torch.Tensor, attn_cls: torch.Tensor, warmup: int = 2):
        # (a class-sharded head issues one NCCL all-gather per step on the class-side stream: NCCL collectives are
        # graph-capturable, every rank captures and replays the same sequence)
        self.head, self.inputs = head, (mid_feat, attn, attn_cls)
        for _ in range(max(1, warmup)):          # lazy initialisation (workspaces, function attributes) outside the capture
            head(mid_feat, attn, attn_cls)
        torch.cuda.synchronize(mid_feat.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = head(mid_feat, attn, attn_cls)

    def replay(self) -> Dict[str, torch.Tensor]:
        self.graph.replay()
        return self.out


class HostPipeline:
    """Host-buffer front end of `SchemaHead`: the call a serving loop makes when the backbone taps arrive in (pinned)
    host memory.  `submit()` enqueues, for one batch, the H2D copies on a copy stream, the whole head on the compute
    stream and the D2H copy of the logits into a pinned buffer; two staging slots let the copies of batch i+1 overlap
    the kernels of batch i.  `result(ticket)` waits for that batch only.  Every batch's inputs cross PCIe exactly once.
    """

    def __init__(self, head: SchemaHead, device, slots: int = 2, use_graphs: bool = False):
        self.head, self.device, self.slots = head, device, slots
        self.use_graphs = use_graphs            # one GraphedHead per staging slot (captured at the slot's second use)
        self.graphs = [None] * slots
        self.uses = [0] * slots
        self.copy_stream = torch.cuda.Stream(device=device)
        self.staging = [None] * slots
        self.h2d_done = [torch.cuda.Event() for _ in range(slots)]
        self.compute_done = [torch.cuda.Event() for _ in range(slots)]
        self.d2h_done = [torch.cuda.Event() for _ in range(slots)]
        self.out_host = [None] * slots
        self.slot_ticket = [-1] * slots         # which submit() currently owns the slot's output buffer
        self.ticket = 0

    def submit(self, mid_feat: torch.Tensor, attn: torch.Tensor, attn_cls: torch.Tensor) -> int:
        s = self.ticket % self.slots
        if self.staging[s] is None:
            self.staging[s] = tuple(torch.empty_like(t, device=self.device) for t in (mid_feat, attn, attn_cls))
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.compute_done[s])          # slot is free once its last consumer finished
            for dst, src in zip(self.staging[s], (mid_feat, attn, attn_cls)):
                dst.copy_(src, non_blocking=True)
            self.h2d_done[s].record(self.copy_stream)
        main.wait_event(self.h2d_done[s])
        if self.use_graphs and self.graphs[s] is None and self.uses[s] >= 1:
            self.graphs[s] = GraphedHead(self.head, *self.staging[s], warmup=1)   # (synchronises once per slot)
        self.uses[s] += 1
        out = self.graphs[s].replay() if self.graphs[s] is not None else self.head(*self.staging[s])
        self.compute_done[s].record(main)
        pred = out["pred"]
        if self.out_host[s] is None:
            self.out_host[s] = torch.empty(pred.shape, dtype=pred.dtype).pin_memory()
        self.out_host[s].copy_(pred, non_blocking=True)
        self.d2h_done[s].record(main)
        self.slot_ticket[s] = self.ticket
        self.ticket += 1
        return self.ticket - 1

    def result(self, ticket: int) -> torch.Tensor:
        s = ticket % self.slots
        if self.slot_ticket[s] != ticket:
            raise RuntimeError(f"HostPipeline.result({ticket}): that batch's output slot was reused by submit #{self.slot_ticket[s]}; "
                               f"collect a result before submitting {self.slots} more batches")
        self.d2h_done[s].synchronize()
        return self.out_host[s]
