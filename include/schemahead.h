/*
 * schemahead.h -- C ABI of libschemahead.so, the B200 (sm_100a) implementation of SchemaNet's schema-inference head.
 *
 * This is the drop-in boundary for the one hot path this repository accelerates (BASELINE.json north_star):
 * discretize -> instance-graph build -> schema match.  Plain pointers and sizes only: no torch / pybind types.
 * Every entry point names the reference interface it replaces (paths relative to the reference repository,
 * zhfeing/SchemaNet-PyTorch).  INTEGRATION.md shows the reference-side binding for each.
 *
 * Conventions
 *   - All `sh_*` functions return 0 on success, non-zero on failure; sh_last_error() returns the message of the
 *     calling thread's last failure.
 *   - Device entry points (`sh_dev_*`): every pointer is a DEVICE pointer on the current CUDA device, the work is
 *     enqueued asynchronously on `stream` (a cudaStream_t), nothing is allocated and nothing synchronises.
 *   - Host entry points (`sh_host_*`): every pointer is a HOST pointer; the call copies inputs to the device,
 *     runs the same kernels, copies results back and returns when they are in the caller's buffers -- the
 *     synchronous, CPU-tensor contract of the reference's pybind `extension` functions
 *     (cpp_extension/src/extension.cpp:6-12).
 *   - Codes/ids are int64 at the boundary because the reference's accessors are `long`.
 *   - Per-image packed graph layout ("slots"), L = tokens per image, n_b = distinct codes of image b:
 *       ids[b*L + k], vertex_w[b*L + k]            k <  n_b
 *       edges[b*L*L + i*L + j]                     i,j < n_b   (row stride L: image b's graph is the top-left
 *                                                              n_b x n_b corner of its own [L, L] slot)
 *       num_vertices[b] = n_b (int32), *max_vertices = max_b n_b (int32, atomically maximised: zero it first)
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef SCHEMAHEAD_H_
#define SCHEMAHEAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *sh_stream_t; /* cudaStream_t */

#define SH_ABI_VERSION 1

/* float "not set" marker for the optional clamp thresholds (SchemaNet(clamp_vertex_attn=None, ...)) */
#define SH_NO_CLAMP (-3.0e38f)

int sh_abi_version(void);
const char *sh_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's `gpu_launches`). */
int64_t sh_launch_count(void);
int sh_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* Per-kernel timing for bench.py's roofline: while enabled, every kernel launch of this library is bracketed by CUDA
 * events on its launching stream.  sh_profile_collect synchronises the device and writes one line
 * "kernel_name\tlaunches\ttotal_ms\n" per kernel into buf (NUL terminated, truncated at cap), then clears. */
int sh_profile_enable(int on);
int sh_profile_collect(char *buf, size_t cap);

/* ------------------------------------------------------------------------------------------------------------
 * Stage 0 -- attention prologue.
 * Replaces IngredientModelWrapper.forward's head-mean + slicing
 * (schema_inference/utils/ingredient_model_wrapper.py:57-69).
 *   extracted [B*H, T, T] raw attention logits (T = L+1, token 0 = cls)
 *   attn      [B, L, L]  = mean_h extracted[b*H+h, 1:, 1:]
 *   attn_cls  [B, L]     = mean_h extracted[b*H+h, 0, 1:]
 * ---------------------------------------------------------------------------------------------------------- */
int sh_dev_attention_prologue(const float *extracted, int B, int H, int T, float *attn, float *attn_cls,
                              sh_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Stage 1 -- discretization: nearest visual word of every token.
 * Replaces Discretization.encode (discretization/discretization.py:58-70):
 *   ingredients = torch.cdist(seq.reshape(R, d), vocabulary.weight).argmin(dim=1); seq' = vocabulary(ingredients)
 *   tokens [R, d] fp32 (row stride d), vocab [M, d] fp32
 *   out_idx [R] int64: argmin_j sqrt(max(|x|^2 + |c_j|^2 - 2 x.c_j, 0)), lowest index on ties
 *   out_seq [R, d] or NULL: the gathered codewords (Discretization(activate) path, :66-67)
 *   the index of row r is stored at out_idx[(r % idx_rows) * idx_row_stride + (r / idx_rows) * idx_col_stride]:
 *     flat [R] result: idx_rows = R, idx_row_stride = 1, idx_col_stride = 0;
 *     the head's token-major rows (r = t*bs + b) written straight into the [bs, L] layout of
 *     ingredient_model_wrapper.py:55: idx_rows = bs, idx_row_stride = L, idx_col_stride = 1.
 *   mode: SH_DISC_AUTO picks a tensor-core path when it applies (fp16 operands if d % 8 == 0, else tf32 if d % 4 == 0);
 *         SH_DISC_EXACT forces the fp32 CUDA-core scan; every mode returns the same indices: the
 *         tensor-core pass only short-lists candidates inside a worst-case error band (Cauchy-Schwarz on the measured
 *         operand rounding residuals); the winner is always decided by the fp32 re-check
 *   workspace: sh_discretize_workspace_bytes(R, d, M) bytes of device scratch.
 * ---------------------------------------------------------------------------------------------------------- */
#define SH_DISC_AUTO 0
#define SH_DISC_EXACT 1
#define SH_DISC_TENSOR 2        /* tcgen05 kind::tf32 coarse pass (fp32 tiles fed as they are) + fp32 re-check */
#define SH_DISC_TENSOR_F16 3    /* tcgen05 kind::f16 coarse pass on fp16 copies (half the operand traffic) + fp32 re-check */
size_t sh_discretize_workspace_bytes(int64_t R, int d, int M);
int sh_dev_discretize(const float *tokens, const float *vocab, int64_t R, int d, int M, int64_t *out_idx,
                      int64_t idx_rows, int64_t idx_row_stride, int64_t idx_col_stride, float *out_seq,
                      void *workspace, size_t workspace_bytes, int mode, sh_stream_t stream);
/* statistics of the last tensor-core discretize call on this stream's device: rows that needed the exact recheck,
 * rows whose candidate list overflowed (rescanned over all M).  Reads two device counters -> synchronises. */
int sh_discretize_stats(const void *workspace, int64_t *recheck_rows, int64_t *overflow_rows);

/* ------------------------------------------------------------------------------------------------------------
 * Stage 2 -- instance IR-graphs (vertices + edges of every image in one launch).
 * Replaces SchemaNet.feat_to_instance_vertices / feat_to_instance_edges (schema_inference/graph/schema_net.py:278-375)
 * and the native functions they call: ext::feat_to_instance_v (cpp_extension/src/large_scale_feat_to_v.cpp:41-143),
 * ext::feat_to_instance_e (cpp_extension/src/large_scale_feat_to_e.cpp:33-150), ext::accumulate (utils.cpp:6-15).
 *   ingredients [B, L] int64
 *   attn        [B, L, L] fp32, attn_cls [B, L] fp32.  If SH_G_RAW_LOGITS is set they are RAW logits and the kernel
 *               applies masked_fill(x < clamp, -inf) + softmax (+ nan_to_num(0) for attn_cls) itself
 *               (schema_net.py:295-297, 334-336); otherwise they are used as given (the ext contract).
 *   If SH_G_FROM_HEADS is set, `attn` is instead the backbone's `extracted` [B*H, L+1, L+1] tensor and `attn_cls`
 *               is ignored: the stage-0 head mean and slicing are fused into the read.
 *   geo_sim     [L, L] fp32 (graph/utils.py:55-81), read through L2
 *   w_vertex, w_edge: device pointers to the two [2,1] attribute-weight parameters
 *   outputs: packed layout above.  Either output group may be NULL to skip it (vertices: ids/vertex_w; edges).
 *   SH_G_SUM: accumulate sums instead of means (the ext functions' mean=false).
 *   SH_G_WRITE_BACK_CLAMP: also store -inf into the caller's attn/attn_cls where x < clamp, the reference's
 *               in-place side effect (schema_net.py:296,335).
 * ---------------------------------------------------------------------------------------------------------- */
#define SH_G_RAW_LOGITS 1
#define SH_G_FROM_HEADS 2
#define SH_G_SUM 4
#define SH_G_WRITE_BACK_CLAMP 8
/* also zero the rest of each [L, L] edge slot, so that edges[b, :N, :N] IS the zero-padded graph that
 * Matcher.forward builds with F.pad (schema_inference/graph/match.py:49-54) */
#define SH_G_ZERO_PAD 16
int sh_dev_instance_graphs(const int64_t *ingredients, float *attn, float *attn_cls, const float *geo_sim, int B,
                           int L, int H, float clamp_vertex, float clamp_edge, const float *w_vertex,
                           const float *w_edge, int flags, int64_t *ids, float *vertex_w, float *edges,
                           int32_t *num_vertices, int32_t *max_vertices, sh_stream_t stream);

/* Dense init-time variants.
 * sh_dev_feat_to_v_attr replaces ext::feat_to_v_attr (cpp_extension/src/feat_to_v_attr.cpp:19-63,74-148):
 *   out [B, n_vertices, 2] = (count, sum-or-mean attention) scattered at the code id; fully overwritten.
 * sh_dev_feat_to_e replaces ext::feat_to_e (cpp_extension/src/feat_to_e.cpp:31-127):
 *   class_ingredients [K, n_max] int64 (the tensor behind the reference's list of {code: index} dictionaries,
 *   schema_net.py:121-126), label [B] int64; out [B, n_max, n_max, 2] = (geo, attn) block sums-or-means at the
 *   class-local indices; fully overwritten. */
int sh_dev_feat_to_v_attr(const int64_t *ingredients, const float *attn_cls, int B, int L, int n_vertices, int mean,
                          int ingredients_only, float *out, sh_stream_t stream);
int sh_dev_feat_to_e(const int64_t *ingredients, const float *attn, const float *geo_sim,
                     const int64_t *class_ingredients, const int64_t *label, int B, int L, int K, int n_max, int mean,
                     float *out, sh_stream_t stream);
/* Per-class running sums of the atlas initialisation: replaces the Python accumulation loops of
 * scripts/init_schema_net.py:31-34 (edges) and :57-59 (vertices),
 *     for cls_id, x_b in zip(label, x): acc[cls_id] += x_b; n_tracked[cls_id] += 1
 *   x [B, N] fp32 (one flattened sample per row), label [B] int64, acc [K, N] fp32 and n_tracked [K] fp32 (or NULL) are
 *   updated in place; samples are added in batch order (the reference's fp32 summation order: bit-exact). */
int sh_dev_class_accumulate(const float *x, const int64_t *label, int B, int64_t N, int K, float *acc, float *n_tracked,
                            sh_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Stage 3a -- class IR-atlas.
 * Replaces SchemaNet.get_class_vertices / get_class_edges / get_atlas (schema_net.py:144-184) and
 * normalize_sum_clamp (graph/utils.py:25-52).
 *   vertex_weights [K, Vc], edge_weights [K, Vc, Vc] (parameters)
 *   class_vertices [K, Vc] = nan_to_num(clamp_min(vw, 1e-5) / sum)
 *   class_edges    [K, Vc, Vc] = nan_to_num(clamp_min(pruned ew, 0) / row sum); prune_threshold < 0 = no pruning
 *   With pruning the reference ALSO zeroes the pruned entries of the edge_weights parameter in place (:164);
 *   so does this call when prune_in_place != 0.
 * ---------------------------------------------------------------------------------------------------------- */
int sh_dev_class_atlas(const float *vertex_weights, float *edge_weights, int K, int Vc, float prune_threshold,
                       int prune_in_place, int remove_self_loop, float *class_vertices, float *class_edges,
                       sh_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Stage 3b -- GNN embedding of a batch of graphs, and the matcher's similarity.
 * sh_dev_gnn_forward replaces GNN.forward (schema_inference/graph/gnn.py:78-98) incl. Layer/GraphConv (:20-46):
 *   feat = Emb[ids]; per layer: feat = relu(LN(mask(((E+E^T)/2 + I) feat W^T + b))); out = fc(mean_n(feat * w)).
 *   G graphs.  ids [G, ld_v] int64, vertex_w [G, ld_v], edges: graph g at edges + g*edge_batch_stride with row
 *   stride (edge_ld > 0 ? edge_ld : n_g).  sizes [G] int32 or NULL (all graphs have n_fixed nodes).
 *   mean_div: device pointer to the int32 divisor of the pooling mean (the PADDED node count of gnn.py:96;
 *   max_b n_b for instance graphs) or NULL to divide by n_fixed.
 *   params: see sh_gnn_params.  workspace: sh_gnn_workspace_bytes().  out [G, D].
 * sh_dev_similarity replaces Matcher._inner_product/_cosine_sim/_euclidean_sim (match.py:21-31) on the expanded
 *   [B,K,D] pair: logits [B, K].
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct sh_gnn_params {
    int num_codes;            /* M; embedding has M+1 rows, row M is the padding row */
    int embed_dim;            /* D */
    int num_layers;
    float ln_eps;
    const float *embedding;   /* [(M+1), D] */
    const float *const *lin_w; /* num_layers pointers to [D, D] (out, in) */
    const float *const *lin_b; /* num_layers pointers to [D] */
    const float *const *ln_w;  /* num_layers pointers to [D] */
    const float *const *ln_b;  /* num_layers pointers to [D] */
    const float *fc_w;        /* [D, D] */
    const float *fc_b;        /* [D] */
} sh_gnn_params;

size_t sh_gnn_workspace_bytes(int G, int n_max, int D);
int sh_dev_gnn_forward(const sh_gnn_params *params, int G, int n_fixed, const int32_t *sizes, const int64_t *ids,
                       const float *vertex_w, int ld_v, const float *edges, int64_t edge_batch_stride, int edge_ld,
                       const int32_t *mean_div, float *out, void *workspace, size_t workspace_bytes,
                       sh_stream_t stream);

/* Stage 3a + class side of stage 3b in one call (what SchemaNetPredictor.forward does with get_atlas() followed by the
 * second self.gnn(...) of Matcher.forward, schema_net.py:177-184 + match.py:66-70): writes class_vertices [K,Vc],
 * class_edges [K,Vc,Vc] and the class embeddings feat_class [K,D].  The tensor-core path compacts every class graph to
 * its un-pruned vertices first (pruned vertices have all-zero edge rows/columns, so the result is unchanged).
 * class_edges may be NULL on the tensor-core path (embed_dim % 256 == 0, embed_dim <= 1024, Vc >= 32): the normalised
 * [K,Vc,Vc] tensor is then not materialised at all -- the adjacency operand is gathered from the (pruned) parameter and
 * per-row normalisers -- which saves its 4*K*Vc*Vc bytes of HBM writes; feat_class is bit-identical either way.  The
 * in-place prune of edge_weights (schema_net.py:164) happens in both cases. */
size_t sh_class_side_workspace_bytes(int K, int Vc, int D);
/* GNN.forward on the class graphs get_atlas() produced (match.py:66-70), given the prune threshold they were built with
 * (< 0: none): same result as sh_dev_gnn_forward, pruned vertices are skipped.  Workspace: sh_class_side_workspace_bytes. */
int sh_dev_gnn_forward_class(const sh_gnn_params *params, int K, int Vc, const float *class_vertices,
                             const float *class_edges, const int64_t *class_ingredients, float prune_threshold,
                             float *feat_class, void *workspace, size_t workspace_bytes, sh_stream_t stream);
int sh_dev_class_side(const sh_gnn_params *params, const float *vertex_weights, float *edge_weights,
                      const int64_t *class_ingredients, int K, int Vc, float prune_threshold, int prune_in_place,
                      int remove_self_loop, float *class_vertices, float *class_edges, float *feat_class,
                      void *workspace, size_t workspace_bytes, sh_stream_t stream);

#define SH_SIM_INNER_PRODUCT 0
#define SH_SIM_COSINE 1
#define SH_SIM_EUCLIDEAN 2
int sh_dev_similarity(const float *feat_instance, const float *feat_class, int B, int K, int D, int kind,
                      float *logits, sh_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Host-buffer entry points: the synchronous CPU-tensor contract of the reference's pybind module `extension`
 * (cpp_extension/src/extension.cpp:6-12; headers cpp_extension/include/feat_to_v.h:6-19, feat_to_e.h:15-33).
 * ---------------------------------------------------------------------------------------------------------- */
/* ext::feat_to_instance_v: attn_cls already soft-maxed.  ids/vertex_w: [B*L] slots, num_vertices [B] int64. */
int sh_host_feat_to_instance_v(const int64_t *ingredients, const float *attn_cls, int B, int L, const float *w_vertex2,
                               int mean, int64_t *ids, float *vertex_w, int64_t *num_vertices);
/* ext::feat_to_instance_e: attn already soft-maxed; the code->index dictionaries are the sorted-unique ranks
 * (schema_net.py:345-348).  edges: [B*L*L] slots (row stride L), num_vertices [B] int64. */
int sh_host_feat_to_instance_e(const int64_t *ingredients, const float *attn, const float *geo_sim, int B, int L,
                               const float *w_edge2, int mean, float *edges, int64_t *num_vertices);
int sh_host_feat_to_v_attr(const int64_t *ingredients, const float *attn_cls, int B, int L, int n_vertices, int mean,
                           int ingredients_only, float *out);
int sh_host_feat_to_e(const int64_t *ingredients, const float *attn, const float *geo_sim,
                      const int64_t *class_ingredients, const int64_t *label, int B, int L, int K, int n_max, int mean,
                      float *out);

#ifdef __cplusplus
}
#endif
#endif /* SCHEMAHEAD_H_ */
