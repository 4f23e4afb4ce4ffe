// gnn_tc.cuh -- tensor-core (tcgen05, 3xTF32) GNN layer path, see gnn_tc.cu
#pragma once
#include "common.cuh"

namespace sh {

bool gnn_tc_supported(int D, int n_fixed);

// Where the final [G, D] features go.  When the pooling was fused into the last GEMM's epilogue the tensor-core path also
// applies mean + fc itself (one kernel instead of reduce + pool_fc) and sets `done`; otherwise the caller finishes from
// `partial` with pool_fc_kernel.
struct TcFinal {
    float *out;
    const int32_t *mean_div;
    bool done;
};
size_t gnn_tc_workspace_bytes(int G, int n_fixed, int D, int chunks);
// Runs all GNN layers on the tensor cores and leaves the vertex-weighted pooling partials [G, chunks, D] in `partial`.
int gnn_forward_tc(const sh_gnn_params *p, int G, int n_fixed, const int32_t *sizes, const int64_t *ids,
                   const float *vertex_w, int ld_v, const float *edges, int64_t edge_batch_stride, int edge_ld, int chunks,
                   float *partial, void *workspace, cudaStream_t st, TcFinal *fin = nullptr);

// Class-side variant: graphs are compacted to their un-pruned vertices first (exact, see gnn_tc.cu).
int gnn_class_forward_tc(const sh_gnn_params *p, int K, int Vc, const float *class_vertices, const float *class_edges,
                         const int64_t *class_ingredients, float prune_threshold, int chunks, float *partial,
                         void *workspace, cudaStream_t st, TcFinal *fin = nullptr);

// sh_dev_class_side on the tensor-core path: atlas edges + class-graph GNN with the un-pruned vertices compacted on the fly
int gnn_class_side_tc(const sh_gnn_params *p, float *edge_weights, int K, int Vc, float prune_threshold, int prune_in_place,
                      int remove_self_loop, const float *class_vertices, float *class_edges,
                      const int64_t *class_ingredients, int chunks, float *partial, void *workspace, cudaStream_t st,
                      TcFinal *fin = nullptr);

// Inner-product logits on the tensor cores (large B * K * D only); scratch: 8 bytes of device memory
bool similarity_tc_supported(int B, int K, int D);
int similarity_tc(const float *feat_instance, const float *feat_class, int B, int K, int D, float *logits, unsigned *scratch,
                  cudaStream_t st);

// atlas.cu
int launch_class_vertices(const float *vertex_weights, int K, int Vc, float *class_vertices, cudaStream_t st);
int launch_class_edges(float *edge_weights, const float *class_vertices, int K, int Vc, float prune_threshold,
                       int prune_in_place, int remove_self_loop, float *class_edges, float *rowinv, cudaStream_t st);

// out[rows, D] = A[rows, D] W^T with W [D, D] (fp32 CUDA-core GEMM from gnn.cu; used for the embedding-table shortcut)
// act (optional): relu(LayerNorm(out + bias)) of the same rows, fused
int launch_rows_linear(const float *A, const float *W, int rows, int D, float *out, cudaStream_t st, const float *bias = nullptr,
                       const float *gamma = nullptr, const float *beta = nullptr, float eps = 0.0f, float *act = nullptr,
                       unsigned *amax_out = nullptr);      // amax_out: atomicMax of the bit patterns of |out|

}  // namespace sh
