// gnn_tc.cu -- stage 3b on tcgen05 tensor cores: fp32-accurate "3xTF32" batched GEMMs with fused epilogues.
//
// Replaces the bmm / Linear / LayerNorm / ReLU chain of GraphConv + Layer (schema_inference/graph/gnn.py:20-46) for
// embed_dim == 256 (every shipped config except ImageNet's 1024, which stays on the fp32 CUDA-core path).
//
// Precision: the north star allows 1e-5 relative error on logits, which plain TF32 (10-bit mantissa) cannot meet.
// Every fp32 operand x is therefore split as x = hi + lo with hi = x & 0xffffe000 (exactly representable in TF32) and
// lo = x - hi (exact in fp32, 13 significant bits), and each product is accumulated as
//       a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi        (three kind::tf32 MMAs into the same fp32 TMEM accumulator)
// which leaves a relative error of ~2^-21 per product -- the same order as fp32 FMA accumulation itself.
// The hi/lo pairs are produced by the kernels that write the operands (adjacency prep, embedding gather, the previous
// GEMM's epilogue), so the GEMM main loop is pure TMA -> tcgen05.mma.
//
// Per layer, for a batch of G graphs with n_g <= n_fixed nodes:
//   adj GEMM     Y[g]  = Adj[g] (n x n, symmetric, K-major)  *  X[g]   given as X^T [D, n] (K-major)   -> Y  [n, D]
//   linear GEMM  Z     = Y (rows x D, K-major) * W^T with W [D_out, D_in] (K-major), + bias, LayerNorm, ReLU fused in
//                the TMEM epilogue (one thread owns one full 256-wide row: LayerNorm needs no cross-thread reduction)
//                -> H^T [D, n] as hi/lo for the next layer's adj GEMM, or H [rows, D] for the pooling.
// Kernel layout is the same as discretize_tc.cu: warp 0 TMA producer, warp 1 MMA issuer / TMEM owner, warps 2-5
// epilogue; 2-stage 96 KB shared-memory ring, 2-stage 256-column TMEM accumulator ring.
#include "common.cuh"
#include "gnn_tc.cuh"
#include "tc_common.cuh"

namespace sh {

using namespace tc;

constexpr int G_BM = 128;
constexpr int G_BN = 256;      // == embed_dim
constexpr int G_BK = 32;
constexpr int G_STAGES = 2;
constexpr int G_THREADS = 192;
constexpr int kABytes = G_BM * G_BK * 4;                 // 16 KB
constexpr int kBBytes = G_BN * G_BK * 4;                 // 32 KB
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;   // hi + lo of both operands: 96 KB
constexpr int kBarOffset = G_STAGES * kStageBytes;
constexpr int kParamOffset = kBarOffset + 256;              // bias / gamma / beta of the fused LayerNorm epilogue
constexpr int kSmemTotal = kParamOffset + 3 * G_BN * 4 + 1024;

enum { EPI_STORE_SPLIT = 0, EPI_LN_RELU_T_SPLIT = 1, EPI_LN_RELU_ROWS = 2 };

struct GemmTcArgs {
    int G;                 // batch of graphs (adj GEMM) or 1 (linear GEMM over flattened rows)
    int rows_per_graph;    // n_fixed
    int M_total;           // rows of the A operand per batch entry
    int K_total;           // reduction length upper bound
    const int32_t *sizes;  // [G] n_g or null
    int batched_b;         // 1: B operand indexed by the graph, 0: shared (weights)
    // epilogue
    float *out_hi, *out_lo;      // EPI_STORE_SPLIT: Y hi/lo [G, n_fixed, 256]; EPI_LN_RELU_T_SPLIT: H^T hi/lo [G, 256, ldk]
    float *out_rows;             // EPI_LN_RELU_ROWS: H [G*n_fixed, 256]
    int ldk;                     // row stride of the transposed output
    const float *bias, *gamma, *beta;
    float eps;
};

__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo)
{
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

template <int EPI>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
              const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, GemmTcArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full = (uint64_t *)(smem + kBarOffset);
    uint64_t *empty = full + G_STAGES;
    uint64_t *tmem_full = empty + G_STAGES;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_ptr = (uint32_t *)(tmem_empty + 2);
    float *s_bias = (float *)(smem + kParamOffset), *s_gamma = s_bias + G_BN, *s_beta = s_gamma + G_BN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (EPI != EPI_STORE_SPLIT)
        for (int i = threadIdx.x; i < G_BN; i += G_THREADS) { s_bias[i] = a.bias[i]; s_gamma[i] = a.gamma[i]; s_beta[i] = a.beta[i]; }
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmAh); tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmBh); tma_prefetch_desc(&tmBl);
        for (int s = 0; s < G_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 2 * G_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int mb_per = ceil_div(a.M_total, G_BM);
    const int tiles = a.G * mb_per;

    // every role walks the same tile list and applies the same skip rule
#define TILE_LOOP_BEGIN                                                                                   \
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {                                                 \
        const int g = t / mb_per, mb = t % mb_per;                                                        \
        const int n_g = (a.sizes && a.G > 1) ? a.sizes[g] : a.K_total;                                    \
        if (a.G > 1 && mb * G_BM >= n_g) continue;                                                        \
        const int kblocks = ceil_div(a.G > 1 ? n_g : a.K_total, G_BK);
#define TILE_LOOP_END }

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            TILE_LOOP_BEGIN
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t *s = smem + stage * kStageBytes;
                    mbar_arrive_expect_tx(&full[stage], kStageBytes);
                    tma_load_3d(s, &tmAh, &full[stage], kb * G_BK, mb * G_BM, g);
                    tma_load_3d(s + kABytes, &tmAl, &full[stage], kb * G_BK, mb * G_BM, g);
                    tma_load_3d(s + 2 * kABytes, &tmBh, &full[stage], kb * G_BK, 0, a.batched_b ? g : 0);
                    tma_load_3d(s + 2 * kABytes + kBBytes, &tmBl, &full[stage], kb * G_BK, 0, a.batched_b ? g : 0);
                    if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
                }
            TILE_LOOP_END
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_tf32(G_BM, G_BN);
        int stage = 0, as = 0;
        uint32_t phase = 0, aphase = 0;
        TILE_LOOP_BEGIN
            mbar_wait(&tmem_empty[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(as * G_BN);
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t s = smem_u32(smem + stage * kStageBytes);
                    const uint64_t ah = make_desc_k_sw128(s), al = make_desc_k_sw128(s + kABytes);
                    const uint64_t bh = make_desc_k_sw128(s + 2 * kABytes), bl = make_desc_k_sw128(s + 2 * kABytes + kBBytes);
#pragma unroll
                    for (int k = 0; k < G_BK / 8; ++k) {
                        const uint64_t o = (uint64_t)(2 * k);
                        umma_tf32(tmem_d, al + o, bh + o, idesc, (kb | k) != 0);   // small terms first
                        umma_tf32(tmem_d, ah + o, bl + o, idesc, 1);
                        umma_tf32(tmem_d, ah + o, bh + o, idesc, 1);
                    }
                    umma_commit(&empty[stage]);
                    if (kb == kblocks - 1) umma_commit(&tmem_full[as]);
                }
                __syncwarp();
                if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        TILE_LOOP_END
    } else {
        const int wq = warp & 3;
        const int row_in_tile = wq * 32 + lane;
        int as = 0;
        uint32_t aphase = 0;
        TILE_LOOP_BEGIN
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * G_BN);
            const int m = mb * G_BM + row_in_tile;           // row inside this batch entry
            if (EPI == EPI_STORE_SPLIT) {
                // Y[g, m, :] as hi/lo (row-major: the K-major A operand of the linear GEMM)
                const bool valid = m < a.rows_per_graph;
                float *oh = a.out_hi + ((size_t)g * a.rows_per_graph + m) * G_BN;
                float *ol = a.out_lo + ((size_t)g * a.rows_per_graph + m) * G_BN;
#pragma unroll 1
                for (int c = 0; c < G_BN / 32; ++c) {
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
                    if (valid) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 h, l;
                            split_tf32(v[4 * q + 0], h.x, l.x); split_tf32(v[4 * q + 1], h.y, l.y);
                            split_tf32(v[4 * q + 2], h.z, l.z); split_tf32(v[4 * q + 3], h.w, l.w);
                            *reinterpret_cast<float4 *>(oh + c * 32 + 4 * q) = h;
                            *reinterpret_cast<float4 *>(ol + c * 32 + 4 * q) = l;
                        }
                    }
                }
            } else {
                // z = acc + bias; LayerNorm over the 256 columns this thread owns; ReLU   (gnn.py:31,45)
                const int gg = m / a.rows_per_graph, i = m % a.rows_per_graph;   // flattened rows -> (graph, node)
                const bool in_range = m < a.M_total;
                const int n_node = (in_range && a.sizes) ? a.sizes[gg] : a.rows_per_graph;
                const bool valid = in_range && i < n_node;
                float sum = 0.0f;
#pragma unroll 1
                for (int c = 0; c < G_BN / 32; ++c) {
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum += v[j] + s_bias[c * 32 + j];
                }
                const float mean = sum / (float)G_BN;
                float var = 0.0f;
#pragma unroll 1
                for (int c = 0; c < G_BN / 32; ++c) {
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) { const float tt = v[j] + s_bias[c * 32 + j] - mean; var = fmaf(tt, tt, var); }
                }
                const float rstd = 1.0f / sqrtf(var / (float)G_BN + a.eps);
#pragma unroll 1
                for (int c = 0; c < G_BN / 32; ++c) {
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = c * 32 + j;
                        const float z = v[j] + s_bias[n];
                        v[j] = fmaxf((z - mean) * rstd * s_gamma[n] + s_beta[n], 0.0f);
                    }
                    if (EPI == EPI_LN_RELU_ROWS) {
                        if (valid) {
                            float *o = a.out_rows + (size_t)m * G_BN + c * 32;
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                *reinterpret_cast<float4 *>(o + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        }
                    } else {
                        // H^T[gg, n, i] as hi/lo: lanes hold consecutive nodes i -> coalesced 128-byte stores; nodes
                        // beyond n_g are written as zeros (they are the zero-padded K range of the next adj GEMM)
                        if (in_range) {
                            float *oh = a.out_hi + ((size_t)gg * G_BN + c * 32) * a.ldk + i;
                            float *ol = a.out_lo + ((size_t)gg * G_BN + c * 32) * a.ldk + i;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float h, l;
                                split_tf32(valid ? v[j] : 0.0f, h, l);
                                oh[(size_t)j * a.ldk] = h;
                                ol[(size_t)j * a.ldk] = l;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        TILE_LOOP_END
    }
#undef TILE_LOOP_BEGIN
#undef TILE_LOOP_END
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * G_BN);
}

// ---------------------------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------------------------
// Adj = (E + E^T)/2 + I  (gnn.py:27-30) as hi/lo, zero outside the n_g x n_g corner.  32x32 tiles, the transposed tile
// goes through shared memory so that both reads are coalesced.
__global__ void __launch_bounds__(256)
adj_prep_kernel(const float *__restrict__ E, int64_t e_batch, int e_ld, const int32_t *__restrict__ sizes, int n_fixed,
                int ldk, float *__restrict__ adj_hi, float *__restrict__ adj_lo)
{
    __shared__ float tile[32][33];
    const int g = blockIdx.z;
    const int n_g = sizes ? sizes[g] : n_fixed;
    const int ld = e_ld > 0 ? e_ld : n_g;
    const float *Eg = E + (size_t)g * e_batch;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {       // transposed tile: rows j0.., cols i0..
        const int jj = j0 + r, ii = i0 + tx;
        tile[r][tx] = (jj < n_g && ii < n_g) ? Eg[(size_t)jj * ld + ii] : 0.0f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, j = j0 + tx;
        if (i < n_fixed && j < ldk) {
            float v = 0.0f;
            if (i < n_g && j < n_g) v = (Eg[(size_t)i * ld + j] + tile[tx][r]) / 2.0f + (i == j ? 1.0f : 0.0f);
            float h, l;
            split_tf32(v, h, l);
            const size_t o = ((size_t)g * n_fixed + i) * ldk + j;
            adj_hi[o] = h;
            adj_lo[o] = l;
        }
    }
}

// X0^T[g, d, i] = Emb[ids[g, i], d] as hi/lo (gnn.py:91), zero for i >= n_g.  32x32 tile transpose.
__global__ void __launch_bounds__(256)
embed_gather_t_kernel(const float *__restrict__ emb, const int64_t *__restrict__ ids, int ld_ids,
                      const int32_t *__restrict__ sizes, int n_fixed, int ldk, int D, float *__restrict__ xt_hi,
                      float *__restrict__ xt_lo)
{
    __shared__ float tile[32][33];
    const int g = blockIdx.z;
    const int n_g = sizes ? sizes[g] : n_fixed;
    const int i0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {       // node i0 + r, features d0 + tx (coalesced along d)
        const int i = i0 + r;
        tile[r][tx] = (i < n_g) ? emb[(size_t)ids[(size_t)g * ld_ids + i] * D + d0 + tx] : 0.0f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {       // feature d0 + r, nodes i0 + tx (coalesced along i)
        const int d = d0 + r, i = i0 + tx;
        if (i < ldk) {
            float h, l;
            split_tf32(tile[tx][r], h, l);
            const size_t o = ((size_t)g * D + d) * ldk + i;
            xt_hi[o] = h;
            xt_lo[o] = l;
        }
    }
}

__global__ void __launch_bounds__(256) split_kernel(const float *__restrict__ x, int64_t n, float *__restrict__ hi, float *__restrict__ lo)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        split_tf32(x[i], hi[i], lo[i]);
}

// pooled partials: partial[g, chunk, d] = sum over the chunk's nodes of H[g, i, d] * w[g, i]   (gnn.py:94-95)
__global__ void __launch_bounds__(256)
pool_rows_kernel(const float *__restrict__ H, const float *__restrict__ vertex_w, int ld_v, const int32_t *__restrict__ sizes,
                 int n_fixed, int D, int chunks, float *__restrict__ partial)
{
    const int g = blockIdx.y, chunk = blockIdx.x;
    const int n_g = sizes ? sizes[g] : n_fixed;
    const int per = ceil_div(n_fixed, chunks);
    const int r0 = chunk * per, r1 = min(n_g, r0 + per);
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.0f;
        for (int r = r0; r < r1; ++r) acc = fmaf(H[((size_t)g * n_fixed + r) * D + d], vertex_w[(size_t)g * ld_v + r], acc);
        partial[((size_t)g * chunks + chunk) * D + d] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------------------------
static size_t al256(size_t x) { return (x + 255) / 256 * 256; }

bool gnn_tc_supported(int D, int n_fixed)
{
    return D == G_BN && n_fixed >= 32 && encode_tiled_fn() != nullptr;
}

size_t gnn_tc_workspace_bytes(int G, int n_fixed, int D, int chunks)
{
    const size_t ldk = (size_t)(n_fixed + 3) / 4 * 4;
    const size_t adj = al256((size_t)G * n_fixed * ldk * 4);
    const size_t xt = al256((size_t)G * D * ldk * 4);
    const size_t y = al256((size_t)G * n_fixed * D * 4);
    const size_t w = al256((size_t)D * D * 4);
    return 2 * adj + 2 * xt + 2 * y + 2 * w + y /* H rows */ + al256((size_t)G * chunks * D * 4) + al256((size_t)G * D * 4) + 4096;
}

template <int EPI>
static int launch_gemm3x(const CUtensorMap *maps, const GemmTcArgs &a, const char *name, cudaStream_t st)
{
    static bool configured = false;
    if (!configured) {
        SH_CHECK_CUDA(cudaFuncSetAttribute(gemm3x_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
        configured = true;
    }
    const int tiles = a.G * ceil_div(a.M_total, G_BM);
    const int grid = min(tiles, sm_count());
    SH_LAUNCH(name, st, gemm3x_kernel<EPI><<<grid, G_THREADS, kSmemTotal, st>>>(maps[0], maps[1], maps[2], maps[3], a));
    SH_CHECK_LAUNCH();
    return 0;
}

// 3-D map over [batch, rows, cols]; a batch of 1 still uses rank 3 so that the kernel issues one kind of TMA
static int tmap3(CUtensorMap *m, const float *p, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t ld, uint64_t bstride,
                 uint32_t box_rows)
{
    EncodeTiledFn fn = encode_tiled_fn();
    SH_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available");
    cuuint64_t dims[3] = {cols, rows, batch};
    cuuint64_t strides[2] = {ld * 4, (batch > 1 ? bstride : rows * ld) * 4};
    cuuint32_t box[3] = {32, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(p), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SH_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

int gnn_forward_tc(const sh_gnn_params *p, int G, int n_fixed, const int32_t *sizes, const int64_t *ids,
                   const float *vertex_w, int ld_v, const float *edges, int64_t edge_batch_stride, int edge_ld, int chunks,
                   float *partial, void *workspace, cudaStream_t st)
{
    const int D = p->embed_dim;
    const int ldk = (n_fixed + 3) / 4 * 4;
    char *ws = (char *)workspace;
    const size_t adj_b = al256((size_t)G * n_fixed * ldk * 4), xt_b = al256((size_t)G * D * ldk * 4);
    const size_t y_b = al256((size_t)G * n_fixed * D * 4), w_b = al256((size_t)D * D * 4);
    float *adj_hi = (float *)ws; ws += adj_b;
    float *adj_lo = (float *)ws; ws += adj_b;
    float *xt_hi = (float *)ws; ws += xt_b;
    float *xt_lo = (float *)ws; ws += xt_b;
    float *y_hi = (float *)ws; ws += y_b;
    float *y_lo = (float *)ws; ws += y_b;
    float *w_hi = (float *)ws; ws += w_b;
    float *w_lo = (float *)ws; ws += w_b;
    float *h_rows = (float *)ws; ws += y_b;

    {
        dim3 grid(ceil_div(ldk, 32), ceil_div(n_fixed, 32), G);
        SH_LAUNCH("gnn_adj_prep", st, adj_prep_kernel<<<grid, 256, 0, st>>>(edges, edge_batch_stride, edge_ld, sizes, n_fixed, ldk, adj_hi, adj_lo));
        SH_CHECK_LAUNCH();
        dim3 grid2(ceil_div(ldk, 32), D / 32, G);
        SH_LAUNCH("gnn_embed_gather", st, embed_gather_t_kernel<<<grid2, 256, 0, st>>>(p->embedding, ids, ld_v, sizes, n_fixed, ldk, D, xt_hi, xt_lo));
        SH_CHECK_LAUNCH();
    }
    CUtensorMap adjm[2], xtm[2], ym[2], wm[2];
    if (tmap3(&adjm[0], adj_hi, n_fixed, n_fixed, G, ldk, (uint64_t)n_fixed * ldk, G_BM)) return 1;
    if (tmap3(&adjm[1], adj_lo, n_fixed, n_fixed, G, ldk, (uint64_t)n_fixed * ldk, G_BM)) return 1;
    if (tmap3(&xtm[0], xt_hi, n_fixed, D, G, ldk, (uint64_t)D * ldk, G_BN)) return 1;
    if (tmap3(&xtm[1], xt_lo, n_fixed, D, G, ldk, (uint64_t)D * ldk, G_BN)) return 1;
    if (tmap3(&ym[0], y_hi, D, (uint64_t)G * n_fixed, 1, D, 0, G_BM)) return 1;
    if (tmap3(&ym[1], y_lo, D, (uint64_t)G * n_fixed, 1, D, 0, G_BM)) return 1;
    if (tmap3(&wm[0], w_hi, D, D, 1, D, 0, G_BN)) return 1;
    if (tmap3(&wm[1], w_lo, D, D, 1, D, 0, G_BN)) return 1;

    for (int l = 0; l < p->num_layers; ++l) {
        const bool last = (l == p->num_layers - 1);
        SH_LAUNCH("gnn_split_weights", st, split_kernel<<<64, 256, 0, st>>>(p->lin_w[l], (int64_t)D * D, w_hi, w_lo));
        SH_CHECK_LAUNCH();
        // Y = Adj X
        GemmTcArgs a{};
        a.G = G; a.rows_per_graph = n_fixed; a.M_total = n_fixed; a.K_total = n_fixed; a.sizes = sizes; a.batched_b = 1;
        a.out_hi = y_hi; a.out_lo = y_lo;
        CUtensorMap m1[4] = {adjm[0], adjm[1], xtm[0], xtm[1]};
        if (launch_gemm3x<EPI_STORE_SPLIT>(m1, a, "gnn_adj_gemm_tc", st)) return 1;
        // H = relu(LN(Y W^T + b))
        GemmTcArgs b{};
        b.G = 1; b.rows_per_graph = n_fixed; b.M_total = G * n_fixed; b.K_total = D; b.sizes = sizes; b.batched_b = 0;
        b.bias = p->lin_b[l]; b.gamma = p->ln_w[l]; b.beta = p->ln_b[l]; b.eps = p->ln_eps;
        b.out_hi = xt_hi; b.out_lo = xt_lo; b.ldk = ldk; b.out_rows = h_rows;
        CUtensorMap m2[4] = {ym[0], ym[1], wm[0], wm[1]};
        if (last) { if (launch_gemm3x<EPI_LN_RELU_ROWS>(m2, b, "gnn_linear_ln_tc", st)) return 1; }
        else { if (launch_gemm3x<EPI_LN_RELU_T_SPLIT>(m2, b, "gnn_linear_ln_tc", st)) return 1; }
    }
    dim3 grid(chunks, G);
    SH_LAUNCH("gnn_pool_rows", st, pool_rows_kernel<<<grid, 256, 0, st>>>(h_rows, vertex_w, ld_v, sizes, n_fixed, D, chunks, partial));
    SH_CHECK_LAUNCH();
    return 0;
}

}  // namespace sh
