// gnn_tc.cu -- stage 3b on tcgen05 tensor cores: fp32-accurate batched GEMMs ("3 x fp16": three kind::f16 MMAs per
// product) with fused epilogues.
//
// Replaces the bmm / Linear / LayerNorm / ReLU chain of GraphConv + Layer (schema_inference/graph/gnn.py:20-46) for
// embed_dim % 256 == 0, embed_dim <= 1024 (256: everything fused as described below; wider: bias in the GEMM epilogue,
// LayerNorm + ReLU as a separate pass).  Other widths stay on the fp32 CUDA-core path (gnn.cu).
//
// Precision: the north star allows 1e-5 relative error on logits, which a single 16-bit or TF32 pass cannot meet.  Every
// fp32 operand x is scaled by a power of two s (exact) and split as  s x = hi + lo + r:
//     hi = (s x) & 0xffffe000   11 significant bits: exactly representable in fp16
//     lo = fp16_rn(s x - hi)    the next 11 bits;  |r| <= 2^-21 |s x|
// and each product is accumulated as  a b ~= a_lo b_hi + a_hi b_lo + a_hi b_hi  by three kind::f16 MMAs into one fp32 TMEM
// accumulator (products of fp16 values are exact in fp32; the dropped a_lo b_lo term is <= 2^-20 |a b|), i.e. the accuracy
// of the "3xTF32" scheme of round 1 at TWICE the tensor rate and half the shared-memory operand bytes per flop.
// s = 2^k with max |s x| in [2^14, 2^15): the maximum of every operand tensor is recorded on the device by the kernel that
// produces it (atomicMax on the bit pattern, `amax` slots in the workspace) and read by the consuming GEMM -- no host
// round trip, no overflow for any input range, and the epilogue undoes both scales with one exact multiplication.
//
// Operands live in HBM ONCE, as plain fp32.  TMA brings the fp32 tiles (32 k-elements x 128 rows, 128-byte swizzle) into a
// 32 KB stage; four transform warps convert them IN PLACE into four 8 KB fp16 tiles (A hi, A lo, B hi, B lo) in the
// un-swizzled K-major core-matrix layout (8 rows x 16 bytes contiguous, LBO 128 B, SBO 512 B) and release the MMA warp.
// Half the bytes per stage of the hi/lo fp32 scheme -> a 6-deep ring instead of 3, which is what hides the TMA latency.
//
// Per layer, for a batch of G graphs with n_g <= n_fixed nodes:
//   adj GEMM     Y[g]  = Adj[g] (n x n, symmetric, K-major)  *  X[g]   given as X^T [D, n] (K-major)   -> Y  [n, D]
//   linear GEMM  Z     = Y (rows x D, K-major) * W^T with W [D_out, D_in] (K-major), + bias, LayerNorm, ReLU fused in
//                the TMEM epilogue (one thread owns one full 256-wide row: LayerNorm needs no cross-thread reduction)
//                -> H^T [D, n] for the next layer's adj GEMM, or H [rows, D] for the pooling.
// Layer 0 is shortened to ONE adjacency GEMM with bias + LayerNorm + ReLU in its epilogue (the first Linear is applied to
// the (M+1)-row embedding table instead, see run_layers_tc); the last layer's epilogue emits vertex-weighted group sums
// instead of activations.  Class graphs run on their un-pruned vertices only (class_perm_kernel, per-code tables).
// Kernel layout: warp 0 TMA producer, warp 1 MMA issuer / TMEM owner, warps 2-5 epilogue, warps 6-9 operand transform;
// CTA pairs (cta_group::2, one UMMA of M = 256 per pair, each CTA stages and converts its 128 A rows and half of the B
// tile), or single CTAs (SCHEMANET_GEMM_CTAS=1, 48 KB stages, 4-deep); 2-stage 256-column TMEM accumulator ring.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "gnn_tc.cuh"
#include "tc_common.cuh"

namespace sh {

using namespace tc;

constexpr int G_BM = 128;
constexpr int G_BN = 256;      // == embed_dim
constexpr int G_BK = 32;
constexpr int kGatherMaxNodes = 1024;                      // node codes of one graph held in shared memory by a gathering GEMM
// Warp budget: producer + MMA issuer + EW epilogue warps + 8 operand-transform warps; EW is a template parameter of the kernel.
// EW = 8 (two warps per TMEM lane quarter, each owns half of a tile's columns; 576 threads, 96 registers, one stage less): the
// embed_dim = 256 launches, whose 6-8 k-block tiles are bound by the LayerNorm epilogue.  EW = 4 (448 threads, 128 registers):
// everything else -- at cfg4 the transform also applies LayerNorm and the register cap of the larger block costs more than the
// second epilogue warpgroup saves.  Same-box A/B on B200 (profiles/r02_experiments.md): cfg2 0.685 -> 0.665 ms, cfg3 1.019 ->
// 0.996 ms with EW = 8; cfg4 14.78 -> 15.05 ms, hence 4 there.  (8 epilogue + 4 transform warps: no gain at cfg2, -4 % at cfg4.)
constexpr int G_XF_WARPS = 8;                              // operand-transform warps
constexpr int kABytes = G_BM * G_BK * 4;                 // 16 KB

enum { EPI_STORE_ROWS = 0, EPI_LN_RELU_T = 1, EPI_LN_RELU_ROWS = 2, EPI_BIAS_ROWS = 3,
       // embed_dim > 256: a LayerNorm row spans several 256-column accumulator tiles, which different work units own.  The
       // epilogue stores z = acc + bias un-normalised (transposed for the next adjacency GEMM, or as rows for the pooling)
       // together with the tile's (mean, M2) of every row; a tiny kernel merges the tiles' statistics (Chan's formula) and the
       // CONSUMER normalises on the fly: the transform warps of the next adjacency GEMM while they convert the B operand, or
       // the pooling kernel.  No separate LayerNorm pass over the activations, no round trip of H through HBM.
       EPI_Z_T_STATS = 4, EPI_Z_ROWS_STATS = 5 };
constexpr int kMaxDim = 1024;   // largest embed_dim the bias staging buffer holds

struct GemmTcArgs {
    int G;                 // batch of graphs (adj GEMM) or 1 (linear GEMM over flattened rows)
    int rows_per_graph;    // n_fixed
    int M_total;           // rows of the A operand per batch entry
    int K_total;           // reduction length upper bound
    int N_total;           // output columns (embed_dim): tiles of 256 columns, ld of the row-major outputs
    const int32_t *k_sizes;   // adj GEMM: [G] active size n_g of each graph (rows and reduction range) or null
    const int32_t *row_sizes; // epilogue: [graphs] number of real rows per graph (masks LayerNorm outputs) or null
    int identity_tail;        // adj GEMM: rows >= n_g carry an identity diagonal (class graphs compacted to their
                              // un-pruned vertices): such row blocks only visit their own diagonal k-blocks
    int batched_b;            // 1: B operand indexed by the graph, 0: shared (weights)
    int skip_masked;          // 1: work units whose rows are all >= their graph's size are skipped in every epilogue
                              // (nothing downstream reads those rows: class graphs reduced to their un-pruned vertices)
    // epilogue
    float *out_t;                // EPI_LN_RELU_T: H^T [G, 256, ldk]
    float *out_rows;             // EPI_STORE_ROWS: Y [G, n_fixed, N_total]; EPI_LN_RELU_ROWS / EPI_BIAS_ROWS: H / Z [rows, N_total]
    // EPI_LN_RELU_ROWS on flattened rows (G == 1) with pool_groups != null: instead of storing H, emit the vertex-weighted
    // column sums of every 32-row group, split at the (at most one) graph boundary inside the group:
    // pool_groups[group, 0 | 1, 256]  (slot 0: rows of the group's first graph, slot 1: rows of the next graph)
    const float *pool_w;         // [graphs, ld_w] vertex weights
    int ld_w;
    float *pool_groups;
    int ldk;                     // row stride of the transposed output
    const float *bias, *gamma, *beta;
    float eps;
    // dynamic operand scaling: bit patterns of max |A|, max |B| (written by the kernels that produced the operands) and the
    // slot this launch records the maximum of its own output in (null: the output is not a GEMM operand)
    const unsigned *amax_a, *amax_b;
    unsigned *amax_out;
    int n_valid;                 // EPI_BIAS_ROWS: output columns that exist (0: all N_total); the row stride is n_valid then
    float *stats;                // EPI_Z_*_STATS: [rows, N_total / 256, 2] per-tile (mean, M2) of z
    // B operand = relu(LayerNorm(z)) of the stored z^T, applied by the transform warps (null: B is used as stored):
    const float *bln_mr;         // [G * rows_per_graph, 2] merged (mean, rstd) per node
    const float *bln_gamma, *bln_beta;   // [N_total]
    // B operand GATHERED instead of loaded by a tiled TMA (layer 0: X_0^T[f, k] = table[ids[g, k], f], so the [G, D, n] copy of
    // the gathered table rows is never written): the producer warp issues one bulk copy per node (its slice of a table row,
    // completing on the stage's mbarrier like a TMA box); the landing tile is [k][f] and the transform warps read it
    // transposed.  bg_table [rows, N_total] fp32, bg_ids [G, bg_ld] node codes, bg_sizes [G] live nodes per graph (null:
    // rows_per_graph); nodes beyond read bg_zero_row, a table row that is all zeros (written by tables_begin behind the table)
    const float *bg_table;
    const int64_t *bg_ids;
    const int32_t *bg_sizes;
    int bg_ld, bg_zero_row;
    int debug;      // timing experiments (SCHEMANET_GEMM_DEBUG): 1 transform skips its work, 2 no MMAs, 4 epilogue skips its work
    long long *trace;   // SCHEMANET_GEMM_TRACE: per CTA 8 cycle counters (see launch_gemm3x_n), null in normal runs
};
#define TR_T0() const long long tr_t0_ = a.trace ? clock64() : 0
#define TR_ADD(var) do { if (a.trace) var += clock64() - tr_t0_; } while (0)

// power-of-two scale that puts a tensor whose largest magnitude has the bit pattern `bits` into [2^14, 2^15)
__device__ __forceinline__ float scale_from_amax(unsigned bits)
{
    int e = (int)((bits >> 23) & 0xffu);                 // max < 2^(e - 126)
    if (bits == 0u) return 1.0f;
    int k = 141 - e;                                     // 2^k * max < 2^15
    k = k > 60 ? 60 : k;                                 // (tensors below 2^-45 are not scaled further: no overflow of s_a s_b)
    return __uint_as_float((unsigned)(127 + k) << 23);
}

__device__ __forceinline__ void record_amax(unsigned *slot, float m)
{
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && slot != nullptr && !(m == 0.0f)) atomicMax(slot, __float_as_uint(m));   // NaN compares above +inf
}

__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo)
{
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// A warp holds a 32 (rows, one per lane) x 32 (columns) fp32 chunk in registers (the TMEM layout).  Writing it row-major
// straight from registers would issue 16-byte pieces of 32 different rows per instruction; going through a padded
// shared-memory tile lets 8 lanes cover one 128-byte row segment, i.e. every store instruction writes 4 full lines.
__device__ __forceinline__ void store_chunk_rows(float *tile /* [32][33] */, const float (&v)[32], int lane, float *dst,
                                                 size_t row_stride, int rows_valid)
{
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = v[j];
    __syncwarp();
    const int sub = lane >> 3, c4 = (lane & 7) * 4;   // 4 rows per instruction, 8 lanes x float4 per row
#pragma unroll
    for (int r0 = 0; r0 < 32; r0 += 4) {
        const int r = r0 + sub;
        if (r < rows_valid) {
            const float *t = tile + r * 33 + c4;
            *reinterpret_cast<float4 *>(dst + (size_t)r * row_stride + c4) = make_float4(t[0], t[1], t[2], t[3]);
        }
    }
}

// Shared-memory matrix descriptor for the fp16 tiles the transform warps write: K-major, no swizzle ("interleaved"): a core
// matrix is 8 rows x 16 bytes, contiguous; the two 16-byte k-chunks of one K = 16 step are LBO = 128 bytes apart, 8-row
// groups SBO = 512 bytes apart (cute::UMMA make_umma_desc<Major::K>, LayoutType::INTERLEAVE; checked on the device by
// tools/probe/umma_probe.cu).  Chunk (row r, k-chunk c of 4) of a tile sits at (r >> 3) * 512 + c * 128 + (r & 7) * 16.
__device__ __forceinline__ uint64_t make_desc_k_interleaved(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           ((uint64_t)1 << 46);
}

__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 x;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr));
    return x;
}
__device__ __forceinline__ float lds32(uint32_t addr)
{
    float x;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
    return x;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// two scaled fp32 values -> packed fp16 pair of their 11-bit heads (exact) and of the rounded remainders
__device__ __forceinline__ void split_pair(float x0, float x1, float s, uint32_t &hi, uint32_t &lo)
{
    x0 *= s; x1 *= s;
    const float h0 = __uint_as_float(__float_as_uint(x0) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(x1) & 0xffffe000u);
    const __half2 hh = __floats2half2_rn(h0, h1), ll = __floats2half2_rn(x0 - h0, x1 - h1);
    hi = *reinterpret_cast<const uint32_t *>(&hh);
    lo = *reinterpret_cast<const uint32_t *>(&ll);
}
// half of a 128-byte landing row (4 float4 in k order = 16 k-elements) -> its 2 hi and 2 lo 16-byte chunks
__device__ __forceinline__ void convert_half_row(const float4 (&v)[4], float s, uint32_t hi_base, uint32_t lo_base)
{
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t h[4], l[4];
        split_pair(v[2 * c].x, v[2 * c].y, s, h[0], l[0]);
        split_pair(v[2 * c].z, v[2 * c].w, s, h[1], l[1]);
        split_pair(v[2 * c + 1].x, v[2 * c + 1].y, s, h[2], l[2]);
        split_pair(v[2 * c + 1].z, v[2 * c + 1].w, s, h[3], l[3]);
        sts128(hi_base + c * 128, h[0], h[1], h[2], h[3]);
        sts128(lo_base + c * 128, l[0], l[1], l[2], l[3]);
    }
}

// Shared-memory plan.  A stage is the fp32 landing area of one k-block: 128 A rows (16 KB) + this CTA's B rows (CTA pair:
// 128 rows, 16 KB; single CTA: 256 rows, 32 KB).  The conversion is in place: each region ends up as [hi tile | lo tile].
template <int CTAS, int EW>
struct GemmPlan {
    static constexpr int kBRows = G_BN / CTAS;
    static constexpr int kBBytesL = kBRows * G_BK * 4;            // landing bytes of the B rows = hi tile + lo tile
    static constexpr int kBTile = kBBytesL / 2;
    static constexpr int kStage = kABytes + kBBytesL;
    static constexpr int kStages = EW == 4 ? (CTAS == 1 ? 4 : 6) : (CTAS == 1 ? 3 : 5);
    static constexpr int kBar = kStages * kStage;
    static constexpr int kParam = kBar + 256;
    static constexpr int kStageOut = kParam + (2 * G_BN + kMaxDim) * 4;   // gamma, beta (LN: 256 wide) + bias (up to kMaxDim)
    static constexpr int kXch = kStageOut + EW * 32 * 33 * 4;      // LayerNorm partials exchanged by the two column halves
    static constexpr int kIds = kXch + 2 * 2 * G_BM * 8;                    // node codes of the current graph (gathered B operand)
    static constexpr int kTotal = kIds + 1024 * 4 + 1024;            // (kGatherMaxNodes int32)
};

template <int EPI, int CTAS, int EW>
__global__ void __launch_bounds__(32 * (2 + EW + G_XF_WARPS), 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmTcArgs a)
{
    constexpr int G_EPI_WARPS = EW;
    constexpr int G_XF_FIRST = 2 + G_EPI_WARPS;                // first transform warp
    constexpr int G_THREADS = 32 * (G_XF_FIRST + G_XF_WARPS);
    using P = GemmPlan<CTAS, EW>;
    constexpr int S = P::kStages;
    const long long tr_entry = a.trace ? clock64() : 0;
    unsigned long long tr_gt0 = 0;
    if (a.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_gt0));
    extern __shared__ uint8_t smem_raw[];
    // 1 KB alignment for the 128-byte-swizzled TMA tiles, as an OFFSET into the shared array: going through uintptr_t makes
    // the compiler lose the address space and emit 64-bit generic LD/ST for every shared-memory access of the epilogue
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = (uint64_t *)(smem + P::kBar);       // TMA bytes of this CTA's slot have landed (local)
    uint64_t *ready = full + S;                          // both CTAs' slots are converted to fp16 hi/lo (lives in the leader)
    uint64_t *empty = ready + S;
    uint64_t *tmem_full = empty + S;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_ptr = (uint32_t *)(tmem_empty + 2);
    float *s_gamma = (float *)(smem + P::kParam), *s_beta = s_gamma + G_BN, *s_bias = s_beta + G_BN;
    float *s_out = (float *)(smem + P::kStageOut);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = CTAS == 2 ? (int)cluster_ctarank() : 0;     // 0 = the CTA that issues the MMAs
    const int unit = (int)blockIdx.x / CTAS, num_units = (int)gridDim.x / CTAS;
    const float scale_a = scale_from_amax(__ldg(a.amax_a)), scale_b = scale_from_amax(__ldg(a.amax_b));
    const float unscale = (1.0f / scale_a) * (1.0f / scale_b);      // powers of two: exact

    if (EPI == EPI_LN_RELU_T || EPI == EPI_LN_RELU_ROWS)
        for (int i = threadIdx.x; i < G_BN; i += G_THREADS) { s_bias[i] = a.bias[i]; s_gamma[i] = a.gamma[i]; s_beta[i] = a.beta[i]; }
    if (EPI == EPI_BIAS_ROWS || EPI == EPI_Z_T_STATS || EPI == EPI_Z_ROWS_STATS)
        for (int i = threadIdx.x; i < a.N_total; i += G_THREADS) s_bias[i] = a.bias ? a.bias[i] : 0.0f;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], G_XF_WARPS * CTAS); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], G_EPI_WARPS * CTAS); }
        fence_barrier_init();
    }
    if (CTAS == 2) cluster_sync();       // the peer's barriers must exist before anything signals them
    if (warp == 1) { if (CTAS == 2) tmem_alloc_pair(tmem_ptr, 2 * G_BN); else tmem_alloc(tmem_ptr, 2 * G_BN); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    constexpr int UB = G_BM * CTAS;                      // rows per work unit (a CTA, or a CTA pair)
    const int ub_per = ceil_div(a.M_total, UB);
    const int nb_per = a.N_total / G_BN;                 // 256-column output tiles (1 unless embed_dim > 256)
    const int tiles = a.G * ub_per * nb_per;

    // every role of every CTA walks the same unit list and applies the same skip rule
#define TILE_LOOP_BEGIN                                                                                   \
    for (int t = unit; t < tiles; t += num_units) {                                                       \
        const int nb = t % nb_per, gu = t / nb_per;   /* column tile fastest: neighbours share the A rows */ \
        const int g = gu / ub_per, ub = gu % ub_per;                                                      \
        const int mb = ub * CTAS + rank;   /* this CTA's 128-row block */                                 \
        const int n_g = (a.k_sizes && a.G > 1) ? a.k_sizes[g] : a.K_total;                                \
        /* units past the graph are skipped, except when the epilogue must still write their (zero) rows */ \
        if ((EPI == EPI_STORE_ROWS || a.skip_masked) && a.G > 1 && !a.identity_tail && ub * UB >= n_g) continue; \
        if (a.skip_masked && a.G == 1 && a.row_sizes) {   /* flattened rows: unit inside one graph, past its size */ \
            const int r0_ = ub * UB, gi_ = r0_ / a.rows_per_graph;                                        \
            if ((r0_ + UB - 1) / a.rows_per_graph == gi_ && r0_ - gi_ * a.rows_per_graph >= a.row_sizes[gi_]) continue; \
        }                                                                                                 \
        /* k-blocks [0, kA) cover the active range; with identity_tail a unit that holds rows >= n_g       \
           additionally visits the k-blocks of its own diagonal that are not in [0, kA) */                 \
        const int kA = (a.G > 1 && a.identity_tail && ub * UB >= n_g) ? 0 : max(1, ceil_div(n_g, G_BK));   \
        int k2s = 0, k2c = 0;                                                                             \
        if (a.identity_tail && (ub + 1) * UB > n_g) {                                                     \
            k2s = max(kA, ub * (UB / G_BK));                                                              \
            k2c = max(0, min((ub + 1) * (UB / G_BK), ceil_div(a.K_total, G_BK)) - k2s);                    \
        }                                                                                                 \
        const int kblocks = kA + k2c;                                                                     \
        if (kblocks == 0) continue;
#define TILE_LOOP_END }

    long long tr0 = 0, tr1 = 0;                            // (trace: cycles this role spent waiting / working)
    const long long tr_start = a.trace ? clock64() : 0;
    if (warp == 0) {
        if (a.bg_table != nullptr) {
            // ===================== producer, gathered B: whole warp =====================
            int *s_ids = reinterpret_cast<int *>(smem + P::kIds);
            int stage = 0;
            uint32_t phase = 0;
            TILE_LOOP_BEGIN
                const int n_live = a.bg_sizes ? a.bg_sizes[g] : a.rows_per_graph;
                __syncwarp();
                for (int i = lane; i < kGatherMaxNodes; i += 32)          // table row of every node of graph g
                    s_ids[i] = i < n_live ? (int)a.bg_ids[(size_t)g * a.bg_ld + i] : a.bg_zero_row;
                __syncwarp();
                for (int kidx = 0; kidx < kblocks; ++kidx) {
                    const int kb = kidx < kA ? kidx : k2s + (kidx - kA);
                    uint8_t *s = smem + stage * P::kStage;
                    if (lane == 0) {
                        { TR_T0(); mbar_wait(&empty[stage], phase ^ 1); TR_ADD(tr0); }
                        mbar_arrive_expect_tx(&full[stage], P::kStage);
                        tma_load_3d(s, &tmA, &full[stage], kb * G_BK, mb * G_BM, g);
                    }
                    __syncwarp();                                           // the slot is free: every lane may write into it
                    if (a.debug & 8) {
                        // (A/B, SCHEMANET_GEMM_DEBUG=8) one 1-D bulk copy per node: 32 copy operations per k-block; the
                        // producer, not the conversion, then bounds the launch (r02_experiments.md)
                        const int k = kb * G_BK + lane;
                        const int row = k < kGatherMaxNodes ? s_ids[k] : a.bg_zero_row;
                        const float *src = a.bg_table + (size_t)row * a.N_total + nb * G_BN + rank * P::kBRows;
                        bulk_copy_g2s(s + kABytes + lane * (P::kBRows * 4), src, P::kBRows * 4, &full[stage]);
                    } else if (lane < G_BK / 4) {
                        // TMA gather: four table rows per operation (tmB is the 2-D map of the table, box {kBRows, 1})
                        const int k = kb * G_BK + 4 * lane;
                        int r[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) r[u] = k + u < kGatherMaxNodes ? s_ids[k + u] : a.bg_zero_row;
                        tma_gather4_2d(s + kABytes + 4 * lane * (P::kBRows * 4), &tmB, &full[stage], nb * G_BN + rank * P::kBRows,
                                       r[0], r[1], r[2], r[3]);
                    }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            TILE_LOOP_END
            if (a.trace && lane == 0) a.trace[blockIdx.x * 8 + 0] = tr0;
        } else if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            TILE_LOOP_BEGIN
                for (int kidx = 0; kidx < kblocks; ++kidx) {
                    const int kb = kidx < kA ? kidx : k2s + (kidx - kA);
                    { TR_T0(); mbar_wait(&empty[stage], phase ^ 1); TR_ADD(tr0); }
                    uint8_t *s = smem + stage * P::kStage;
                    const int gb = a.batched_b ? g : 0;
                    // every CTA loads its own fp32 tiles (128 A rows, its share of the B tile) and tracks them on its OWN
                    // barrier: its transform warps wait there
                    mbar_arrive_expect_tx(&full[stage], P::kStage);
                    tma_load_3d(s, &tmA, &full[stage], kb * G_BK, mb * G_BM, g);
                    tma_load_3d(s + kABytes, &tmB, &full[stage], kb * G_BK, nb * G_BN + rank * P::kBRows, gb);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            TILE_LOOP_END
            if (a.trace) a.trace[blockIdx.x * 8 + 0] = tr0;
        }
    } else if (warp == 1) {
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_f16(G_BM * CTAS, G_BN);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            TILE_LOOP_BEGIN
                { TR_T0(); mbar_wait(&tmem_empty[as], aphase ^ 1); TR_ADD(tr1); }
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * G_BN);
                for (int kb = 0; kb < kblocks; ++kb) {
                    { TR_T0(); mbar_wait(&ready[stage], phase); TR_ADD(tr0); }     // fp16 hi/lo tiles of both CTAs' slots are in place
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t s = smem_u32(smem + stage * P::kStage);
                        const uint64_t ah = make_desc_k_interleaved(s), al = make_desc_k_interleaved(s + kABytes / 2);
                        const uint64_t bh = make_desc_k_interleaved(s + kABytes), bl = make_desc_k_interleaved(s + kABytes + P::kBTile);
#pragma unroll
                        for (int k = 0; k < ((a.debug & 2) ? 0 : G_BK / 16); ++k) {          // one UMMA consumes 16 halves of K: two 16-byte chunks, 256 bytes on
                            const uint64_t o = (uint64_t)(16 * k);
                            if (CTAS == 1) {
                                umma_f16(tmem_d, al + o, bh + o, idesc, (kb | k) != 0);   // small terms first
                                umma_f16(tmem_d, ah + o, bl + o, idesc, 1);
                                umma_f16(tmem_d, ah + o, bh + o, idesc, 1);
                            } else {
                                umma_f16_pair(tmem_d, al + o, bh + o, idesc, (kb | k) != 0);
                                umma_f16_pair(tmem_d, ah + o, bl + o, idesc, 1);
                                umma_f16_pair(tmem_d, ah + o, bh + o, idesc, 1);
                            }
                        }
                        if (CTAS == 1) {
                            umma_commit(&empty[stage]);
                            if (kb == kblocks - 1) umma_commit(&tmem_full[as]);
                        } else {
                            umma_commit_pair(&empty[stage], 3);                    // frees the slot in both CTAs
                            if (kb == kblocks - 1) umma_commit_pair(&tmem_full[as], 3);
                        }
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                if (++as == 2) { as = 0; aphase ^= 1; }
            TILE_LOOP_END
            if (a.trace && lane == 0) { a.trace[blockIdx.x * 8 + 3] = tr0; a.trace[blockIdx.x * 8 + 4] = tr1; }
        }
    } else if (warp < G_XF_FIRST) {
        // ===================== epilogue: two warps per TMEM lane quarter, each owns half of the tile's 8 column chunks =====================
        const int wq = warp & 3;
        constexpr int kCPW = (G_BN / 32) / (G_EPI_WARPS / 4);                   // column chunks per warp
        const int half = (warp - 2) >> 2;
        const int c_lo = half * kCPW, c_hi = c_lo + kCPW;
        const int row_in_tile = wq * 32 + lane;
        float *s_tile = s_out + (wq + 4 * half) * 32 * 33;                     // this warp's staging tile
        float2 *s_xch = reinterpret_cast<float2 *>(smem + P::kXch);             // [accumulator][half][row]: (mean, M2) of a half row
        // the two warps of a quarter combine their halves' statistics (equal counts n = 128: Chan et al.)
        auto combine_halves = [&](int as_, float mean_h, float m2_h, float &mean, float &m2) {
            if (G_EPI_WARPS == 4) { mean = mean_h; m2 = m2_h; return; }          // one warp owns the whole row
            s_xch[(as_ * 2 + half) * G_BM + row_in_tile] = make_float2(mean_h, m2_h);
            asm volatile("bar.sync %0, 64;" ::"r"(3 + wq) : "memory");
            const float2 o = s_xch[(as_ * 2 + (half ^ 1)) * G_BM + row_in_tile];
            const float dlt = mean_h - o.x;
            mean = 0.5f * (mean_h + o.x);
            m2 = m2_h + o.y + (float)(G_BN / 4) * dlt * dlt;
        };
        int as = 0;
        uint32_t aphase = 0;
        TILE_LOOP_BEGIN
            { TR_T0(); mbar_wait(&tmem_full[as], aphase); TR_ADD(tr0); }
            const long long tr_w0 = a.trace ? clock64() : 0;
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * G_BN);
            const int m = mb * G_BM + row_in_tile;           // row inside this batch entry
            float out_max = 0.0f;                            // largest magnitude this thread writes (the next GEMM's operand scale)
            if (a.debug & 4) {
            } else if (EPI == EPI_STORE_ROWS) {
                // Y[g, m, :] (row-major: the K-major A operand of the linear GEMM)
                // warp-level view: this warp owns rows [mb*128 + wq*32, +32) of graph g
                const int m_warp = mb * G_BM + wq * 32;
                const int rows_valid = min(32, a.rows_per_graph - m_warp);
                float *o = a.out_rows + ((size_t)g * a.rows_per_graph + m_warp) * a.N_total + nb * G_BN;
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) { v[j] *= unscale; out_max = fmaxf(out_max, fabsf(v[j])); }
                    store_chunk_rows(s_tile, v, lane, o + c * 32, a.N_total, rows_valid);
                }
            } else if (EPI == EPI_BIAS_ROWS) {
                // Z[m, nb*256 + :] = acc + bias (embed_dim > 256: LayerNorm needs the whole row and runs as its own kernel)
                const int m_warp = mb * G_BM + wq * 32;
                float *o = a.out_rows + (size_t)m_warp * a.N_total + nb * G_BN;
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], unscale, s_bias[nb * G_BN + c * 32 + j]);
                    if (a.amax_out != nullptr && m < a.M_total) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) out_max = fmaxf(out_max, fabsf(v[j]));
                    }
                    if (a.n_valid == 0) {
                        store_chunk_rows(s_tile, v, lane, o + c * 32, a.N_total, min(32, a.M_total - m_warp));
                    } else if (m < a.M_total) {
                        // ragged output width (logits [B, K]): this thread's row, the columns that exist
                        float *orow = a.out_rows + (size_t)m * a.n_valid + nb * G_BN + c * 32;
                        const int cols = a.n_valid - (nb * G_BN + c * 32);
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < cols) orow[j] = v[j];
                    }
                }
            } else if (EPI == EPI_Z_T_STATS || EPI == EPI_Z_ROWS_STATS) {
                // z = acc + bias of this 256-column tile (masked rows: z = 0, gnn.py:43-44) + the tile's LayerNorm statistics
                const int gg = a.G > 1 ? g : m / a.rows_per_graph, i = a.G > 1 ? m : m % a.rows_per_graph;
                const bool in_range = a.G > 1 ? m < a.rows_per_graph : m < a.M_total;
                const int n_node = (in_range && a.row_sizes) ? a.row_sizes[gg] : a.rows_per_graph;
                const bool valid = in_range && i < n_node;
                const float *bias_t = s_bias + nb * G_BN;
                float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f}, shift = 0.0f;
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = valid ? fmaf(v[j], unscale, bias_t[c * 32 + j]) : 0.0f;
                    if (c == c_lo) {
                        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) { q0 += v[j]; q1 += v[j + 1]; q2 += v[j + 2]; q3 += v[j + 3]; }
                        shift = ((q0 + q1) + (q2 + q3)) * (1.0f / 32.0f);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float tt = v[j] - shift;
                        p1[j & 3] += tt;
                        p2[j & 3] = fmaf(tt, tt, p2[j & 3]);
                    }
                    if (EPI == EPI_Z_T_STATS) {
                        if (in_range) {
                            float *o = a.out_t + ((size_t)gg * a.N_total + nb * G_BN + c * 32) * a.ldk + i;
#pragma unroll
                            for (int j = 0; j < 32; ++j) o[(size_t)j * a.ldk] = v[j];
                        }
                    } else {
                        const int m_warp = mb * G_BM + wq * 32;
                        const size_t row0 = (a.G > 1 ? (size_t)g * a.rows_per_graph : 0) + m_warp;
                        store_chunk_rows(s_tile, v, lane, a.out_rows + row0 * a.N_total + nb * G_BN + c * 32, a.N_total,
                                         min(32, (a.G > 1 ? a.rows_per_graph : a.M_total) - m_warp));
                    }
                }
                {
                    const float s1 = (p1[0] + p1[1]) + (p1[2] + p1[3]), s2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
                    const float dm = s1 / (float)(32 * kCPW);
                    float2 st;                                  // mean of the tile's 256 values, sum of squared deviations from it
                    combine_halves(as, shift + dm, fmaxf(s2 - s1 * dm, 0.0f), st.x, st.y);
                    if (in_range && half == 0)
                        *reinterpret_cast<float2 *>(a.stats + (((size_t)gg * a.rows_per_graph + i) * nb_per + nb) * 2) = st;
                }
            } else {
                // z = acc + bias; LayerNorm over the 256 columns this thread owns; ReLU   (gnn.py:31,45)
                // linear GEMM: flattened rows -> (graph, node); adjacency GEMM with a fused LayerNorm: (g, row in graph)
                const int gg = a.G > 1 ? g : m / a.rows_per_graph, i = a.G > 1 ? m : m % a.rows_per_graph;
                const bool in_range = a.G > 1 ? m < a.rows_per_graph : m < a.M_total;
                const int n_node = (in_range && a.row_sizes) ? a.row_sizes[gg] : a.rows_per_graph;
                const bool valid = in_range && i < n_node;
                // mean and variance in ONE pass over the accumulator row (each pass is 8 tcgen05.ld + 256 adds per thread, and
                // the epilogue, not the MMA, bounds these kernels): sums of (z - K) and (z - K)^2 with the shift K = the mean of
                // the row's first 32 features, so that var = S2/n - (S1/n)^2 subtracts quantities of the size of the variance
                // (shifted-data algorithm; ~2e-7 relative, the same as the two-pass form at this width)
                float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f}, shift = 0.0f;   // 4 independent chains each
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
                    if (c == c_lo) {       // shift = mean of this half's first 32 features (robust against a single outlier)
                        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            q0 += fmaf(v[j], unscale, s_bias[c * 32 + j]); q1 += fmaf(v[j + 1], unscale, s_bias[c * 32 + j + 1]);
                            q2 += fmaf(v[j + 2], unscale, s_bias[c * 32 + j + 2]); q3 += fmaf(v[j + 3], unscale, s_bias[c * 32 + j + 3]);
                        }
                        shift = ((q0 + q1) + (q2 + q3)) * (1.0f / 32.0f);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float tt = fmaf(v[j], unscale, s_bias[c * 32 + j]) - shift;
                        p1[j & 3] += tt;
                        p2[j & 3] = fmaf(tt, tt, p2[j & 3]);
                    }
                }
                const float s1 = (p1[0] + p1[1]) + (p1[2] + p1[3]), s2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
                const float dm = s1 / (float)(32 * kCPW);
                float mean, var;                                     // of the whole 256-wide row: var = sum (z - mean)^2
                combine_halves(as, shift + dm, fmaxf(s2 - s1 * dm, 0.0f), mean, var);
                const float rstd = 1.0f / sqrtf(var / (float)G_BN + a.eps);
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = c * 32 + j;
                        const float z = fmaf(v[j], unscale, s_bias[n]);
                        v[j] = fmaxf((z - mean) * rstd * s_gamma[n] + s_beta[n], 0.0f);
                    }
                    if (EPI == EPI_LN_RELU_ROWS && a.pool_groups != nullptr) {
                        // fused weighted pooling (gnn.py:94-95): the activations never go to memory
                        const int m_warp = mb * G_BM + wq * 32;                       // flattened row of lane 0
                        const int g0 = m_warp / a.rows_per_graph;
                        const int split = min(32, (g0 + 1) * a.rows_per_graph - m_warp);   // rows of the group's first graph
                        const float wrow = valid ? __ldg(a.pool_w + (size_t)gg * a.ld_w + i) : 0.0f;
                        float *tile = s_tile;
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = v[j] * wrow;
                        __syncwarp();
                        float s0 = 0.0f, s1 = 0.0f;
                        if (split == 32) {       // usual case: one graph in the group; four independent chains, fixed order
                            float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
#pragma unroll
                            for (int r = 0; r < 32; r += 4) {
                                p0 += tile[r * 33 + lane];
                                p1 += tile[(r + 1) * 33 + lane];
                                p2 += tile[(r + 2) * 33 + lane];
                                p3 += tile[(r + 3) * 33 + lane];
                            }
                            s0 = (p0 + p1) + (p2 + p3);
                        } else {
                            for (int r = 0; r < split; ++r) s0 += tile[r * 33 + lane];
                            for (int r = split; r < 32; ++r) s1 += tile[r * 33 + lane];
                        }
                        if (m_warp < a.M_total) {
                            float *pg = a.pool_groups + (size_t)(m_warp / 32) * 2 * G_BN + c * 32 + lane;
                            pg[0] = s0;
                            pg[G_BN] = s1;
                        }
                    } else if (EPI == EPI_LN_RELU_ROWS) {
                        // rows of masked nodes are never read downstream (pooling stops at n_g): store all rows in range
                        const int m_warp = mb * G_BM + wq * 32;
                        const size_t row0 = (a.G > 1 ? (size_t)g * a.rows_per_graph : 0) + m_warp;
                        store_chunk_rows(s_tile, v, lane, a.out_rows + row0 * G_BN + c * 32, G_BN,
                                         min(32, (a.G > 1 ? a.rows_per_graph : a.M_total) - m_warp));
                    } else {
                        // H^T[gg, n, i]: lanes hold consecutive nodes i -> coalesced 128-byte stores; nodes beyond n_g
                        // are written as zeros (they are the zero-padded K range of the next adj GEMM)
                        if (in_range) {
                            float *o = a.out_t + ((size_t)gg * G_BN + c * 32) * a.ldk + i;
#pragma unroll
                            for (int j = 0; j < 32; ++j) o[(size_t)j * a.ldk] = valid ? v[j] : 0.0f;
                        }
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) out_max = fmaxf(out_max, v[j]);      // (after ReLU: v >= 0)
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (EPI == EPI_STORE_ROWS || EPI == EPI_LN_RELU_T || EPI == EPI_BIAS_ROWS) record_amax(a.amax_out, out_max);
            if (a.trace) tr1 += clock64() - tr_w0;
            if (lane == 0) {
                if (CTAS == 1) mbar_arrive(&tmem_empty[as]);
                else mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0));   // the leader owns the accumulator ring
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        TILE_LOOP_END
        if (a.trace && threadIdx.x == 64) { a.trace[blockIdx.x * 8 + 5] = tr0; a.trace[blockIdx.x * 8 + 6] = tr1; }
    } else {
        // ===================== operand transform: fp32 landing tiles -> fp16 hi / lo tiles, in place =====================
        // Thread t owns landing row (t & 127) of A and of B (+128) -- all of it with four transform warps, the half
        // k = 16 h .. 16 h + 15 (h = t >> 7) with eight.  It reads its 16-byte pieces (piece c of row r was stored by TMA at
        // position c ^ (r & 7); a quarter-warp touches 8 rows = 8 distinct positions: no bank conflicts), waits until all
        // transform threads have read (the fp16 tiles overwrite other threads' landing rows), then writes 2 + 2 chunks per half
        // row (a quarter-warp writes one contiguous 128-byte core matrix).  (Eight warps measured the same 1000 cycles per
        // k-block as four -- shared-memory bandwidth, not issue slots, bounds the conversion -- so four of them went to the epilogue.)
        constexpr int kXfThreads = 32 * G_XF_WARPS, kHP = 256 / kXfThreads;       // half rows per thread
        const int tid = (int)threadIdx.x - G_XF_FIRST * 32;
        const int trow = tid & 127, th0 = tid >> 7;
        const uint32_t ready_lead = mapa_u32(smem_u32(&ready[0]), 0);
        const uint32_t swz = (uint32_t)(trow & 7);
        const uint32_t row_off = (uint32_t)((trow >> 3) * 512 + (trow & 7) * 16);   // of this row inside an fp16 tile
        constexpr int kBPer = P::kBRows / 128;                      // B rows per thread
        int stage = 0;
        uint32_t phase = 0;
        const bool gather = a.bg_table != nullptr;
        TILE_LOOP_BEGIN
            (void)mb;
            for (int kidx = 0; kidx < kblocks; ++kidx) {
                { TR_T0(); mbar_wait(&full[stage], phase); TR_ADD(tr0); }     // this CTA's fp32 tiles have landed
                const long long tr_w0 = a.trace ? clock64() : 0;
                const uint32_t sb = smem_u32(smem + stage * P::kStage);
                float4 va[kHP][4], vb[kBPer][kHP][4];
#pragma unroll
                for (int q = 0; q < kHP; ++q) {
                    const int th = th0 + q * (kXfThreads / 128);
#pragma unroll
                    for (int c = 0; c < 4; ++c) va[q][c] = lds128(sb + (uint32_t)trow * 128u + (((uint32_t)(4 * th + c) ^ swz) << 4));
                    if (!gather) {
#pragma unroll
                        for (int h = 0; h < kBPer; ++h)
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                vb[h][q][c] = lds128(sb + kABytes + (uint32_t)(trow + 128 * h) * 128u + (((uint32_t)(4 * th + c) ^ swz) << 4));
                    } else {
                        // gathered B: the landing tile is [k][f] (one bulk copy per node): read this thread's nodes of feature row f
                        // (32 lanes = 32 consecutive f: conflict-free)
#pragma unroll
                        for (int h = 0; h < kBPer; ++h)
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const uint32_t base = sb + kABytes + (uint32_t)((16 * th + 4 * c) * (P::kBRows * 4) + (trow + 128 * h) * 4);
                                vb[h][q][c].x = lds32(base);
                                vb[h][q][c].y = lds32(base + P::kBRows * 4);
                                vb[h][q][c].z = lds32(base + 2 * P::kBRows * 4);
                                vb[h][q][c].w = lds32(base + 3 * P::kBRows * 4);
                            }
                    }
                }
                asm volatile("bar.sync 2, %0;" ::"n"(kXfThreads) : "memory");       // every landing row is in registers
                if (a.bln_mr != nullptr) {
                    // B holds z^T of the previous layer: h = relu((z - mean_k) rstd_k gamma_f + beta_f) for feature row f, node k
                    const int kb = kidx < kA ? kidx : k2s + (kidx - kA);
                    const float2 *mr = reinterpret_cast<const float2 *>(a.bln_mr) + (size_t)(a.batched_b ? g : 0) * a.rows_per_graph;
#pragma unroll
                    for (int h = 0; h < kBPer; ++h) {
                        const int f = nb * G_BN + rank * P::kBRows + trow + 128 * h;
                        const float gam = __ldg(a.bln_gamma + f), bet = __ldg(a.bln_beta + f);
#pragma unroll
                        for (int q = 0; q < kHP; ++q) {
                            const int k0 = kb * G_BK + 16 * (th0 + q * (kXfThreads / 128));
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                float *e = reinterpret_cast<float *>(&vb[h][q][c]);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int k = k0 + 4 * c + j;
                                    const float2 s2 = k < a.rows_per_graph ? __ldg(mr + k) : make_float2(0.0f, 0.0f);
                                    e[j] = fmaxf(fmaf(e[j] - s2.x, s2.y * gam, bet), 0.0f);
                                }
                            }
                        }
                    }
                }
                if (!(a.debug & 1)) {
#pragma unroll
                    for (int q = 0; q < kHP; ++q) {
                        const uint32_t ro = row_off + (uint32_t)((th0 + q * (kXfThreads / 128)) * 256);
                        convert_half_row(va[q], scale_a, sb + ro, sb + kABytes / 2 + ro);
#pragma unroll
                        for (int h = 0; h < kBPer; ++h)
                            convert_half_row(vb[h][q], scale_b, sb + kABytes + ro + (uint32_t)(h * 16 * 512),
                                             sb + kABytes + P::kBTile + ro + (uint32_t)(h * 16 * 512));
                    }
                }
                fence_proxy_async();                                // generic-proxy stores -> visible to the UMMA (async proxy)
                __syncwarp();
                if (lane == 0) {
                    if (rank == 0) mbar_arrive(&ready[stage]);
                    else mbar_arrive_cluster(ready_lead + (uint32_t)(stage * 8));
                }
                if (a.trace) tr1 += clock64() - tr_w0;
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        TILE_LOOP_END
        if (a.trace && tid == 0) { a.trace[blockIdx.x * 8 + 1] = tr0; a.trace[blockIdx.x * 8 + 2] = tr1; }
    }
    if (a.trace && threadIdx.x == 0) a.trace[blockIdx.x * 8 + 7] = clock64() - tr_start;
    if (a.trace && threadIdx.x == 32) {        // (second table after the first 8 * 512 entries)
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        a.trace[4096 + blockIdx.x * 4 + 0] = tr_start - tr_entry;
        a.trace[4096 + blockIdx.x * 4 + 1] = (long long)tr_gt0;
        a.trace[4096 + blockIdx.x * 4 + 2] = (long long)gt1;
    }
#undef TILE_LOOP_BEGIN
#undef TILE_LOOP_END
    tc_fence_before();
    __syncthreads();
    if (CTAS == 2) cluster_sync();       // the peer may still be reading this CTA's shared memory / signalling its barriers
    if (warp == 1) { if (CTAS == 2) tmem_dealloc_pair(tmem_base, 2 * G_BN); else tmem_dealloc(tmem_base, 2 * G_BN); }
}

// ---------------------------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------------------------
// Adj = (E + E^T)/2 + I  (gnn.py:27-30), zero outside the n_g x n_g corner.  32x32 tiles, the transposed tile
// goes through shared memory so that both reads are coalesced.
__global__ void __launch_bounds__(256)
adj_prep_kernel(const float *__restrict__ E, int64_t e_batch, int e_ld, const int32_t *__restrict__ sizes, int n_fixed,
                int ldk, float *__restrict__ adj, unsigned *amax)
{
    __shared__ float tile[32][33];
    float mx = 0.0f;
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
            adj[((size_t)g * n_fixed + i) * ldk + j] = v;
            mx = fmaxf(mx, fabsf(v));
        }
    }
    record_amax(amax, mx);
}

// Same result as adj_prep_kernel with the work organised like class_adj_raw_kernel below: (E + E^T) / 2 is symmetric,
// so a WARP owns the tile pair (I, J) / (J, I), I <= J, reads both source tiles once (16 independent row loads per lane
// in flight) and writes both results; the remaining items of a graph's work list are 32-row strips of zero padding.
// ctas_per_graph CTAs of 4 warps walk each graph's list.  (One CTA per 32x32 tile was latency-bound: 0.045 ms for 118 MB.)
__device__ __forceinline__ void adj_zero_strip(float *__restrict__ adj, size_t base, int ldk, int rows, int i0, int c0, int c1,
                                               int lane)
{
    const int nr = min(32, rows - i0);
    for (int j = c0 + lane; j < c1; j += 32) {
        float *ph = adj + base + (size_t)i0 * ldk + j;
        for (int r = 0; r < nr; ++r) { *ph = 0.0f; ph += ldk; }
    }
}

__global__ void __launch_bounds__(128)
adj_sym_kernel(const float *__restrict__ E, int64_t e_batch, int e_ld, const int32_t *__restrict__ sizes, int n_fixed, int ldk,
               float *__restrict__ adj, int ctas_per_graph, unsigned *amax)
{
    __shared__ float tiles[4][2][32][33];
    float mx = 0.0f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x / ctas_per_graph;
    const int wid = (blockIdx.x % ctas_per_graph) * 4 + warp, nw = ctas_per_graph * 4;
    const int n_g = sizes ? sizes[g] : n_fixed;
    const int ld = e_ld > 0 ? e_ld : n_g;
    const float *Eg = E + (size_t)g * e_batch;
    const size_t base = (size_t)g * n_fixed * ldk;
    const int nt = (n_g + 31) / 32, pairs = nt * (nt + 1) / 2, TR = (n_fixed + 31) / 32;
    float (*T)[33] = tiles[warp][0];
    float (*S)[33] = tiles[warp][1];
    for (int item = wid; item < pairs + TR; item += nw) {
        if (item >= pairs) {                 // zero padding right of (rows < n_g) / instead of (rows >= n_g) the corner
            const int i0 = (item - pairs) * 32;
            adj_zero_strip(adj, base, ldk, n_fixed, i0, i0 < n_g ? nt * 32 : 0, ldk, lane);
            continue;
        }
        int J = (int)((sqrtf(8.0f * (float)item + 1.0f) - 1.0f) * 0.5f);
        while ((J + 1) * (J + 2) / 2 <= item) ++J;
        while (J * (J + 1) / 2 > item) --J;
        const int I = item - J * (J + 1) / 2;
        const int i0 = I * 32, j0 = J * 32;
        const bool vi = i0 + lane < n_g, vj = j0 + lane < n_g;
        const int nvi = min(32, n_g - i0), nvj = min(32, n_g - j0);
        // T[r][lane] = E[j0 + r][i0 + lane]
#pragma unroll 1
        for (int rb = 0; rb < 32; rb += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = (vi && rb + u < nvj) ? __ldg(Eg + (size_t)(j0 + rb + u) * ld + i0 + lane) : 0.0f;
#pragma unroll
            for (int u = 0; u < 16; ++u) T[rb + u][lane] = v[u];
        }
        __syncwarp();
        const bool col_ok = j0 + lane < ldk;
        float *ph = adj + base + (size_t)i0 * ldk + j0 + lane;
        const int diag = (I == J) ? lane : -1;
#pragma unroll 1
        for (int rb = 0; rb < 32; rb += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = (vj && rb + u < nvi) ? __ldg(Eg + (size_t)(i0 + rb + u) * ld + j0 + lane) : 0.0f;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int r = rb + u;
                // outside the n_g x n_g corner both terms are 0 and there is no identity (padded vertices: gnn.py:27-30 on
                // zero-padded edges would add it, but those rows are masked and their columns multiply zero features)
                float sym = 0.0f;
                if (r < nvi && vj) sym = (v[u] + T[lane][r]) / 2.0f + ((r == diag) ? 1.0f : 0.0f);
                S[r][lane] = sym;
                mx = fmaxf(mx, fabsf(sym));
                if (col_ok && i0 + r < n_fixed) *ph = sym;
                ph += ldk;
            }
        }
        __syncwarp();
        if (I != J && i0 + lane < ldk) {
            ph = adj + base + (size_t)j0 * ldk + i0 + lane;
            const int rows = min(32, n_fixed - j0);
#pragma unroll 4
            for (int r = 0; r < rows; ++r) {
                *ph = S[lane][r];
                ph += ldk;
            }
        }
        __syncwarp();
    }
    record_amax(amax, mx);
}

// X0^T[g, d, i] = Emb[ids[g, i], d] (gnn.py:91), zero for i >= n_g.  One CTA per (32 nodes, 256 features, graph):
// each warp reads whole 1 KB table rows (8 coalesced loads per node, the id fetched once), the slab is transposed through
// shared memory and written as 128-byte row segments of X^T.
__global__ void __launch_bounds__(256)
embed_gather_t_kernel(const float *__restrict__ emb, const int64_t *__restrict__ ids, int ld_ids,
                      const int32_t *__restrict__ sizes, int n_fixed, int ldk, int D, float *__restrict__ xt, unsigned *amax)
{
    __shared__ float tile[32][257];
    const int g = blockIdx.z;
    const int i0 = blockIdx.x * 32, d0 = blockIdx.y * 256;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_g = sizes ? sizes[g] : n_fixed;
    // the codes of this warp's four nodes are requested together with the graph's size (every slot < n_fixed is readable), so
    // the kernel is three dependent round trips long -- size + codes, table rows, stores -- instead of four
    int64_t code[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {                     // node i0 + warp + 8 q
        const int i = i0 + warp + 8 * q;
        code[q] = (i < n_fixed) ? ids[(size_t)g * ld_ids + i] : 0;
    }
    if (i0 >= max(32, (n_g + 31) / 32 * 32)) return;   // past the k-blocks the adjacency GEMM reads for this graph
    const int dn = min(256, D - d0);                    // features of this slab (multiple of 32)
    // all 32 loads of a warp (4 nodes x 8 row segments) are issued before the first one is used: written as a plain
    // load-then-store loop the compiler kept one load in flight per warp and the kernel ran at L2 latency (28 us at cfg2)
    const float *rows[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) rows[q] = (i0 + warp + 8 * q < n_g) ? emb + (size_t)code[q] * D + d0 : nullptr;
    float v[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 8; ++c) v[q][c] = (rows[q] != nullptr && c * 32 < dn) ? __ldg(rows[q] + c * 32 + lane) : 0.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c * 32 < dn) tile[warp + 8 * q][c * 32 + lane] = v[q][c];
    __syncthreads();
    const int i = i0 + lane;
    float mx = 0.0f;
    if (i < ldk)
        for (int d = warp; d < dn; d += 8) {           // feature d0 + d, nodes i0 .. i0 + 31 (coalesced along i)
            const float t = tile[lane][d];
            xt[((size_t)g * D + d0 + d) * ldk + i] = t;
            mx = fmaxf(mx, fabsf(t));
        }
    record_amax(amax, mx);
}

constexpr int kTableSlices = 8;   // CTAs per class summing the pruned vertices' table rows (pool_table_rows_kernel)

// pooled partials: partial[g, chunk, d] = sum over the chunk's nodes of H[g, i, d] * w[g, i]   (gnn.py:94-95)
__global__ void __launch_bounds__(256)
pool_rows_kernel(const float *__restrict__ H, const float *__restrict__ vertex_w, int ld_v, const int32_t *__restrict__ sizes,
                 int n_fixed, int D, int chunks, float *__restrict__ partial, const float *__restrict__ extra)
{
    const int g = blockIdx.y, chunk = blockIdx.x;
    const int n_g = sizes ? sizes[g] : n_fixed;
    const int per = ceil_div(n_fixed, chunks);
    const int r0 = chunk * per, r1 = min(n_g, r0 + per);
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.0f;
        for (int r = r0; r < r1; ++r) acc = fmaf(H[((size_t)g * n_fixed + r) * D + d], vertex_w[(size_t)g * ld_v + r], acc);
        if (extra && chunk == 0) {                                 // vertices handled outside the GEMMs (class side)
            float e = 0.0f;
            for (int y = 0; y < kTableSlices; ++y) e += extra[((size_t)g * kTableSlices + y) * D + d];
            acc += e;
        }
        partial[((size_t)g * chunks + chunk) * D + d] = acc;
    }
}

// partial[g, 0, d] = sum of the 32-row group sums the last GEMM's epilogue emitted for graph g (+ the table share of the
// class side's pruned vertices), partial[g, c > 0, d] = 0: the layout pool_fc_kernel consumes.  Groups are summed in
// ascending row order, so the result does not depend on the launch configuration.
__global__ void __launch_bounds__(256)
pool_groups_reduce_kernel(const float *__restrict__ groups, const int32_t *__restrict__ sizes, int n_fixed, int D, int chunks,
                          float *__restrict__ partial, const float *__restrict__ extra)
{
    const int g = blockIdx.x, d = threadIdx.x;
    if (d >= D) return;
    const int n_g = sizes ? sizes[g] : n_fixed;
    float acc = 0.0f;
    if (n_g > 0) {
        const int r0 = g * n_fixed, q0 = r0 / 32, q1 = (r0 + n_g - 1) / 32;
        for (int q = q0; q <= q1; ++q) {
            const int slot = ((q * 32) / n_fixed == g) ? 0 : 1;      // is g the first or the second graph of the group
            acc += groups[((size_t)q * 2 + slot) * D + d];
        }
    }
    if (extra) {
        float e = 0.0f;
        for (int y = 0; y < kTableSlices; ++y) e += extra[((size_t)g * kTableSlices + y) * D + d];
        acc += e;
    }
    partial[((size_t)g * chunks) * D + d] = acc;
    for (int c = 1; c < chunks; ++c) partial[((size_t)g * chunks + c) * D + d] = 0.0f;
}

// The same group sum followed by mean + fc (gnn.py:96-97) in one kernel: out[g, :] = fc(pooled[g] / N).  grid (G, 8): every
// CTA rebuilds the pooled vector of its graph in shared memory (<= n/32 + 1 group rows) and computes 1/8 of the outputs.
__global__ void __launch_bounds__(256)
pool_groups_fc_kernel(const float *__restrict__ groups, const int32_t *__restrict__ sizes, int n_fixed, int D,
                      const float *__restrict__ extra, const int32_t *__restrict__ mean_div, const float *__restrict__ fc_w,
                      const float *__restrict__ fc_b, float *__restrict__ out)
{
    extern __shared__ float pooled[];
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_g = sizes ? sizes[g] : n_fixed;
    const float div = (float)(mean_div ? *mean_div : n_fixed);
    // (every loop below issues its batch of independent loads before the first use: these kernels are a few dependent L2
    // round trips long, and a load-use-load-use chain made them 19 us at cfg2)
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.0f;
        if (n_g > 0) {
            const int r0 = g * n_fixed, q0 = r0 / 32, q1 = (r0 + n_g - 1) / 32;
            for (int qb = q0; qb <= q1; qb += 8) {
                float t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int q = qb + u;
                    // slot 0: g is the first graph of the 32-row group, i.e. the group starts inside g (no division: q * 32 <= r0 +
                    // n_g - 1 < r0 + n_fixed already holds for q <= q1)
                    t[u] = q <= q1 ? groups[((size_t)q * 2 + ((q * 32 >= r0) ? 0 : 1)) * D + d] : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (qb + u <= q1) acc += t[u];
            }
        }
        if (extra) {
            float t[kTableSlices];
#pragma unroll
            for (int y = 0; y < kTableSlices; ++y) t[y] = extra[((size_t)g * kTableSlices + y) * D + d];
            float e = 0.0f;
#pragma unroll
            for (int y = 0; y < kTableSlices; ++y) e += t[y];
            acc += e;
        }
        pooled[d] = acc / div;
    }
    __syncthreads();
    const int per = (D + gridDim.y - 1) / gridDim.y;
    const int o_end = min(D, (int)(blockIdx.y + 1) * per);
    const int nwarps = (int)(blockDim.x >> 5);
    // a warp computes four outputs at a time; per output the products are folded in ascending d per lane, then across lanes
    for (int o0 = blockIdx.y * per + warp; o0 < o_end; o0 += 4 * nwarps) {
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int d0 = lane; d0 < D; d0 += 8 * kWarp) {
            float wv[4][8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int o = o0 + j * nwarps;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int d = d0 + u * kWarp;
                    wv[j][u] = (o < o_end && d < D) ? __ldg(fc_w + (size_t)o * D + d) : 0.0f;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int d = d0 + u * kWarp;
                    if (d < D) acc[j] = fmaf(pooled[d], wv[j][u], acc[j]);
                }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = o0 + j * nwarps;
            const float r = warp_sum(acc[j]);
            if (lane == 0 && o < o_end) out[(size_t)g * D + o] = r + fc_b[o];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// class side: pruned vertices without GEMM rows
// ---------------------------------------------------------------------------------------------------------------
// A pruned class vertex has an all-zero row and column in class_edges, so in every layer ((E + E^T) / 2 + I) X keeps just
// its own feature: its activations depend on its CODE only,
//     T_0[c] = relu(LN(Emb[c] W_0^T + b_0)),   T_l[c] = relu(LN(T_{l-1}[c] W_l^T + b_l)),
// (M + 1)-row tables that cost two tiny GEMMs, instead of ~55 % of the rows of every class-side GEMM.  The weighted
// pooling of those vertices is a gather from the last table.  (The tables come from rows_linear_kernel, gnn.cu, with
// the bias + LayerNorm + ReLU fused.)
// extra[k, y, d] = sum over the pruned vertices i >= n_act[k], i = n_act[k] + y (mod kTableSlices) (permuted order) of
// w[k, i] * T[ids[k, i], d]; grid (K, kTableSlices), one thread per feature (D <= 256)
__global__ void __launch_bounds__(256)
pool_table_rows_kernel(const float *__restrict__ T, const int64_t *__restrict__ ids, const float *__restrict__ w,
                       const int32_t *__restrict__ n_act, int Vc, int D, float *__restrict__ extra)
{
    // the slice's codes and weights go to shared memory in one round trip (they are the same for every thread), then 16 table
    // entries per thread are in flight at a time; the sum is folded in the order of the plain loop over i
    extern __shared__ unsigned char pool_smem[];
    int *s_code = reinterpret_cast<int *>(pool_smem);
    float *s_w = reinterpret_cast<float *>(s_code + (Vc + kTableSlices - 1) / kTableSlices);
    const int k = blockIdx.x, d = threadIdx.x;
    const int64_t *idk = ids + (size_t)k * Vc;
    const float *wk = w + (size_t)k * Vc;
    const int first = n_act[k] + blockIdx.y;
    const int cnt = first < Vc ? (Vc - first + kTableSlices - 1) / kTableSlices : 0;
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
        s_code[j] = (int)idk[first + j * kTableSlices];
        s_w[j] = wk[first + j * kTableSlices];
    }
    __syncthreads();
    if (d >= D) return;
    float acc = 0.0f;
    int j = 0;
    for (; j + 16 <= cnt; j += 16) {
        float t[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) t[u] = __ldg(T + (size_t)s_code[j + u] * D + d);
#pragma unroll
        for (int u = 0; u < 16; ++u) acc = fmaf(t[u], s_w[j + u], acc);
    }
    for (; j < cnt; ++j) acc = fmaf(__ldg(T + (size_t)s_code[j] * D + d), s_w[j], acc);
    extra[((size_t)k * kTableSlices + blockIdx.y) * D + d] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// class side: compaction of each class graph to its un-pruned vertices
// ---------------------------------------------------------------------------------------------------------------
// A vertex whose normalised weight is <= prune_node_threshold has an all-zero row and column in class_edges
// (schema_net.py:157-166), so in ((E+E^T)/2 + I) X it only keeps its own feature.  Vertices are therefore reordered
// "active first" (stable), the adjacency is built for the active x active corner only, and the GEMM's row blocks of
// inactive vertices visit just their identity diagonal -- same result, a fraction of the flops and bytes.
// One CTA per class: stable partition by prefix sums over Vc flags.
__global__ void __launch_bounds__(1024)
class_perm_kernel(const float *__restrict__ cv, const int64_t *__restrict__ ci, int Vc, float thr, int prune,
                  int32_t *__restrict__ n_act, int32_t *__restrict__ old_of_new, int64_t *__restrict__ pid,
                  float *__restrict__ pvw)
{
    __shared__ int warp_tot[32];
    __shared__ int base_act, total_act;
    const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *cvk = cv + (size_t)k * Vc;
    if (tid == 0) base_act = 0;
    __syncthreads();
    // pass 1: count active; pass 2: place.  Chunks of blockDim vertices keep the partition stable.
    int total = 0;
    for (int i0 = 0; i0 < Vc; i0 += blockDim.x) {
        const int i = i0 + tid;
        total += (i < Vc && (!prune || cvk[i] > thr)) ? 1 : 0;
    }
    total = (int)warp_sum((float)total);   // counts <= 1024 per warp: exact in fp32
    if (lane == 0) warp_tot[warp] = total;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w];
        total_act = t;
        n_act[k] = t;
    }
    __syncthreads();
    const int nA = total_act;
    int done_act = 0, done_in = 0;   // running offsets (uniform across the block)
    for (int i0 = 0; i0 < Vc; i0 += blockDim.x) {
        const int i = i0 + tid;
        const bool in = i < Vc;
        const bool act = in && (!prune || cvk[i] > thr);
        const unsigned bal = __ballot_sync(kFull, act);
        const unsigned bal_in = __ballot_sync(kFull, in);
        const int before = __popc(bal & ((1u << lane) - 1));
        const int before_in = __popc(bal_in & ((1u << lane) - 1));
        __syncthreads();
        if (lane == 0) warp_tot[warp] = __popc(bal) | (__popc(bal_in) << 16);
        __syncthreads();
        int wa = 0, wi = 0, ta = 0, ti = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            const int a_ = warp_tot[w] & 0xffff, n_ = warp_tot[w] >> 16;
            if (w < warp) { wa += a_; wi += n_; }
            ta += a_; ti += n_;
        }
        if (in) {
            const int rank_act = done_act + wa + before;
            const int rank_inact = done_in + (wi - wa) + (before_in - before);
            const int pos = act ? rank_act : nA + rank_inact;
            old_of_new[(size_t)k * Vc + pos] = i;
            pid[(size_t)k * Vc + pos] = ci[(size_t)k * Vc + i];
            pvw[(size_t)k * Vc + pos] = cvk[i];
        }
        done_act += ta;
        done_in += ti - ta;
    }
}

// Compacted class adjacency.  Only the parts the GEMM reads are written: the active corner (plus its padding
// up to the tile edges) and the 128x128 diagonal blocks of row blocks that contain inactive vertices.
__global__ void __launch_bounds__(256)
class_adj_prep_kernel(const float *__restrict__ ce, int K, int Vc, int ldk, int unit_rows, int with_tail,
                      const int32_t *__restrict__ n_act, const int32_t *__restrict__ old_of_new, float *__restrict__ adj,
                      unsigned *amax)
{
    // One CTA per 32x32 tile; tiles outside the regions the GEMM reads exit at once.  (A persistent variant that walked
    // the tile space with two block barriers per tile measured 45 % slower.)
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    {
        const int k = blockIdx.z;
        const int pi0 = blockIdx.y * 32, pj0 = blockIdx.x * 32;
        const int nA = n_act[k];
        // unit_rows = rows one GEMM work unit covers (128 for a single CTA, 256 for a CTA pair): see TILE_LOOP in gemm3x_kernel
        const int rowsA = min(Vc, (nA + unit_rows - 1) / unit_rows * unit_rows), colsA = (nA + G_BK - 1) / G_BK * G_BK;
        const int ub = pi0 / unit_rows;
        const bool in_a = pi0 < rowsA && pj0 < colsA;
        const bool in_b = with_tail && (ub + 1) * unit_rows > nA && pj0 / unit_rows == ub;
        if (!in_a && !in_b) return;
        float mx = 0.0f;
        const float *cek = ce + (size_t)k * Vc * Vc;
        const int32_t *old = old_of_new + (size_t)k * Vc;
        const bool any_active = pi0 < nA && pj0 < nA;
        float direct[4];
        if (any_active) {
            // all eight gathered loads of a thread are issued before any is used
            const int oi = (pi0 + tx < nA) ? old[pi0 + tx] : -1;
            float tr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {       // transposed tile: ce[old[pj0 + r]][old[pi0 + tx]]
                const int r = ty + 8 * u;
                tr[u] = (pj0 + r < nA && oi >= 0) ? __ldg(cek + (size_t)old[pj0 + r] * Vc + oi) : 0.0f;
            }
            const int oj = (pj0 + tx < nA) ? old[pj0 + tx] : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = ty + 8 * u;
                direct[u] = (pi0 + r < nA && oj >= 0) ? __ldg(cek + (size_t)old[pi0 + r] * Vc + oj) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) tile[ty + 8 * u][tx] = tr[u];
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = ty + 8 * u;
            const int pi = pi0 + r, pj = pj0 + tx;
            if (pi < Vc && pj < ldk) {
                float v = (pi == pj) ? 1.0f : 0.0f;
                if (any_active && pi < nA && pj < nA) v = (direct[u] + tile[tx][r]) / 2.0f + v;
                adj[((size_t)k * Vc + pi) * ldk + pj] = v;
                mx = fmaxf(mx, fabsf(v));
            }
        }
        record_amax(amax, mx);
    }
}

// Same output as class_adj_prep_kernel without a materialised class_edges tensor: entry (i, j) of class_edges is rebuilt
// from the (pruned) edge parameter and the per-row normalisers the atlas pass left in rowinv,
//   ce[i][j] = nan_to_num(max(ew[i][j], 0) * rowinv[i]),  0 on the diagonal when self loops are removed
// (bit-identical to what class_edges_*_kernel would have stored).  (E + E^T) / 2 is symmetric: a WARP owns the pair of
// 32x32 tiles (I, J) and (J, I), I <= J, gathers both source tiles once (8 independent loads per lane in flight, row
// indices and normalisers passed around by shuffle) and writes both results.  A fixed number of CTAs per class walks
// that class's work list (tile pairs of the active corner, then the identity / zero padding tiles the GEMM reads), so no
// CTA is launched just to exit.
constexpr int kAdjWarps = 4;
constexpr int kAdjBatch = 16;   // gathered loads a lane keeps in flight (the kernel is bound by DRAM latency)

// One 32-row strip of padding: columns [c0, c1) of rows [pi0, pi0 + 32) get the identity / zero pattern.
__device__ __forceinline__ void adj_fill_strip(float *__restrict__ adj, size_t base, int ldk, int Vc, int pi0, int c0, int c1,
                                               int lane)
{
    c1 = min(c1, ldk);
    const int rows = min(32, Vc - pi0);
    for (int pj = c0 + lane; pj < c1; pj += 32) {
        float *ph = adj + base + (size_t)pi0 * ldk + pj;
        const int diag = pj - pi0;          // row of this strip that holds the 1 of column pj
        for (int r = 0; r < rows; ++r) {
            *ph = (r == diag) ? 1.0f : 0.0f;
            ph += ldk;
        }
    }
}

// The work list of class k (tile pairs of the active corner, then the padding strips) walked by worker `wid` of `nw` warps;
// T, S: two 32 x 33 shared-memory tiles owned by this warp.  Returns the largest magnitude it wrote.
__device__ __forceinline__ float class_adj_items(const float *__restrict__ ew, const float *__restrict__ rowinv, int k, int Vc, int ldk,
                                                 int unit_rows, int with_tail, int remove_self_loop,
                                                 const int32_t *__restrict__ n_act, const int32_t *__restrict__ old_of_new,
                                                 float *__restrict__ adj, int wid, int nw, float (*T)[33], float (*S)[33])
{
    const int lane = threadIdx.x & 31;
    float mx = 1.0f;                      // (the identity / padding strips hold ones)
    const int nA = n_act[k];
    const int rowsA = min(Vc, (nA + unit_rows - 1) / unit_rows * unit_rows), colsA = (nA + G_BK - 1) / G_BK * G_BK;
    const int nt = (nA + 31) / 32, pairs = nt * (nt + 1) / 2;
    const int TR = (Vc + 31) / 32;
    const size_t base = (size_t)k * Vc * ldk;
    const float *ewk = ew + (size_t)k * Vc * Vc;
    const float *rik = rowinv + (size_t)k * Vc;
    const int32_t *old = old_of_new + (size_t)k * Vc;
    // work list of the class: tile pairs of the active corner first (the expensive items), then one item per 32-row strip
    // of padding
    for (int item = wid; item < pairs + TR; item += nw) {
        if (item >= pairs) {
            const int pi0 = (item - pairs) * 32;
            const int act_end = pi0 < nA ? nt * 32 : 0;                 // columns [0, act_end) belong to tile pairs
            const int a_end = pi0 < rowsA ? colsA : 0;                  // region A: the active corner up to the tile edges
            if (a_end > act_end) adj_fill_strip(adj, base, ldk, Vc, pi0, act_end, a_end, lane);
            const int ub = pi0 / unit_rows;                             // region B: diagonal blocks with inactive vertices
            if (with_tail && (ub + 1) * unit_rows > nA) {
                const int b0 = max(ub * unit_rows, max(act_end, a_end)), b1 = (ub + 1) * unit_rows;
                if (b1 > b0) adj_fill_strip(adj, base, ldk, Vc, pi0, b0, b1, lane);
            }
            continue;
        }
        // item -> (I, J), I <= J: item = J (J + 1) / 2 + I
        int J = (int)((sqrtf(8.0f * (float)item + 1.0f) - 1.0f) * 0.5f);
        while ((J + 1) * (J + 2) / 2 <= item) ++J;
        while (J * (J + 1) / 2 > item) --J;
        const int I = item - J * (J + 1) / 2;
        const int pi0 = I * 32, pj0 = J * 32;
        // lanes past the active corner read element 0 of the class with a zero normaliser: their values become exact zeros
        const bool vi = pi0 + lane < nA, vj = pj0 + lane < nA;
        const int oi = vi ? old[pi0 + lane] : 0, oj = vj ? old[pj0 + lane] : 0;
        const float si = vi ? __ldg(rik + oi) : 0.0f, sj = vj ? __ldg(rik + oj) : 0.0f;
        const float *col_i = ewk + oi, *col_j = ewk + oj;
        const int nvi = min(32, nA - pi0), nvj = min(32, nA - pj0);   // valid rows of the two source tiles
        // rows whose normaliser is not an ordinary positive number (empty, infinite or NaN row sums) need nan_to_num
        const bool special = __any_sync(kFull, (vi && !(si > 0.0f && si <= FLT_MAX)) || (vj && !(sj > 0.0f && sj <= FLT_MAX)));
        // source tile of the mirror: T[r][lane] = ce[old_j(r)][old_i(lane)]
#pragma unroll 1
        for (int rb = 0; rb < 32; rb += kAdjBatch) {
            float v[kAdjBatch];
#pragma unroll
            for (int u = 0; u < kAdjBatch; ++u) v[u] = __ldg(col_i + (size_t)__shfl_sync(kFull, oj, rb + u) * Vc);
#pragma unroll
            for (int u = 0; u < kAdjBatch; ++u) {
                float e = fmaxf(v[u], 0.0f) * __shfl_sync(kFull, sj, rb + u);
                if (!vi || rb + u >= nvj) e = 0.0f;
                if (special) e = nan_to_num0(e);
                if (remove_self_loop && __shfl_sync(kFull, oj, rb + u) == oi) e = 0.0f;
                T[rb + u][lane] = e;
            }
        }
        __syncwarp();
        // this tile: S[r][lane] = (ce[old_i(r)][old_j(lane)] + ce[old_j(lane)][old_i(r)]) / 2 + identity
        const bool col_ok = pj0 + lane < ldk;
        float *ph = adj + base + (size_t)pi0 * ldk + pj0 + lane;
        const int diag = (I == J) ? lane : -1;
#pragma unroll 1
        for (int rb = 0; rb < 32; rb += kAdjBatch) {
            float v[kAdjBatch];
#pragma unroll
            for (int u = 0; u < kAdjBatch; ++u) v[u] = __ldg(col_j + (size_t)__shfl_sync(kFull, oi, rb + u) * Vc);
#pragma unroll
            for (int u = 0; u < kAdjBatch; ++u) {
                const int r = rb + u;
                float e = fmaxf(v[u], 0.0f) * __shfl_sync(kFull, si, r);
                if (!vj || r >= nvi) e = 0.0f;
                if (special) e = nan_to_num0(e);
                if (remove_self_loop && __shfl_sync(kFull, oi, r) == oj) e = 0.0f;
                const float sym = (e + T[lane][r]) / 2.0f + ((r == diag) ? 1.0f : 0.0f);
                S[r][lane] = sym;
                mx = fmaxf(mx, fabsf(sym));
                if (col_ok && pi0 + r < Vc) *ph = sym;
                ph += ldk;
            }
        }
        __syncwarp();
        if (I != J && pi0 + lane < ldk) {
            ph = adj + base + (size_t)pj0 * ldk + pi0 + lane;          // mirror tile (J, I)
            const int rows = min(32, Vc - pj0);
#pragma unroll 4
            for (int r = 0; r < rows; ++r) {
                *ph = S[lane][r];
                ph += ldk;
            }
        }
        __syncwarp();
    }
    return mx;
}

__global__ void __launch_bounds__(kAdjWarps * 32)
class_adj_raw_kernel(const float *__restrict__ ew, const float *__restrict__ rowinv, int K, int Vc, int ldk, int unit_rows,
                     int with_tail, int remove_self_loop, const int32_t *__restrict__ n_act,
                     const int32_t *__restrict__ old_of_new, float *__restrict__ adj, int ctas_per_class, unsigned *amax)
{
    __shared__ float tiles[kAdjWarps][2][32][33];
    const int warp = threadIdx.x >> 5;
    // last classes first: their edge parameters are what the atlas pass, which ran just before, left in L2
    const int k = K - 1 - blockIdx.x / ctas_per_class;
    const int wid = (blockIdx.x % ctas_per_class) * kAdjWarps + warp, nw = ctas_per_class * kAdjWarps;
    const float mx = class_adj_items(ew, rowinv, k, Vc, ldk, unit_rows, with_tail, remove_self_loop, n_act, old_of_new, adj, wid, nw,
                                     tiles[warp][0], tiles[warp][1]);
    record_amax(amax, mx);
}

// amax slots: the adjacency, the Linear weights of layer l, the node features entering layer l (l = num_layers: never
// consumed), the adjacency product Y of layer l
constexpr int kMaxTcLayers = 16;
constexpr int AM_ADJ = 0, AM_W = 1, AM_X = 1 + kMaxTcLayers, AM_Y = 2 + 2 * kMaxTcLayers;
// a-priori bound of relu(LayerNorm_l(.)) = max_f |gamma_f| sqrt(D) + |beta_f| (operands normalised on the fly are never
// materialised, so nothing can record their maximum), and the embedding table's maximum
constexpr int AM_LN = 2 + 3 * kMaxTcLayers, AM_EMB = 2 + 4 * kMaxTcLayers, AM_FC = AM_EMB + 1, AM_POOL = AM_EMB + 2;   // 128 slots in all

// Operand bounds that depend on the parameters only, one launch per forward.  blockIdx.y < layers: the largest magnitude of
// layer y's Linear weights (B operand of the linear GEMMs) and, by block (0, y), the a-priori bound of that layer's
// relu(LayerNorm(.)) output, max_f |gamma_f| sqrt(D) + |beta_f|  (|normalised value| <= sqrt(D - 1));
// blockIdx.y == layers: the largest magnitude of the embedding table (A operand of the table product P_0 = Emb W_0^T).
struct ParamPtrs { const float *w[16], *gamma[16], *beta[16]; const float *emb, *fc; long long emb_n; int layers; };
__global__ void __launch_bounds__(256) param_bounds_kernel(ParamPtrs p, int D, unsigned *amax)
{
    const int y = blockIdx.y;
    float mx = 0.0f;
    if (y < p.layers) {
        const float4 *w = reinterpret_cast<const float4 *>(p.w[y]);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D * D / 4; i += gridDim.x * blockDim.x) {
            const float4 v = __ldg(w + i);
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        }
        record_amax(amax + AM_W + y, mx);
        if (blockIdx.x == 0) {
            float b = 0.0f;
            const float rt = sqrtf((float)D);
            for (int i = threadIdx.x; i < D; i += blockDim.x) b = fmaxf(b, fmaf(fabsf(p.gamma[y][i]), rt, fabsf(p.beta[y][i])));
            record_amax(amax + AM_LN + y, b);
        }
    } else if (y == p.layers) {                 // fc weights (the pooled features are multiplied by them on the tensor cores)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D * D; i += gridDim.x * blockDim.x) mx = fmaxf(mx, fabsf(__ldg(p.fc + i)));
        record_amax(amax + AM_FC, mx);
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.emb_n; i += (long long)gridDim.x * blockDim.x)
            mx = fmaxf(mx, fabsf(__ldg(p.emb + i)));
        record_amax(amax + AM_EMB, mx);
    }
}

// embed_dim > 256: merge the per-tile LayerNorm statistics of every row (Chan et al.: tiles of equal size T = 256) into
// (mean, rstd):  mean = avg_t mean_t,  M2 = sum_t M2_t + T sum_t (mean_t - mean)^2,  rstd = 1 / sqrt(M2 / D + eps)
__global__ void __launch_bounds__(256)
ln_stats_merge_kernel(const float *__restrict__ stats, int64_t rows, int nt, float eps, float *__restrict__ mr)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float2 *s = reinterpret_cast<const float2 *>(stats) + r * nt;
    float mean = 0.0f;
    for (int t = 0; t < nt; ++t) mean += s[t].x;
    mean /= (float)nt;
    float m2 = 0.0f;
    for (int t = 0; t < nt; ++t) { const float dlt = s[t].x - mean; m2 += s[t].y + (float)G_BN * dlt * dlt; }
    reinterpret_cast<float2 *>(mr)[r] = make_float2(mean, 1.0f / sqrtf(m2 / (float)(nt * G_BN) + eps));
}

// embed_dim > 256, last layer: pooled partials straight from the un-normalised z rows,
//   partial[g, chunk, d] = sum over the chunk's nodes of relu((z[g,i,d] - mean_i) rstd_i gamma_d + beta_d) * w[g, i]   (gnn.py:45,94-95)
__global__ void __launch_bounds__(256)
pool_ln_rows_kernel(const float *__restrict__ Z, const float *__restrict__ mr, const float *__restrict__ gamma,
                    const float *__restrict__ beta, const float *__restrict__ vertex_w, int ld_v, const int32_t *__restrict__ sizes,
                    int n_fixed, int D, int chunks, float *__restrict__ partial)
{
    // a thread owns 4 consecutive features (one float4 per row: a warp reads 512 contiguous bytes), 4 rows in flight
    const int g = blockIdx.y, chunk = blockIdx.x;
    const int n_g = sizes ? sizes[g] : n_fixed;
    const int per = ceil_div(n_fixed, chunks);
    const int r0 = chunk * per, r1 = min(n_g, r0 + per);
    const float2 *mrg = reinterpret_cast<const float2 *>(mr) + (size_t)g * n_fixed;
    const float *wg = vertex_w + (size_t)g * ld_v;
    for (int d = threadIdx.x * 4; d < D; d += blockDim.x * 4) {
        const float4 gam = *reinterpret_cast<const float4 *>(gamma + d), bet = *reinterpret_cast<const float4 *>(beta + d);
        const float *zp = Z + ((size_t)g * n_fixed + r0) * D + d;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int r = r0;
        for (; r + 4 <= r1; r += 4) {
            float4 z[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) z[u] = __ldcs(reinterpret_cast<const float4 *>(zp + (size_t)u * D));
#pragma unroll
            for (int u = 0; u < 4; ++u) {           // rows in ascending order: the sum does not depend on the unrolling
                const float2 s = mrg[r + u];
                const float w = wg[r + u];
                acc.x = fmaf(fmaxf(fmaf(z[u].x - s.x, s.y * gam.x, bet.x), 0.0f), w, acc.x);
                acc.y = fmaf(fmaxf(fmaf(z[u].y - s.x, s.y * gam.y, bet.y), 0.0f), w, acc.y);
                acc.z = fmaf(fmaxf(fmaf(z[u].z - s.x, s.y * gam.z, bet.z), 0.0f), w, acc.z);
                acc.w = fmaf(fmaxf(fmaf(z[u].w - s.x, s.y * gam.w, bet.w), 0.0f), w, acc.w);
            }
            zp += (size_t)4 * D;
        }
        for (; r < r1; ++r) {
            const float4 z = __ldcs(reinterpret_cast<const float4 *>(zp));
            const float2 s = mrg[r];
            const float w = wg[r];
            acc.x = fmaf(fmaxf(fmaf(z.x - s.x, s.y * gam.x, bet.x), 0.0f), w, acc.x);
            acc.y = fmaf(fmaxf(fmaf(z.y - s.x, s.y * gam.y, bet.y), 0.0f), w, acc.y);
            acc.z = fmaf(fmaxf(fmaf(z.z - s.x, s.y * gam.z, bet.z), 0.0f), w, acc.z);
            acc.w = fmaf(fmaxf(fmaf(z.w - s.x, s.y * gam.w, bet.w), 0.0f), w, acc.w);
            zp += D;
        }
        *reinterpret_cast<float4 *>(partial + ((size_t)g * chunks + chunk) * D + d) = acc;
    }
}

// max |x| of a small tensor into an amax slot
__global__ void __launch_bounds__(256) absmax_kernel(const float *__restrict__ x, int64_t n, unsigned *slot)
{
    float mx = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) mx = fmaxf(mx, fabsf(__ldg(x + i)));
    record_amax(slot, mx);
}

// ---------------------------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------------------------
// embed_dim > 256: LayerNorm + ReLU as its own pass over Z = Y W^T + b (the GEMM epilogue cannot hold a row wider than
// one 256-column accumulator).  One CTA per (32-node slab, graph); a warp normalises 4 rows (two-pass mean / variance).
// kTranspose: stage the slab in shared memory and write H^T for the next layer's adjacency GEMM (coalesced along
// the node index, zeros for nodes >= n_g); otherwise overwrite Z with H in place (the pooling reads rows).
template <bool kTranspose>
__global__ void __launch_bounds__(256)
ln_relu_wide_kernel(float *__restrict__ Z, const int32_t *__restrict__ row_sizes, int n_fixed, int ldk, int D,
                    const float *__restrict__ gamma, const float *__restrict__ beta, float eps, float *__restrict__ xt,
                    unsigned *amax)
{
    extern __shared__ float slab[];   // kTranspose: [32][D + 1]
    const int g = blockIdx.y, i0 = blockIdx.x * 32;
    const int n_g = row_sizes ? row_sizes[g] : n_fixed;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int rr = warp; rr < 32; rr += 8) {
        const int i = i0 + rr;
        const bool live = i < n_g;
        float *z = Z + ((size_t)g * n_fixed + i) * D;
        float mean = 0.0f, rstd = 0.0f;
        if (live) {
            float sum = 0.0f;
            for (int dd = lane; dd < D; dd += kWarp) sum += z[dd];
            mean = warp_sum(sum) / (float)D;
            float var = 0.0f;
            for (int dd = lane; dd < D; dd += kWarp) { const float t = z[dd] - mean; var = fmaf(t, t, var); }
            rstd = 1.0f / sqrtf(warp_sum(var) / (float)D + eps);
        }
        for (int dd = lane; dd < D; dd += kWarp) {
            const float y = live ? fmaxf((z[dd] - mean) * rstd * gamma[dd] + beta[dd], 0.0f) : 0.0f;
            if (kTranspose) slab[rr * (D + 1) + dd] = y;
            else if (live) z[dd] = y;
        }
    }
    if (kTranspose) {
        __syncthreads();
        const int i = i0 + lane;
        float mx = 0.0f;
        if (i < ldk)
            for (int dd = warp; dd < D; dd += 8) {
                const float v = i < n_fixed ? slab[lane * (D + 1) + dd] : 0.0f;
                xt[((size_t)g * D + dd) * ldk + i] = v;
                mx = fmaxf(mx, v);
            }
        record_amax(amax, mx);
    }
}

static size_t al256(size_t x) { return (x + 255) / 256 * 256; }

bool gnn_tc_supported(int D, int n_fixed)
{
    return D % G_BN == 0 && D <= kMaxDim && n_fixed >= 32 && encode_tiled_fn() != nullptr;
}

struct TcBuffers {
    float *adj, *xt, *xt2, *y, *tab, *h_rows;     // tab: (M+1)-row table scratch (P_0 = Emb W_0^T), same size as y
    unsigned *amax;                               // operand maxima (bit patterns), see AM_* below; zeroed per forward
    float *stats, *mr;                            // embed_dim > 256: per-tile LayerNorm statistics and the merged (mean, rstd)
    int32_t *n_act, *old_of_new;
    float *rowinv, *pool_extra, *pool_groups;
    int64_t *pid;
    float *pvw;
    int ldk;
    size_t bytes;
};

static TcBuffers carve_tc(void *base, int G, int n_fixed, int D)
{
    TcBuffers b{};
    b.ldk = (n_fixed + 31) / 32 * 32;   // rows of the K-major operands start on 128-byte lines: a TMA box row is one line, not two halves
    char *p = (char *)base;
    size_t off = 0;
    const size_t adj_b = al256((size_t)G * n_fixed * b.ldk * 4), xt_b = al256((size_t)G * D * b.ldk * 4);
    const size_t y_b = al256((size_t)G * n_fixed * D * 4);
    b.adj = (float *)(p + off); off += adj_b;
    b.xt = (float *)(p + off); off += xt_b;
    b.xt2 = (float *)(p + off); off += xt_b;
    b.y = (float *)(p + off); off += y_b;
    b.tab = (float *)(p + off); off += y_b + al256((size_t)D * 4);      // (+ the all-zero row the in-GEMM gather reads for absent nodes)
    b.h_rows = (float *)(p + off); off += y_b;
    b.n_act = (int32_t *)(p + off); off += al256((size_t)G * 4);
    b.old_of_new = (int32_t *)(p + off); off += al256((size_t)G * n_fixed * 4);
    b.rowinv = (float *)(p + off); off += al256((size_t)G * n_fixed * 4);
    b.pool_extra = (float *)(p + off); off += al256((size_t)G * kTableSlices * D * 4);
    b.pool_groups = (float *)(p + off); off += al256(((size_t)G * n_fixed / 32 + 2) * 2 * D * 4);
    b.pid = (int64_t *)(p + off); off += al256((size_t)G * n_fixed * 8);
    b.pvw = (float *)(p + off); off += al256((size_t)G * n_fixed * 4);
    b.amax = (unsigned *)(p + off); off += 512;
    b.stats = (float *)(p + off); off += al256((size_t)G * n_fixed * (D / G_BN) * 2 * 4);
    b.mr = (float *)(p + off); off += al256((size_t)G * n_fixed * 2 * 4);
    b.bytes = off;
    return b;
}

size_t gnn_tc_workspace_bytes(int G, int n_fixed, int D, int chunks)
{
    (void)chunks;
    return carve_tc(nullptr, G, n_fixed, D).bytes + 4096;
}

static int gemm_ctas()
{
    static int v = 0;
    if (v == 0) {
        const char *e = getenv("SCHEMANET_GEMM_CTAS");
        v = (e && e[0] == '1') ? 1 : 2;
    }
    return v;
}

template <int EPI, int CTAS, int EW>
static int launch_gemm3x_n(const CUtensorMap *maps, const GemmTcArgs &a, const char *name, cudaStream_t st)
{
    using P = GemmPlan<CTAS, EW>;
    static bool configured = false;
    if (!configured) {
        SH_CHECK_CUDA(cudaFuncSetAttribute(gemm3x_kernel<EPI, CTAS, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::kTotal));
        configured = true;
    }
    const int units = a.G * ceil_div(a.M_total, G_BM * CTAS) * (a.N_total / G_BN);
    const int num_units = min(units, sm_count() / CTAS);
    static const int dbg = [] { const char *e = getenv("SCHEMANET_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
    GemmTcArgs a2 = a;
    a2.debug = dbg;
    // SCHEMANET_GEMM_TRACE=1: per-role wait / work cycle counters of every launch, printed after a device synchronisation
    static const bool tracing = getenv("SCHEMANET_GEMM_TRACE") != nullptr;
    static long long *trace_buf = nullptr;
    if (tracing) {
        if (!trace_buf) SH_CHECK_CUDA(cudaMalloc(&trace_buf, sizeof(long long) * 8 * 1024));
        SH_CHECK_CUDA(cudaMemsetAsync(trace_buf, 0, sizeof(long long) * 8 * 1024, st));
        a2.trace = trace_buf;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(num_units * CTAS);
    cfg.blockDim = dim3(32 * (2 + EW + G_XF_WARPS));
    cfg.dynamicSmemBytes = P::kTotal;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    prof_begin(name, st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm3x_kernel<EPI, CTAS, EW>, maps[0], maps[1], a2);
    prof_end(st);
    if (e != cudaSuccess) { set_error("%s launch -> %s", name, cudaGetErrorString(e)); return 1; }
    SH_CHECK_LAUNCH();
    if (tracing) {
        static long long host[8 * 1024];
        SH_CHECK_CUDA(cudaStreamSynchronize(st));
        SH_CHECK_CUDA(cudaMemcpy(host, trace_buf, sizeof(long long) * 8 * num_units * CTAS, cudaMemcpyDeviceToHost));
        double m[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tmin = 1e30, tmax = 0;
        for (int c = 0; c < num_units * CTAS; ++c) {
            for (int j = 0; j < 8; ++j) m[j] += (double)host[c * 8 + j] / (num_units * CTAS);
            tmin = host[c * 8 + 7] < tmin ? (double)host[c * 8 + 7] : tmin;
            tmax = host[c * 8 + 7] > tmax ? (double)host[c * 8 + 7] : tmax;
        }
        {
            static long long h2[4 * 512];
            SH_CHECK_CUDA(cudaMemcpy(h2, trace_buf + 4096, sizeof(long long) * 4 * num_units * CTAS, cudaMemcpyDeviceToHost));
            double pro = 0; long long g0 = (1LL << 62), g0max = 0, g1 = 0, g1min = (1LL << 62);
            for (int c = 0; c < num_units * CTAS; ++c) {
                pro += (double)h2[c * 4] / (num_units * CTAS);
                g0 = h2[c * 4 + 1] < g0 ? h2[c * 4 + 1] : g0; g0max = h2[c * 4 + 1] > g0max ? h2[c * 4 + 1] : g0max;
                g1 = h2[c * 4 + 2] > g1 ? h2[c * 4 + 2] : g1; g1min = h2[c * 4 + 2] < g1min ? h2[c * 4 + 2] : g1min;
            }
            fprintf(stderr, "gemm trace %-18s main loops min %.1f max %.1f kcyc | prologue %.1f kcyc | globaltimer: first entry -> last exit %.1f us, "
                            "entry spread %.1f us, exit spread %.1f us\n", name, tmin / 1e3, tmax / 1e3, pro / 1e3, (g1 - g0) / 1e3,
                    (g0max - g0) / 1e3, (g1 - g1min) / 1e3);
        }
        fprintf(stderr, "gemm trace %-18s units %4d ctas %3d | kcycles per CTA: total %.1f | producer wait-empty %.1f | transform wait-full %.1f "
                        "work %.1f | mma(leader, x2) wait-ready %.1f wait-tmem %.1f | epilogue wait-full %.1f work %.1f\n",
                name, units, num_units * CTAS, m[7] / 1e3, m[0] / 1e3, m[1] / 1e3, m[2] / 1e3, 2 * m[3] / 1e3, 2 * m[4] / 1e3, m[5] / 1e3, m[6] / 1e3);
    }
    return 0;
}

// maps: {A, B} for one CTA per row block, maps2: the same with 128-row B boxes for CTA pairs
template <int EPI, int EW = 4>
static int launch_gemm3x(const CUtensorMap *maps, const CUtensorMap *maps2, const GemmTcArgs &a, const char *name, cudaStream_t st)
{
    if (gemm_ctas() == 2) return launch_gemm3x_n<EPI, 2, EW>(maps2, a, name, st);
    return launch_gemm3x_n<EPI, 1, EW>(maps, a, name, st);
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

// 2-D map over a row-major fp32 table [rows, cols] with a box of {box_cols, 1} and no swizzle: the form tile::gather4 loads take
static int tmap_rows(CUtensorMap *m, const float *p, uint64_t cols, uint64_t rows, uint32_t box_cols)
{
    EncodeTiledFn fn = encode_tiled_fn();
    SH_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 4};
    cuuint32_t box[2] = {box_cols, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(p), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SH_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (gather) failed with CUresult %d", (int)r);
    return 0;
}

// All GNN layers given a prepared adjacency in b.adj.  k_sizes: active size per graph for the adjacency GEMM
// (null = n_fixed); identity_tail: see GemmTcArgs; row_sizes: real rows per graph for masking (null = all).
static bool layer0_fused(const sh_gnn_params *p, int G, int n_fixed)
{
    return p->embed_dim == G_BN && (int64_t)(p->num_codes + 1) <= (int64_t)G * n_fixed;
}
// embed_dim = 256 with the layer-0 shortcut: the B operand of the layer-0 adjacency GEMM (rows of P_0 by node code) is gathered by
// the GEMM's producer (TMA tile::gather4) instead of being materialised as X_0^T by embed_gather_t_kernel.
// SCHEMANET_EMBED_GATHER_KERNEL=1 keeps the kernel (A/B).
static bool narrow_gather(const sh_gnn_params *p, int G, int n_fixed)
{
    static const bool off = getenv("SCHEMANET_EMBED_GATHER_KERNEL") != nullptr;
    return !off && layer0_fused(p, G, n_fixed) && n_fixed <= kGatherMaxNodes;
}
// embed_dim > 256 (ImageNet configuration: 1024): the same layer-0 shortcut -- the first Linear is applied to the (M+1)-row
// embedding table, on the tensor cores -- and LayerNorm applied by the consumers of z (EPI_Z_*_STATS), see run_layers_wide
static bool wide_fused(const sh_gnn_params *p, int G, int n_fixed)
{
    static const bool off = getenv("SCHEMANET_WIDE_UNFUSED") != nullptr;
    return !off && p->embed_dim > G_BN && (int64_t)(p->num_codes + 1) <= (int64_t)G * n_fixed;
}
static int tmap3(CUtensorMap *m, const float *p, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t ld, uint64_t bstride,
                 uint32_t box_rows);

// The (M+1)-row table work of a forward (layer-0 product P_0, the activation tables of the class side, the pooled share of
// its pruned vertices) depends on the parameters only -- not on the adjacency operand -- and consists of small,
// latency-bound kernels.  It is forked onto an auxiliary stream so that it runs under the HBM-bound operand preparation
// (atlas pass, adjacency prep) instead of in front of the GEMMs, and joined before the first consumer (embed_gather).
// Fork/join are event dependencies, so the pattern is also valid inside a CUDA-graph capture of `st`.
struct AuxLane {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static AuxLane *aux_lane(int which)   // 0: instance-side forwards, 1: class-side forwards (no false dependencies between them)
{
    static AuxLane lanes[64][2];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    AuxLane &l = lanes[dev][which];
    if (l.stream == nullptr) {
        if (cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    return &l;
}

// table_tail (class side, needs layer0_fused): rows >= row_sizes[g] are pruned vertices; they get no GEMM rows, their
// pooled contribution comes from the per-code activation tables (see pool_table_rows_kernel).
// Launches the table work after everything already enqueued on `st` (ids / vertex_w / row_sizes must be ready by then) and
// returns with it in flight on the auxiliary stream; tables_join() makes `st` wait for it.
static int tables_begin(const sh_gnn_params *p, int G, int n_fixed, const int32_t *row_sizes, const int64_t *ids, int ld_ids,
                        const float *vertex_w, const TcBuffers &b, cudaStream_t st, bool table_tail)
{
    const int D = p->embed_dim;
    SH_REQUIRE(p->num_layers <= kMaxTcLayers, "gnn: at most %d layers on the tensor-core path", kMaxTcLayers);
    // operand maxima of this forward: reset, then the Linear weights' (every other slot is written by the kernel that
    // produces the operand)
    SH_CHECK_CUDA(cudaMemsetAsync(b.amax, 0, 512, st));
    // row M + 1 of the table buffer: zeros, what the in-GEMM gather of the layer-0 operand reads for nodes past a graph's size
    if (layer0_fused(p, G, n_fixed) || wide_fused(p, G, n_fixed))       // (otherwise there is no table: it may not even fit)
        SH_CHECK_CUDA(cudaMemsetAsync(b.tab + (size_t)(p->num_codes + 1) * D, 0, (size_t)D * 4, st));
    auto param_bounds = [&](cudaStream_t s_) -> int {
        ParamPtrs pp{};
        for (int l = 0; l < p->num_layers; ++l) { pp.w[l] = p->lin_w[l]; pp.gamma[l] = p->ln_w[l]; pp.beta[l] = p->ln_b[l]; }
        pp.emb = p->embedding; pp.fc = p->fc_w; pp.emb_n = (long long)(p->num_codes + 1) * D; pp.layers = p->num_layers;
        const bool wide_tables = D > G_BN && wide_fused(p, G, n_fixed);       // (only that path multiplies the table / fc on the tensor cores)
        SH_LAUNCH("gnn_param_bounds", s_, param_bounds_kernel<<<dim3(wide_tables ? 64 : 16, p->num_layers + (wide_tables ? 2 : 0)), 256, 0, s_>>>(pp, D, b.amax));
        SH_CHECK_LAUNCH();
        return 0;
    };
    // (with an auxiliary lane the bounds are its first kernel: their consumers are the GEMMs, which wait for the lane anyway,
    // and the operand-preparation chain on `st` starts 6-8 us earlier)
    const bool bounds_on_lane = !wide_fused(p, G, n_fixed) && layer0_fused(p, G, n_fixed);
    if (!bounds_on_lane && param_bounds(st)) return 1;
    if (wide_fused(p, G, n_fixed)) {
        // P_0 = Emb W_0^T into tab on the tensor cores (no bias here: b_0 is added after the adjacency product), then X_0^T = the
        // rows of P_0 gathered by node id
        SH_REQUIRE((((uintptr_t)p->embedding | (uintptr_t)p->lin_w[0]) & 15) == 0, "gnn: parameters must be 16-byte aligned for TMA");
        const int rows = p->num_codes + 1;
        CUtensorMap em, wm, wm2;
        if (tmap3(&em, p->embedding, D, rows, 1, D, 0, G_BM)) return 1;
        if (tmap3(&wm, p->lin_w[0], D, D, 1, D, 0, G_BN)) return 1;
        if (tmap3(&wm2, p->lin_w[0], D, D, 1, D, 0, G_BN / 2)) return 1;
        GemmTcArgs c{};
        c.G = 1; c.rows_per_graph = rows; c.M_total = rows; c.K_total = D; c.N_total = D; c.batched_b = 0;
        c.out_rows = b.tab; c.amax_a = b.amax + AM_EMB; c.amax_b = b.amax + AM_W;
        c.amax_out = b.amax + AM_X;        // the largest table entry bounds every gathered X_0 entry
        CUtensorMap m[2] = {em, wm}, mp[2] = {em, wm2};
        if (launch_gemm3x<EPI_BIAS_ROWS>(m, mp, c, "gnn_embed_table_tc", st)) return 1;
        if (n_fixed > kGatherMaxNodes) {      // (the in-GEMM gather keeps a graph's node codes in shared memory)
            dim3 grid2(ceil_div(b.ldk, 32), ceil_div(D, 256), G);
            SH_LAUNCH("gnn_embed_gather", st, embed_gather_t_kernel<<<grid2, 256, 0, st>>>(b.tab, ids, ld_ids, row_sizes, n_fixed, b.ldk, D, b.xt, b.amax + AM_X));
            SH_CHECK_LAUNCH();
        }
        // otherwise no gather kernel: the layer-0 adjacency GEMM reads the table rows itself (run_layers_wide)
        return 0;
    }
    if (!layer0_fused(p, G, n_fixed)) return 0;
    static const bool serial = getenv("SCHEMANET_TABLES_INLINE") != nullptr;
    AuxLane *lane = (serial || prof_on()) ? nullptr : aux_lane(table_tail ? 1 : 0);   // profiling: a kernel's time is its own
    cudaStream_t ts = lane ? lane->stream : st;
    if (lane) {
        SH_CHECK_CUDA(cudaEventRecord(lane->fork, st));
        SH_CHECK_CUDA(cudaStreamWaitEvent(ts, lane->fork, 0));
    }
    if (param_bounds(ts)) return 1;
    // P_0 = Emb W_0^T into tab; with a table tail the same launch also emits T_0 = relu(LN(P_0 + b_0)) into h_rows
    if (launch_rows_linear(p->embedding, p->lin_w[0], p->num_codes + 1, D, b.tab, ts, p->lin_b[0], p->ln_w[0], p->ln_b[0],
                           p->ln_eps, table_tail ? b.h_rows : nullptr, narrow_gather(p, G, n_fixed) ? b.amax + AM_X : nullptr))
        return 1;
    if (table_tail) {
        // T_l ping-pongs between h_rows and y: both are free until the GEMMs reach them, and the pooled sums of the
        // pruned vertices are taken (into pool_extra) before that
        const int rows = p->num_codes + 1;
        float *cur = b.h_rows, *nxt = b.y;
        for (int l = 1; l < p->num_layers; ++l) {
            if (launch_rows_linear(cur, p->lin_w[l], rows, D, nullptr, ts, p->lin_b[l], p->ln_w[l], p->ln_b[l], p->ln_eps, nxt))
                return 1;
            float *t = cur; cur = nxt; nxt = t;
        }
        static bool pool_smem_set = false;           // (Vc up to 65535: 64 KB of codes + weights per slice, above the 48 KB default)
        if (!pool_smem_set) {
            SH_CHECK_CUDA(cudaFuncSetAttribute(pool_table_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            pool_smem_set = true;
        }
        SH_LAUNCH("gnn_pool_table_rows", ts, pool_table_rows_kernel<<<dim3(G, kTableSlices), 256, (size_t)ceil_div(n_fixed, kTableSlices) * 8, ts>>>(cur, ids, vertex_w, row_sizes, n_fixed, D, b.pool_extra));
        SH_CHECK_LAUNCH();
    }
    if (!narrow_gather(p, G, n_fixed)) {
        // X_0^T of the fused layer 0 (rows of P_0 gathered by node id): L2-write bound, it also runs under the HBM-read-bound
        // operand preparation
        dim3 grid2(ceil_div(b.ldk, 32), ceil_div(D, 256), G);
        SH_LAUNCH("gnn_embed_gather", ts, embed_gather_t_kernel<<<grid2, 256, 0, ts>>>(b.tab, ids, ld_ids, row_sizes, n_fixed, b.ldk, D, b.xt, b.amax + AM_X));
        SH_CHECK_LAUNCH();
    }
    if (lane) SH_CHECK_CUDA(cudaEventRecord(lane->join, ts));
    return 0;
}

static int tables_join(const sh_gnn_params *p, int G, int n_fixed, cudaStream_t st, bool table_tail)
{
    if (!layer0_fused(p, G, n_fixed)) return 0;
    static const bool serial = getenv("SCHEMANET_TABLES_INLINE") != nullptr;
    AuxLane *lane = (serial || prof_on()) ? nullptr : aux_lane(table_tail ? 1 : 0);   // profiling: a kernel's time is its own
    if (lane) SH_CHECK_CUDA(cudaStreamWaitEvent(st, lane->join, 0));
    return 0;
}

// pooled[g, d] = (sum over chunks of partial[g, chunk, d]) / N   (gnn.py:96: mean over the padded length), + its maximum
__global__ void __launch_bounds__(256)
pool_finish_kernel(const float *__restrict__ partial, int chunks, int D, int n_fixed, const int32_t *__restrict__ mean_div,
                   float *__restrict__ pooled, unsigned *amax)
{
    const int g = blockIdx.x;
    const float div = (float)(mean_div ? *mean_div : n_fixed);
    float mx = 0.0f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.0f;
        for (int c = 0; c < chunks; ++c) acc += partial[((size_t)g * chunks + c) * D + d];
        acc /= div;
        pooled[(size_t)g * D + d] = acc;
        mx = fmaxf(mx, fabsf(acc));
    }
    record_amax(amax, mx);
}

// embed_dim > 256 with the layer-0 shortcut (wide_fused).  Per layer l (z_l = the pre-LayerNorm activations):
//   l = 0:  z_0 = Adj (P_0 rows) + b_0                                  one adjacency GEMM, epilogue stores z_0 (+ tile statistics)
//   l > 0:  Y = Adj relu(LN_{l-1}(z_{l-1}))                             adjacency GEMM, LayerNorm + ReLU applied to the B operand by the
//           z_l = Y W_l^T + b_l                                         transform warps; linear GEMM, epilogue stores z_l (+ statistics)
//   end:    pooled = sum_i relu(LN_{L-1}(z_{L-1}))_i w_i                 pool_ln_rows_kernel, straight from the z rows
// z is stored transposed (the next adjacency GEMM's K-major B operand) except for the last layer (rows, for the pooling).
static int run_layers_wide(const sh_gnn_params *p, int G, int n_fixed, const int32_t *k_sizes, int identity_tail,
                           const int32_t *row_sizes, const int64_t *ids, int ld_ids, const float *vertex_w, int ld_v, const TcBuffers &b,
                           int chunks, float *partial, cudaStream_t st, TcFinal *fin)
{
    const int D = p->embed_dim, ldk = b.ldk, nt = D / G_BN, L = p->num_layers;
    const int64_t rows = (int64_t)G * n_fixed;
    CUtensorMap adjm, ym;
    if (tmap3(&adjm, b.adj, n_fixed, n_fixed, G, ldk, (uint64_t)n_fixed * ldk, G_BM)) return 1;
    if (tmap3(&ym, b.y, D, (uint64_t)rows, 1, D, 0, G_BM)) return 1;
    float *xin = b.xt, *xout = b.xt2;
    auto merge = [&]() -> int {
        SH_LAUNCH("gnn_ln_stats_merge", st, ln_stats_merge_kernel<<<(unsigned)ceil_div64(rows, 256), 256, 0, st>>>(b.stats, rows, nt, p->ln_eps, b.mr));
        SH_CHECK_LAUNCH();
        return 0;
    };
    for (int l = 0; l < L; ++l) {
        const bool last = (l == L - 1);
        CUtensorMap xtm, xtm2;
        if (tmap3(&xtm, xin, n_fixed, D, G, ldk, (uint64_t)D * ldk, G_BN)) return 1;
        if (tmap3(&xtm2, xin, n_fixed, D, G, ldk, (uint64_t)D * ldk, G_BN / 2)) return 1;
        CUtensorMap m1[2] = {adjm, xtm}, m1p[2] = {adjm, xtm2};
        GemmTcArgs a{};
        a.G = G; a.rows_per_graph = n_fixed; a.M_total = n_fixed; a.K_total = n_fixed; a.N_total = D;
        a.k_sizes = k_sizes; a.identity_tail = identity_tail; a.batched_b = 1;
        a.amax_a = b.amax + AM_ADJ; a.ldk = ldk; a.stats = b.stats; a.eps = p->ln_eps;
        if (l == 0) {
            a.amax_b = b.amax + AM_X; a.row_sizes = row_sizes; a.bias = p->lin_b[0];
            if (n_fixed <= kGatherMaxNodes) {     // B = rows of the table P_0 gathered by node code inside the GEMM
                a.bg_table = b.tab; a.bg_ids = ids; a.bg_ld = ld_ids; a.bg_sizes = row_sizes;
                a.bg_zero_row = p->num_codes + 1;      // the explicit zero row behind the table (tables_begin)
                // the B map of this launch is the table itself, for the producer's TMA gather (box {B columns of a CTA, 1 row})
                if (tmap_rows(&m1[1], b.tab, D, p->num_codes + 2, G_BN)) return 1;
                if (tmap_rows(&m1p[1], b.tab, D, p->num_codes + 2, G_BN / 2)) return 1;
            }
            if (last) { a.out_rows = b.h_rows; if (launch_gemm3x<EPI_Z_ROWS_STATS>(m1, m1p, a, "gnn_adj_z_tc", st)) return 1; }
            else { a.out_t = xout; if (launch_gemm3x<EPI_Z_T_STATS>(m1, m1p, a, "gnn_adj_z_tc", st)) return 1; }
            if (merge()) return 1;
            float *t = xin; xin = xout; xout = t;
            continue;
        }
        // Y = Adj relu(LN(z_{l-1})): the B operand is normalised while it is converted
        a.bln_mr = b.mr; a.bln_gamma = p->ln_w[l - 1]; a.bln_beta = p->ln_b[l - 1];
        a.amax_b = b.amax + AM_LN + (l - 1);
        a.out_rows = b.y; a.amax_out = b.amax + AM_Y + l;
        if (launch_gemm3x<EPI_STORE_ROWS>(m1, m1p, a, "gnn_adj_gemm_tc", st)) return 1;
        // z_l = Y W_l^T + b_l
        SH_REQUIRE(((uintptr_t)p->lin_w[l] & 15) == 0, "gnn: Linear weights must be 16-byte aligned for TMA");
        CUtensorMap wm, wm2;
        if (tmap3(&wm, p->lin_w[l], D, D, 1, D, 0, G_BN)) return 1;
        if (tmap3(&wm2, p->lin_w[l], D, D, 1, D, 0, G_BN / 2)) return 1;
        GemmTcArgs c{};
        c.G = 1; c.rows_per_graph = n_fixed; c.M_total = (int)rows; c.K_total = D; c.N_total = D; c.row_sizes = row_sizes; c.batched_b = 0;
        c.bias = p->lin_b[l]; c.eps = p->ln_eps; c.ldk = ldk; c.stats = b.stats;
        c.amax_a = b.amax + AM_Y + l; c.amax_b = b.amax + AM_W + l;
        CUtensorMap m2[2] = {ym, wm}, m2p[2] = {ym, wm2};
        if (last) { c.out_rows = b.h_rows; if (launch_gemm3x<EPI_Z_ROWS_STATS>(m2, m2p, c, "gnn_linear_z_tc", st)) return 1; }
        else { c.out_t = xin; if (launch_gemm3x<EPI_Z_T_STATS>(m2, m2p, c, "gnn_linear_z_tc", st)) return 1; }
        if (merge()) return 1;
    }
    dim3 grid(chunks, G);
    SH_LAUNCH("gnn_pool_ln_rows", st, pool_ln_rows_kernel<<<grid, 256, 0, st>>>(b.h_rows, b.mr, p->ln_w[L - 1], p->ln_b[L - 1], vertex_w, ld_v,
                                                                                row_sizes, n_fixed, D, chunks, partial));
    SH_CHECK_LAUNCH();
    if (fin != nullptr && G >= 64) {
        // out = fc(pooled / N)  (gnn.py:96-97) as one more tensor-core GEMM: [G, D] x fc_w^T (a CTA per graph re-reading the
        // D x D weights, pool_fc_kernel, costs 0.4 ms at G = 1000, D = 1024)
        SH_LAUNCH("gnn_pool_finish", st, pool_finish_kernel<<<G, 256, 0, st>>>(partial, chunks, D, n_fixed, fin->mean_div, b.y, b.amax + AM_POOL));
        SH_CHECK_LAUNCH();
        SH_REQUIRE(((uintptr_t)p->fc_w & 15) == 0, "gnn: fc weights must be 16-byte aligned for TMA");
        CUtensorMap pm, fm, fm2;
        if (tmap3(&pm, b.y, D, (uint64_t)G, 1, D, 0, G_BM)) return 1;
        if (tmap3(&fm, p->fc_w, D, D, 1, D, 0, G_BN)) return 1;
        if (tmap3(&fm2, p->fc_w, D, D, 1, D, 0, G_BN / 2)) return 1;
        GemmTcArgs c{};
        c.G = 1; c.rows_per_graph = G; c.M_total = G; c.K_total = D; c.N_total = D; c.batched_b = 0;
        c.bias = p->fc_b; c.out_rows = fin->out; c.amax_a = b.amax + AM_POOL; c.amax_b = b.amax + AM_FC;
        CUtensorMap m[2] = {pm, fm}, mp[2] = {pm, fm2};
        if (launch_gemm3x<EPI_BIAS_ROWS>(m, mp, c, "gnn_fc_tc", st)) return 1;
        fin->done = true;
    }
    return 0;
}

// tables_begin() must have been called on `st` before (after the kernels that produce ids / vertex_w / row_sizes).
static int run_layers_tc(const sh_gnn_params *p, int G, int n_fixed, const int32_t *k_sizes, int identity_tail,
                         const int32_t *row_sizes, const int64_t *ids, int ld_ids, const float *vertex_w, int ld_v,
                         const TcBuffers &b, int chunks, float *partial, cudaStream_t st, bool table_tail = false,
                         TcFinal *fin = nullptr)
{
    const int D = p->embed_dim, ldk = b.ldk;
    // Layer-0 shortcut (embed_dim 256): (Adj X0) W0^T = Adj (X0 W0^T) and X0 = Emb[ids], so the first Linear is applied to
    // the (M+1)-row embedding TABLE once (a tiny fp32 GEMM) instead of to every node of every graph; layer 0 then is a
    // single adjacency GEMM with bias + LayerNorm + ReLU fused in its epilogue.  The table product is staged in the
    // `tab` buffer (as large as Y), which it fits whenever the batch has at least M+1 node slots.
    if (wide_fused(p, G, n_fixed))
        return run_layers_wide(p, G, n_fixed, k_sizes, identity_tail, row_sizes, ids, ld_ids, vertex_w, ld_v, b, chunks, partial, st, fin);
    const bool fuse0 = layer0_fused(p, G, n_fixed);
    SH_REQUIRE(!table_tail || (fuse0 && row_sizes && !identity_tail), "run_layers_tc: table tail needs the fused layer 0");
    const float *table = fuse0 ? b.tab : p->embedding;
    if (tables_join(p, G, n_fixed, st, table_tail)) return 1;
    if (!fuse0) {      // (fused layer 0: already gathered from P_0 by tables_begin)
        dim3 grid2(ceil_div(ldk, 32), ceil_div(D, 256), G);
        SH_LAUNCH("gnn_embed_gather", st, embed_gather_t_kernel<<<grid2, 256, 0, st>>>(table, ids, ld_ids, row_sizes, n_fixed, ldk, D, b.xt, b.amax + AM_X));
        SH_CHECK_LAUNCH();
    }
    CUtensorMap adjm, ym;
    if (tmap3(&adjm, b.adj, n_fixed, n_fixed, G, ldk, (uint64_t)n_fixed * ldk, G_BM)) return 1;
    if (tmap3(&ym, b.y, D, (uint64_t)G * n_fixed, 1, D, 0, G_BM)) return 1;
    // node features ping-pong between two X^T buffers (a fused-LayerNorm adjacency GEMM must not overwrite its own B)
    float *xin = b.xt, *xout = b.xt2;
    bool pooled_in_epilogue = false;

    for (int l = 0; l < p->num_layers; ++l) {
        const bool last = (l == p->num_layers - 1);
        CUtensorMap xtm, xtm2;            // B operand of the adjacency GEMM: whole 256-row tile / half of it (CTA pairs)
        if (tmap3(&xtm, xin, n_fixed, D, G, ldk, (uint64_t)D * ldk, G_BN)) return 1;
        if (tmap3(&xtm2, xin, n_fixed, D, G, ldk, (uint64_t)D * ldk, G_BN / 2)) return 1;
        CUtensorMap m1[2] = {adjm, xtm};
        CUtensorMap m1p[2] = {adjm, xtm2};
        GemmTcArgs a{};
        a.G = G; a.rows_per_graph = n_fixed; a.M_total = n_fixed; a.K_total = n_fixed; a.N_total = D;
        a.k_sizes = k_sizes; a.identity_tail = identity_tail; a.batched_b = 1; a.skip_masked = table_tail ? 1 : 0;
        a.amax_a = b.amax + AM_ADJ; a.amax_b = b.amax + AM_X + l;
        if (l == 0 && fuse0) {
            // H1 = relu(LN(Adj (X0 W0^T) + b0)) in one kernel
            a.row_sizes = row_sizes;
            if (narrow_gather(p, G, n_fixed)) {
                // B = rows of P_0 gathered by node code by the producer (nodes >= row_sizes[g]: the table's zero row)
                a.bg_table = b.tab; a.bg_ids = ids; a.bg_ld = ld_ids; a.bg_sizes = row_sizes;
                a.bg_zero_row = p->num_codes + 1;      // the explicit zero row behind the table (tables_begin)
                if (tmap_rows(&m1[1], b.tab, D, p->num_codes + 2, G_BN)) return 1;
                if (tmap_rows(&m1p[1], b.tab, D, p->num_codes + 2, G_BN / 2)) return 1;
            }
            a.bias = p->lin_b[0]; a.gamma = p->ln_w[0]; a.beta = p->ln_b[0]; a.eps = p->ln_eps;
            a.out_t = xout; a.ldk = ldk; a.out_rows = b.h_rows; a.amax_out = b.amax + AM_X + l + 1;
            if (last) { if (launch_gemm3x<EPI_LN_RELU_ROWS, 8>(m1, m1p, a, "gnn_adj_ln_tc", st)) return 1; }
            else { if (launch_gemm3x<EPI_LN_RELU_T, 8>(m1, m1p, a, "gnn_adj_ln_tc", st)) return 1; }
            float *t = xin; xin = xout; xout = t;
            continue;
        }
        // Y = Adj X
        a.out_rows = b.y; a.amax_out = b.amax + AM_Y + l;
        if (launch_gemm3x<EPI_STORE_ROWS, 8>(m1, m1p, a, "gnn_adj_gemm_tc", st)) return 1;
        // H = relu(LN(Y W^T + b)); the weight matrix is read where it lies (split in shared memory like every operand)
        SH_REQUIRE(((uintptr_t)p->lin_w[l] & 15) == 0, "gnn: Linear weights must be 16-byte aligned for TMA");
        CUtensorMap wm, wm2;
        if (tmap3(&wm, p->lin_w[l], D, D, 1, D, 0, G_BN)) return 1;
        if (tmap3(&wm2, p->lin_w[l], D, D, 1, D, 0, G_BN / 2)) return 1;
        GemmTcArgs c{};
        c.G = 1; c.rows_per_graph = n_fixed; c.M_total = G * n_fixed; c.K_total = D; c.N_total = D; c.row_sizes = row_sizes; c.batched_b = 0;
        c.skip_masked = table_tail ? 1 : 0;
        c.bias = p->lin_b[l]; c.gamma = p->ln_w[l]; c.beta = p->ln_b[l]; c.eps = p->ln_eps;
        c.out_t = xin; c.ldk = ldk; c.out_rows = b.h_rows;
        c.amax_a = b.amax + AM_Y + l; c.amax_b = b.amax + AM_W + l; c.amax_out = b.amax + AM_X + l + 1;
        CUtensorMap m2[2] = {ym, wm};
        CUtensorMap m2p[2] = {ym, wm2};
        if (D == G_BN) {
            if (last) {
                // weighted pooling fused into the epilogue: H of the last layer is never written
                c.pool_w = vertex_w; c.ld_w = ld_v; c.pool_groups = b.pool_groups;
                pooled_in_epilogue = true;
                if (launch_gemm3x<EPI_LN_RELU_ROWS, 8>(m2, m2p, c, "gnn_linear_ln_tc", st)) return 1;
            }
            else { if (launch_gemm3x<EPI_LN_RELU_T, 8>(m2, m2p, c, "gnn_linear_ln_tc", st)) return 1; }
        } else {
            // wide embeddings: bias in the GEMM epilogue, LayerNorm + ReLU (+ transpose) as a separate pass
            if (launch_gemm3x<EPI_BIAS_ROWS>(m2, m2p, c, "gnn_linear_tc", st)) return 1;
            dim3 grid(ceil_div(ldk, 32), G);
            if (last) {
                SH_LAUNCH("gnn_ln_relu_wide", st, ln_relu_wide_kernel<false><<<grid, 256, 0, st>>>(b.h_rows, row_sizes, n_fixed, ldk, D, p->ln_w[l],
                                                                                                  p->ln_b[l], p->ln_eps, nullptr, nullptr));
            } else {
                const size_t smem = (size_t)32 * (D + 1) * sizeof(float);
                static bool configured = false;
                if (!configured) {
                    SH_CHECK_CUDA(cudaFuncSetAttribute(ln_relu_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * (kMaxDim + 1) * 4));
                    configured = true;
                }
                SH_LAUNCH("gnn_ln_relu_wide", st, ln_relu_wide_kernel<true><<<grid, 256, smem, st>>>(b.h_rows, row_sizes, n_fixed, ldk, D, p->ln_w[l],
                                                                                                    p->ln_b[l], p->ln_eps, xin, b.amax + AM_X + l + 1));
            }
            SH_CHECK_LAUNCH();
        }
    }
    if (pooled_in_epilogue && fin != nullptr) {
        SH_LAUNCH("gnn_pool_fc", st,
                  pool_groups_fc_kernel<<<dim3(G, G >= 1024 ? 1 : (G >= 192 ? 4 : 8)), 256, (size_t)D * sizeof(float), st>>>(
                      b.pool_groups, row_sizes, n_fixed, D, table_tail ? b.pool_extra : nullptr, fin->mean_div, p->fc_w, p->fc_b, fin->out));
        fin->done = true;
    } else if (pooled_in_epilogue) {
        SH_LAUNCH("gnn_pool_rows", st, pool_groups_reduce_kernel<<<G, 256, 0, st>>>(b.pool_groups, row_sizes, n_fixed, D, chunks, partial,
                                                                                     table_tail ? b.pool_extra : nullptr));
    } else {
        dim3 grid(chunks, G);
        SH_LAUNCH("gnn_pool_rows", st, pool_rows_kernel<<<grid, 256, 0, st>>>(b.h_rows, vertex_w, ld_v, row_sizes, n_fixed, D, chunks, partial,
                                                                              table_tail ? b.pool_extra : nullptr));
    }
    SH_CHECK_LAUNCH();
    return 0;
}

int gnn_forward_tc(const sh_gnn_params *p, int G, int n_fixed, const int32_t *sizes, const int64_t *ids,
                   const float *vertex_w, int ld_v, const float *edges, int64_t edge_batch_stride, int edge_ld, int chunks,
                   float *partial, void *workspace, cudaStream_t st, TcFinal *fin)
{
    TcBuffers b = carve_tc(workspace, G, n_fixed, p->embed_dim);
    if (tables_begin(p, G, n_fixed, sizes, ids, ld_v, vertex_w, b, st, false)) return 1;
    if (getenv("SCHEMANET_ADJ_TILED") != nullptr) {
        dim3 grid(ceil_div(b.ldk, 32), ceil_div(n_fixed, 32), G);
        SH_LAUNCH("gnn_adj_prep", st, adj_prep_kernel<<<grid, 256, 0, st>>>(edges, edge_batch_stride, edge_ld, sizes, n_fixed, b.ldk, b.adj, b.amax + AM_ADJ));
    } else {
        const int tr = ceil_div(n_fixed, 32);
        const int cpg = max(1, min(16, ceil_div(tr * (tr + 1) / 2 + tr, 4 * 4)));   // ~4 items per warp at full size
        SH_LAUNCH("gnn_adj_prep", st, adj_sym_kernel<<<G * cpg, 128, 0, st>>>(edges, edge_batch_stride, edge_ld, sizes, n_fixed, b.ldk, b.adj, cpg, b.amax + AM_ADJ));
    }
    SH_CHECK_LAUNCH();
    return run_layers_tc(p, G, n_fixed, sizes, 0, sizes, ids, ld_v, vertex_w, ld_v, b, chunks, partial, st, false, fin);
}

// The class graphs' layers after the adjacency operand is in place.  embed_dim 256: pruned vertices are served from the
// per-code activation tables and the GEMMs only see the un-pruned ones; wider embeddings keep every vertex in the GEMMs
// and let the row blocks of pruned vertices visit just their identity diagonal.
static bool class_table_tail(const sh_gnn_params *p, int K, int Vc) { return layer0_fused(p, K, Vc); }

static int class_layers_tc(const sh_gnn_params *p, int K, int Vc, const TcBuffers &b, int chunks, float *partial, cudaStream_t st,
                           TcFinal *fin)
{
    if (class_table_tail(p, K, Vc))
        return run_layers_tc(p, K, Vc, b.n_act, 0, b.n_act, b.pid, Vc, b.pvw, Vc, b, chunks, partial, st, true, fin);
    return run_layers_tc(p, K, Vc, b.n_act, 1, nullptr, b.pid, Vc, b.pvw, Vc, b, chunks, partial, st, false, fin);
}

// Class side with the graphs compacted to their un-pruned vertices (see class_perm_kernel), from a materialised atlas.
int gnn_class_forward_tc(const sh_gnn_params *p, int K, int Vc, const float *class_vertices, const float *class_edges,
                         const int64_t *class_ingredients, float prune_threshold, int chunks, float *partial,
                         void *workspace, cudaStream_t st, TcFinal *fin)
{
    TcBuffers b = carve_tc(workspace, K, Vc, p->embed_dim);
    SH_REQUIRE(Vc <= 65535, "class side: Vc too large");
    const int prune = prune_threshold >= 0.0f ? 1 : 0;
    SH_LAUNCH("class_perm_kernel", st,
              class_perm_kernel<<<K, 1024, 0, st>>>(class_vertices, class_ingredients, Vc, prune_threshold, prune, b.n_act,
                                                    b.old_of_new, b.pid, b.pvw));
    SH_CHECK_LAUNCH();
    if (tables_begin(p, K, Vc, class_table_tail(p, K, Vc) ? b.n_act : nullptr, b.pid, Vc, b.pvw, b, st, class_table_tail(p, K, Vc))) return 1;   // (identity tail: every vertex keeps its GEMM row)
    SH_LAUNCH("class_adj_prep_kernel", st,
              class_adj_prep_kernel<<<dim3(ceil_div(b.ldk, 32), ceil_div(Vc, 32), K), 256, 0, st>>>(
                  class_edges, K, Vc, b.ldk, G_BM * gemm_ctas(), class_table_tail(p, K, Vc) ? 0 : 1, b.n_act, b.old_of_new,
                  b.adj, b.amax + AM_ADJ));
    SH_CHECK_LAUNCH();
    return class_layers_tc(p, K, Vc, b, chunks, partial, st, fin);
}

// Stage 3a + class side fused (sh_dev_class_side on the tensor-core path): the atlas pass leaves the per-row normalisers,
// and the adjacency operand of the un-pruned vertices is gathered from the (pruned) edge parameter itself, so the full
// [K, Vc, Vc] class_edges tensor is only written when the caller asks for it (class_edges != null) and never read back.
int gnn_class_side_tc(const sh_gnn_params *p, float *edge_weights, int K, int Vc, float prune_threshold, int prune_in_place,
                      int remove_self_loop, const float *class_vertices, float *class_edges,
                      const int64_t *class_ingredients, int chunks, float *partial, void *workspace, cudaStream_t st,
                      TcFinal *fin)
{
    TcBuffers b = carve_tc(workspace, K, Vc, p->embed_dim);
    SH_REQUIRE(Vc <= 65535, "class side: Vc too large");
    const int prune = prune_threshold >= 0.0f ? 1 : 0;
    SH_LAUNCH("class_perm_kernel", st,
              class_perm_kernel<<<K, 1024, 0, st>>>(class_vertices, class_ingredients, Vc, prune_threshold, prune, b.n_act,
                                                    b.old_of_new, b.pid, b.pvw));
    SH_CHECK_LAUNCH();
    if (tables_begin(p, K, Vc, class_table_tail(p, K, Vc) ? b.n_act : nullptr, b.pid, Vc, b.pvw, b, st, class_table_tail(p, K, Vc))) return 1;   // (identity tail: every vertex keeps its GEMM row)
    // (Measured and not kept, r02: ONE kernel with a cluster of 4 / 8 CTAs per class -- normalisers in phase 1, cluster barrier,
    // adjacency gathered from L2 in phase 2.  189-203 us against 90 + 77 us for the two kernels below: a class only gets its
    // cluster's share of the HBM bandwidth, HBM idles during phase 2, and 18-37 classes in flight leave the latency-bound gather
    // with a third of the warps.  profiles/r02_experiments.md)
    if (launch_class_edges(edge_weights, class_vertices, K, Vc, prune_threshold, prune_in_place, remove_self_loop, class_edges,
                           b.rowinv, st))
        return 1;
    {
        // a few tile pairs per warp for the largest work list (all vertices active: TR (TR + 1) / 2 pairs); r01 sweep at
        // K = 100, Vc = 1024: 12 CTAs/class 0.143 ms, 24: 0.133, 48: 0.122, 96: 0.123
        const int tr = ceil_div(Vc, 32);
        static const int forced = [] { const char *e = getenv("SCHEMANET_ADJ_CTAS"); return e ? atoi(e) : 0; }();
        int cpc = max(4, min(64, ceil_div(tr * (tr + 1) / 2, kAdjWarps * 2)));
        if (forced > 0) cpc = forced;
        SH_LAUNCH("class_adj_prep_kernel", st,
                  class_adj_raw_kernel<<<K * cpc, kAdjWarps * 32, 0, st>>>(edge_weights, b.rowinv, K, Vc, b.ldk, G_BM * gemm_ctas(),
                                                                           class_table_tail(p, K, Vc) ? 0 : 1, remove_self_loop,
                                                                           b.n_act, b.old_of_new, b.adj, cpc, b.amax + AM_ADJ));
        SH_CHECK_LAUNCH();
    }
    return class_layers_tc(p, K, Vc, b, chunks, partial, st, fin);
}

// Inner-product logits [B, K] = feat_instance [B, D] . feat_class [K, D]^T (match.py:29-31) as one tensor-core GEMM; K need not
// be a multiple of the 256-column tile (TMA zero-fills the missing class rows, the epilogue stores the columns that exist).
// scratch: 2 amax slots.
bool similarity_tc_supported(int B, int K, int D)
{
    return D % G_BK == 0 && D >= 64 && (int64_t)B * K * D >= (1LL << 27) && encode_tiled_fn() != nullptr;
}

int similarity_tc(const float *fi, const float *fk, int B, int K, int D, float *logits, unsigned *scratch, cudaStream_t st)
{
    SH_REQUIRE((((uintptr_t)fi | (uintptr_t)fk) & 15) == 0, "similarity: features must be 16-byte aligned for TMA");
    SH_CHECK_CUDA(cudaMemsetAsync(scratch, 0, 8, st));
    SH_LAUNCH("similarity_absmax", st, absmax_kernel<<<32, 256, 0, st>>>(fi, (int64_t)B * D, scratch));
    SH_CHECK_LAUNCH();
    SH_LAUNCH("similarity_absmax", st, absmax_kernel<<<32, 256, 0, st>>>(fk, (int64_t)K * D, scratch + 1));
    SH_CHECK_LAUNCH();
    const int n_pad = (K + G_BN - 1) / G_BN * G_BN;
    CUtensorMap am, bm, bm2;
    if (tmap3(&am, fi, D, (uint64_t)B, 1, D, 0, G_BM)) return 1;
    if (tmap3(&bm, fk, D, (uint64_t)K, 1, D, 0, G_BN)) return 1;
    if (tmap3(&bm2, fk, D, (uint64_t)K, 1, D, 0, G_BN / 2)) return 1;
    GemmTcArgs c{};
    c.G = 1; c.rows_per_graph = B; c.M_total = B; c.K_total = D; c.N_total = n_pad; c.n_valid = K; c.batched_b = 0;
    c.out_rows = logits; c.amax_a = scratch; c.amax_b = scratch + 1;
    SH_REQUIRE(n_pad <= kMaxDim, "similarity: more than %d classes on the tensor-core path", kMaxDim);
    CUtensorMap m[2] = {am, bm}, mp[2] = {am, bm2};
    return launch_gemm3x<EPI_BIAS_ROWS>(m, mp, c, "similarity_tc", st);
}

}  // namespace sh
