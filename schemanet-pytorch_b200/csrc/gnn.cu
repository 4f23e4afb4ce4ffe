// gnn.cu -- stage 3b: GNN embedding of instance / class graphs and the matcher similarity, fp32 CUDA-core path.
//
// Replaces GNN.forward / Layer.forward / GraphConv.forward (schema_inference/graph/gnn.py:20-98) and
// Matcher._inner_product/_cosine_sim/_euclidean_sim (schema_inference/graph/match.py:21-31).
//
// The reference pads every instance graph to the batch maximum N, materialises adj = E + E^T, a dense identity,
// adj/2 + I, then bmm -> Linear -> masked_fill -> LayerNorm -> ReLU as separate ATen ops.  Here:
//   * graphs are processed at their true size n_g (padded rows only ever contribute the divisor N of the final
//     mean, gnn.py:96 -- see DESIGN.md "padding algebra"), read straight from stage 2's packed slots;
//   * ((E+E^T)/2 + I) is formed on the fly while the A tile is staged in shared memory (E and E^T tiles are both
//     read coalesced), the embedding gather is fused into the first layer's B-tile load;
//   * bias add is fused in the GEMM epilogue; LayerNorm+ReLU is one warp-per-row pass with shuffle reductions; the
//     last layer's LN pass also produces the vertex-weighted pooling partials, reduced without atomics.
// All products accumulate in fp32 FMA (the north star's 1e-5 logit tolerance rules out plain TF32 here).
#include "common.cuh"
#include <stdlib.h>
#include "gnn_tc.cuh"

namespace sh {

enum { A_ROWMAJOR = 0, A_ADJ = 1 };
enum { B_ROWMAJOR = 0, B_TRANSPOSED = 1 };

struct GemmArgs {
    // C[g] (M_g x N) = A[g] (M_g x K_g) * B[g] (K_g x N)
    const float *A; int64_t a_batch; int lda;        // lda <= 0: compact graphs, lda = n_g
    const float *B; int64_t b_batch; int ldb;
    float *C; int64_t c_batch; int ldc;
    const float *bias;                                // [N] or null
    const int32_t *sizes;                             // [G] rows (and K for A_ADJ) per graph, or null
    const int64_t *b_gather; int ld_gather;           // B row k of graph g is B[b_gather[g*ld_gather + k]] (embedding)
    int M, N, K;                                      // upper bounds (fixed sizes when sizes == null)
    int G;
};

template <int BM, int BN, int BK, int TM, int TN, int AMODE, int BMODE>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(GemmArgs g)
{
    constexpr int NT = (BM / TM) * (BN / TN);
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int gi = blockIdx.z;
    const int n_g = g.sizes ? g.sizes[gi] : g.M;
    const int Mg = n_g;
    const int Kg = (AMODE == A_ADJ) ? n_g : g.K;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= Mg) return;
    const int lda = g.lda > 0 ? g.lda : n_g;
    const float *A = g.A + (size_t)gi * g.a_batch;
    const float *B = g.B + (size_t)gi * g.b_batch;
    const int64_t *gather = g.b_gather ? g.b_gather + (size_t)gi * g.ld_gather : nullptr;
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < Kg; k0 += BK) {
        // ---- A tile -> As[kk][mm]
        if (AMODE == A_ROWMAJOR) {
            for (int e = tid; e < BM * BK; e += NT) {
                const int mm = e / BK, kk = e % BK;
                const int m = m0 + mm, k = k0 + kk;
                As[kk][mm] = (m < Mg && k < Kg) ? A[(size_t)m * lda + k] : 0.0f;
            }
        } else {
            // adjacency (E + E^T)/2 + I (gnn.py:27-30): first E[m][k] (coalesced along k) ...
            for (int e = tid; e < BM * BK; e += NT) {
                const int mm = e / BK, kk = e % BK;
                const int m = m0 + mm, k = k0 + kk;
                As[kk][mm] = (m < Mg && k < Kg) ? A[(size_t)m * lda + k] : 0.0f;
            }
            __syncthreads();
            // ... then E[k][m] (coalesced along m), halve, add the identity
            for (int e = tid; e < BM * BK; e += NT) {
                const int kk = e / BM, mm = e % BM;
                const int m = m0 + mm, k = k0 + kk;
                const float et = (m < Mg && k < Kg) ? A[(size_t)k * lda + m] : 0.0f;
                As[kk][mm] = (As[kk][mm] + et) / 2.0f + ((m == k && m < Mg) ? 1.0f : 0.0f);
            }
        }
        // ---- B tile -> Bs[kk][nn]
        if (BMODE == B_ROWMAJOR) {
            for (int e = tid; e < BK * BN; e += NT) {
                const int kk = e / BN, nn = e % BN;
                const int k = k0 + kk, n = n0 + nn;
                float v = 0.0f;
                if (k < Kg && n < g.N) {
                    const size_t row = gather ? (size_t)gather[k] : (size_t)k;
                    v = B[row * g.ldb + n];
                }
                Bs[kk][nn] = v;
            }
        } else {
            for (int e = tid; e < BK * BN; e += NT) {
                const int nn = e / BK, kk = e % BK;
                const int k = k0 + kk, n = n0 + nn;
                Bs[kk][nn] = (k < Kg && n < g.N) ? B[(size_t)n * g.ldb + k] : 0.0f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    float *C = g.C + (size_t)gi * g.c_batch;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= Mg) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n < g.N) C[(size_t)m * g.ldc + n] = acc[i][j] + (g.bias ? g.bias[n] : 0.0f);
        }
    }
}

template <int AMODE, int BMODE>
static int launch_sgemm(const GemmArgs &g, cudaStream_t st, const char *name)
{
    if (g.M >= 256) {
        dim3 grid(ceil_div(g.N, 128), ceil_div(g.M, 128), g.G);
        SH_LAUNCH(name, st, sgemm_kernel<128, 128, 8, 8, 8, AMODE, BMODE><<<grid, 256, 0, st>>>(g));
    } else {
        dim3 grid(ceil_div(g.N, 64), ceil_div(g.M, 64), g.G);
        SH_LAUNCH(name, st, sgemm_kernel<64, 64, 16, 4, 4, AMODE, BMODE><<<grid, 256, 0, st>>>(g));
    }
    SH_CHECK_LAUNCH();
    return 0;
}

// LayerNorm + ReLU, one warp per node row (gnn.py:41-46; rows >= n_g are never produced: DESIGN.md padding algebra).
// kPool: instead of storing the activations, multiply by the vertex weight and emit per-(graph, row-chunk) partial
// sums for the weighted mean pooling (gnn.py:94-96).
template <bool kPool>
__global__ void __launch_bounds__(256)
ln_relu_kernel(const float *__restrict__ Z, float *__restrict__ Hout, int G, int n_fixed, const int32_t *__restrict__ sizes,
               int ld_rows, int D, const float *__restrict__ gamma, const float *__restrict__ beta, float eps,
               const float *__restrict__ vertex_w, int ld_v, float *__restrict__ pool_partial, int chunks)
{
    extern __shared__ float sm[];   // kPool: [warps][D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int gidx = blockIdx.y;
    const int n_g = sizes ? sizes[gidx] : n_fixed;
    const int chunk = blockIdx.x;                  // rows [chunk*rows_per_chunk, ...)
    const int rows_per_chunk = ceil_div(n_fixed, chunks);
    const int r_begin = chunk * rows_per_chunk;
    const int r_end = min(n_g, r_begin + rows_per_chunk);
    float *mypool = kPool ? sm + (size_t)warp * D : nullptr;
    if (kPool)
        for (int d = lane; d < D; d += kWarp) mypool[d] = 0.0f;
    for (int r = r_begin + warp; r < r_end; r += wpb) {
        const float *z = Z + ((size_t)gidx * ld_rows + r) * D;
        float s = 0.0f;
        for (int d = lane; d < D; d += kWarp) s += z[d];
        const float mean = warp_sum(s) / (float)D;
        float v = 0.0f;
        for (int d = lane; d < D; d += kWarp) { const float t = z[d] - mean; v += t * t; }
        const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)D + eps);
        const float wr = kPool ? vertex_w[(size_t)gidx * ld_v + r] : 0.0f;
        for (int d = lane; d < D; d += kWarp) {
            const float y = fmaxf((z[d] - mean) * rstd * gamma[d] + beta[d], 0.0f);
            if (kPool) mypool[d] += y * wr;
            else Hout[((size_t)gidx * ld_rows + r) * D + d] = y;
        }
    }
    if (kPool) {
        __syncthreads();
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            float t = 0.0f;
            for (int w = 0; w < wpb; ++w) t += sm[(size_t)w * D + d];
            pool_partial[((size_t)gidx * chunks + chunk) * D + d] = t;
        }
    }
}

// pooled[g][d] = (sum over chunks of partials) / N   (mean over the PADDED node count, gnn.py:96)
__global__ void pool_finish_kernel(const float *__restrict__ partial, int G, int chunks, int D, int n_fixed,
                                   const int32_t *__restrict__ mean_div, float *__restrict__ pooled)
{
    const int gidx = blockIdx.x;
    const float div = (float)(mean_div ? *mean_div : n_fixed);
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float t = 0.0f;
        for (int c = 0; c < chunks; ++c) t += partial[((size_t)gidx * chunks + c) * D + d];
        pooled[(size_t)gidx * D + d] = t / div;
    }
}

// match.py:21-31 on the expanded [B, K, D] pair; one warp per (b, k)
__global__ void __launch_bounds__(256)
similarity_kernel(const float *__restrict__ fi, const float *__restrict__ fk, int B, int K, int D, int kind,
                  float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t pairs = (int64_t)B * K;
    for (int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < pairs;
         p += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const float *a = fi + (p / K) * D, *b = fk + (p % K) * D;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        // 16 independent loads per lane in flight (the kernel is the tail of the step: two L2 round trips, not 2 D / 32),
        // folded in the order of the plain loop
        for (int d0 = lane; d0 < D; d0 += 8 * kWarp) {
            float xv[8], yv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int d = d0 + u * kWarp;
                xv[u] = d < D ? a[d] : 0.f;
                yv[u] = d < D ? b[d] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (d0 + u * kWarp >= D) break;
                const float x = xv[u], y = yv[u];
                if (kind == SH_SIM_EUCLIDEAN) { const float t = x - y; s0 += t * t; }
                else { s0 += x * y; s1 += x * x; s2 += y * y; }
            }
        }
        s0 = warp_sum(s0);
        if (kind == SH_SIM_COSINE) { s1 = warp_sum(s1); s2 = warp_sum(s2); }
        if (lane == 0) {
            float r;
            if (kind == SH_SIM_INNER_PRODUCT) r = s0;
            else if (kind == SH_SIM_COSINE) {
                // torch.cosine_similarity: x.y / max(|x| |y|, eps) with eps = 1e-8 ; match.py:22-23
                r = (s0 / fmaxf(sqrtf(s1) * sqrtf(s2), 1e-8f) + 1.0f) / 2.0f;
            } else r = 1.0f / (1.0f + sqrtf(s0));
            out[p] = r;
        }
    }
}

// pooled[g] = (sum of the chunk partials) / N  (mean over the PADDED node count, gnn.py:96), then out = fc(pooled)
// (gnn.py:97).  One CTA per graph: the pooled vector lives in shared memory, one warp per output feature reads a
// contiguous row of fc_w.
__global__ void __launch_bounds__(256)
pool_fc_kernel(const float *__restrict__ partial, int chunks, int D, int n_fixed, const int32_t *__restrict__ mean_div,
               const float *__restrict__ fc_w, const float *__restrict__ fc_b, float *__restrict__ out)
{
    extern __shared__ float pooled[];
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float div = (float)(mean_div ? *mean_div : n_fixed);
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float t = 0.0f;
        for (int c = 0; c < chunks; ++c) t += partial[((size_t)g * chunks + c) * D + d];
        pooled[d] = t / div;
    }
    __syncthreads();
    // blockIdx.y splits the output features so that small batches still fill the machine
    const int per = (D + gridDim.y - 1) / gridDim.y;
    const int o_end = min(D, (int)(blockIdx.y + 1) * per);
    for (int o = blockIdx.y * per + warp; o < o_end; o += (int)(blockDim.x >> 5)) {
        const float *w = fc_w + (size_t)o * D;
        float acc = 0.0f;
        for (int d = lane; d < D; d += kWarp) acc = fmaf(pooled[d], w[d], acc);
        acc = warp_sum(acc);
        if (lane == 0) out[(size_t)g * D + o] = acc + fc_b[o];
    }
}

// out[r, o] = sum_k A[r, k] W[o, k] for a tall-skinny A (the (M+1)-row embedding / activation tables), D <= 256, D % 32 == 0;
// optionally also act[r, :] = relu(LayerNorm(out[r, :] + bias)) (one layer of the per-code activation tables, gnn_tc.cu).
// (The generic tiled GEMM launches only ~18 CTAs for this shape and took 0.11 ms; a warp-per-output version with shuffle
// reductions 0.028 ms; thread-per-output streaming its own W row 0.018 ms, L1-tag bound on 32 lines per request.)
// kRows rows per CTA staged in shared memory; thread o owns output feature o for all kRows rows -- kRows independent FMA
// chains per thread, no reductions.  W streams through a 3-stage shared-memory ring in 32-column slabs by cp.async (16-byte
// pieces, rows padded to 36 floats: the 16-byte reads of a quarter-warp hit 32 distinct banks).  With the previous
// register-staged, one-slab-ahead pipeline every slab exposed an L2 round trip (8 slabs: 20 us for 134 MFLOP); two slabs
// are now in flight while one is used.  The products of an output are folded in ascending k as before.
constexpr int kRlPitch = 36, kRlStages = 3;
__device__ __forceinline__ void rl_cp_async16(float *dst_smem, const float *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
template <int kRows>
__global__ void __launch_bounds__(256)
rows_linear_kernel(const float *__restrict__ A, const float *__restrict__ W, int rows, int D, float *__restrict__ out,
                   const float *__restrict__ bias, const float *__restrict__ gamma, const float *__restrict__ beta, float eps,
                   float *__restrict__ act, unsigned *__restrict__ amax_out)
{
    extern __shared__ __align__(16) float rl_smem[];
    float (*As)[256] = reinterpret_cast<float (*)[256]>(rl_smem);                  // [kRows][256]
    float *Wring = rl_smem + kRows * 256;                                          // [kRlStages][256][kRlPitch]
    __shared__ float red[8][kRows];
    const int r0 = blockIdx.x * kRows, o = threadIdx.x, lane = o & 31, warp = o >> 5;
    // slab k0 of W -> stage: 256 rows x 8 pieces of 16 bytes, 8 pieces per thread (rows >= D: never read by a live thread)
    auto issue = [&](int k0, int stage) {
        float *dst = Wring + (size_t)stage * 256 * kRlPitch;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = j * 256 + o, row = c >> 3, piece = c & 7;
            if (row < D) rl_cp_async16(dst + row * kRlPitch + piece * 4, W + (size_t)row * D + k0 + piece * 4);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int slabs = D / 32;
    for (int sidx = 0; sidx < kRlStages - 1; ++sidx) {
        if (sidx < slabs) issue(sidx * 32, sidx);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    {
        float a[kRows];
#pragma unroll
        for (int i = 0; i < kRows; ++i) a[i] = (o < D && r0 + i < rows) ? __ldg(A + (size_t)(r0 + i) * D + o) : 0.0f;
#pragma unroll
        for (int i = 0; i < kRows; ++i) As[i][o] = a[i];
    }
    float acc[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) acc[i] = 0.0f;
    for (int it = 0; it < slabs; ++it) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kRlStages - 2) : "memory");      // slab `it` has landed (this thread's pieces)
        __syncthreads();                      // ... everyone's pieces; slab it - 1 is consumed by all (first pass: As visible)
        if (it + kRlStages - 1 < slabs) issue((it + kRlStages - 1) * 32, (it + kRlStages - 1) % kRlStages);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        const float *wrow = Wring + (size_t)(it % kRlStages) * 256 * kRlPitch + o * kRlPitch;
        const int k0 = it * 32;
        // the shared-memory reads of step k4 + 1 are issued before the FMAs of step k4 (two warps per scheduler do not hide
        // the read latency by themselves: 28 % of the kernel's stall samples were FMAs waiting for their operands)
        float4 wn = *reinterpret_cast<const float4 *>(wrow), avn[kRows];
#pragma unroll
        for (int i = 0; i < kRows; ++i) avn[i] = *reinterpret_cast<const float4 *>(&As[i][k0]);
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            const float4 w = wn;
            float4 av[kRows];
#pragma unroll
            for (int i = 0; i < kRows; ++i) av[i] = avn[i];
            if (k4 + 1 < 8) {
                wn = *reinterpret_cast<const float4 *>(wrow + 4 * (k4 + 1));
#pragma unroll
                for (int i = 0; i < kRows; ++i) avn[i] = *reinterpret_cast<const float4 *>(&As[i][k0 + 4 * (k4 + 1)]);
            }
#pragma unroll
            for (int i = 0; i < kRows; ++i) {
                acc[i] = fmaf(av[i].x, w.x, acc[i]);
                acc[i] = fmaf(av[i].y, w.y, acc[i]);
                acc[i] = fmaf(av[i].z, w.z, acc[i]);
                acc[i] = fmaf(av[i].w, w.w, acc[i]);
            }
        }
    }
    const bool live = o < D;
    if (out && live) {
#pragma unroll
        for (int i = 0; i < kRows; ++i)
            if (r0 + i < rows) out[(size_t)(r0 + i) * D + o] = acc[i];
    }
    if (amax_out != nullptr) {      // largest magnitude written to `out` (bit pattern; the consumer scales its operand by it)
        float mx = 0.0f;
#pragma unroll
        for (int i = 0; i < kRows; ++i)
            if (live && r0 + i < rows) mx = fmaxf(mx, fabsf(acc[i]));
        mx = warp_max(mx);
        if (lane == 0 && !(mx == 0.0f)) atomicMax(amax_out, __float_as_uint(mx));      // NaN compares above +inf
    }
    if (act == nullptr) return;
    // LayerNorm over the D features of each row (two-pass: mean, then centred variance), block reductions via `red`
    const float bo = live ? bias[o] : 0.0f;
    float mean[kRows], rstd[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
        acc[i] = live ? acc[i] + bo : 0.0f;
        const float s = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
        float s = 0.0f;
        for (int w = 0; w < 8; ++w) s += red[w][i];
        mean[i] = s / (float)D;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
        const float t = live ? acc[i] - mean[i] : 0.0f;
        const float s = warp_sum(t * t);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
        float s = 0.0f;
        for (int w = 0; w < 8; ++w) s += red[w][i];
        rstd[i] = 1.0f / sqrtf(s / (float)D + eps);
    }
    if (live) {
        const float g = gamma[o], be = beta[o];
#pragma unroll
        for (int i = 0; i < kRows; ++i)
            if (r0 + i < rows) act[(size_t)(r0 + i) * D + o] = fmaxf((acc[i] - mean[i]) * rstd[i] * g + be, 0.0f);
    }
}

int launch_rows_linear(const float *A, const float *W, int rows, int D, float *out, cudaStream_t st, const float *bias,
                       const float *gamma, const float *beta, float eps, float *act, unsigned *amax_out)
{
    SH_REQUIRE(D <= 256 && D % 32 == 0, "rows_linear: D <= 256, D %% 32 == 0 expected");
    SH_REQUIRE(out || act, "rows_linear: no output");
    SH_REQUIRE((((uintptr_t)A | (uintptr_t)W) & 15) == 0, "rows_linear: operands must be 16-byte aligned");
    constexpr size_t smem = (size_t)(8 * 256 + kRlStages * 256 * kRlPitch) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        SH_CHECK_CUDA(cudaFuncSetAttribute(rows_linear_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    SH_LAUNCH("gnn_embed_table_linear", st,
              rows_linear_kernel<8><<<ceil_div(rows, 8), 256, smem, st>>>(A, W, rows, D, out, bias, gamma, beta, eps, act, amax_out));
    SH_CHECK_LAUNCH();
    return 0;
}

}  // namespace sh

using namespace sh;

static int launch_pool_fc(const sh_gnn_params *p, const float *partial, int G, int chunks, int n_fixed,
                          const int32_t *mean_div, float *out, cudaStream_t st)
{
    const int D = p->embed_dim;
    SH_LAUNCH("gnn_pool_fc", st,
              pool_fc_kernel<<<dim3(G, G >= 1024 ? 1 : 8), 256, (size_t)D * sizeof(float), st>>>(partial, chunks, D, n_fixed, mean_div, p->fc_w, p->fc_b, out));
    SH_CHECK_LAUNCH();
    return 0;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static int pool_chunks(int n_max) { return n_max >= 512 ? 8 : (n_max >= 128 ? 4 : 1); }

extern "C" size_t sh_gnn_workspace_bytes(int G, int n_max, int D)
{
    const size_t slab = align_up((size_t)G * n_max * D * sizeof(float), 256);
    const size_t part = align_up((size_t)G * pool_chunks(n_max) * D * sizeof(float), 256);
    const size_t pooled = align_up((size_t)G * D * sizeof(float), 256);
    const size_t tc = gnn_tc_supported(D, n_max) ? gnn_tc_workspace_bytes(G, n_max, D, pool_chunks(n_max)) : 0;
    return 2 * slab + part + pooled + tc;
}

extern "C" int sh_dev_gnn_forward(const sh_gnn_params *p, int G, int n_fixed, const int32_t *sizes, const int64_t *ids,
                                  const float *vertex_w, int ld_v, const float *edges, int64_t edge_batch_stride,
                                  int edge_ld, const int32_t *mean_div, float *out, void *workspace,
                                  size_t workspace_bytes, sh_stream_t stream)
{
    SH_REQUIRE(p && G > 0 && n_fixed > 0, "gnn_forward: bad arguments");
    const int D = p->embed_dim;
    SH_REQUIRE(D > 0 && p->num_layers >= 1, "gnn_forward: bad params D=%d layers=%d", D, p->num_layers);
    SH_REQUIRE(workspace_bytes >= sh_gnn_workspace_bytes(G, n_fixed, D), "gnn_forward: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t slab = align_up((size_t)G * n_fixed * D * sizeof(float), 256);
    const int chunks = pool_chunks(n_fixed);
    char *ws = (char *)workspace;
    float *Y = (float *)ws;                       // adjacency product
    float *Hbuf = (float *)(ws + slab);           // linear output, then activations (in place)
    float *partial = (float *)(ws + 2 * slab);
    float *pooled = (float *)(ws + 2 * slab + align_up((size_t)G * chunks * D * sizeof(float), 256));

    const bool tensor_path = gnn_tc_supported(D, n_fixed) && getenv("SCHEMANET_GNN_SIMT") == nullptr;
    if (tensor_path) {
        char *tc_ws = ws + 2 * slab + align_up((size_t)G * chunks * D * sizeof(float), 256) + align_up((size_t)G * D * sizeof(float), 256);
        TcFinal fin{out, mean_div, false};
        if (gnn_forward_tc(p, G, n_fixed, sizes, ids, vertex_w, ld_v, edges, edge_batch_stride, edge_ld, chunks, partial,
                           tc_ws, st, &fin)) return 1;
        if (fin.done) return 0;
    }
    for (int l = 0; l < p->num_layers && !tensor_path; ++l) {
        // Y = ((E + E^T)/2 + I) X          (gnn.py:27-30); layer 0 gathers X from the embedding table (:91)
        GemmArgs a{};
        a.A = edges; a.a_batch = edge_batch_stride; a.lda = edge_ld;
        a.sizes = sizes; a.M = n_fixed; a.K = n_fixed; a.N = D; a.G = G;
        if (l == 0) { a.B = p->embedding; a.b_batch = 0; a.ldb = D; a.b_gather = ids; a.ld_gather = ld_v; }
        else { a.B = Hbuf; a.b_batch = (int64_t)n_fixed * D; a.ldb = D; }
        a.C = Y; a.c_batch = (int64_t)n_fixed * D; a.ldc = D;
        if (launch_sgemm<A_ADJ, B_ROWMAJOR>(a, st, "gnn_adj_gemm")) return 1;
        // Z = Y W^T + b                    (gnn.py:31)
        GemmArgs b{};
        b.A = Y; b.a_batch = (int64_t)n_fixed * D; b.lda = D;
        b.B = p->lin_w[l]; b.b_batch = 0; b.ldb = D; b.bias = p->lin_b[l];
        b.C = Hbuf; b.c_batch = (int64_t)n_fixed * D; b.ldc = D;
        b.sizes = sizes; b.M = n_fixed; b.K = D; b.N = D; b.G = G;
        if (launch_sgemm<A_ROWMAJOR, B_TRANSPOSED>(b, st, "gnn_linear_gemm")) return 1;
        // H = relu(LN(Z))                  (gnn.py:45); last layer: fused vertex-weighted pooling (:94-96)
        const bool last = (l == p->num_layers - 1);
        dim3 grid(chunks, G);
        if (last) {
            const size_t smem = (size_t)8 * D * sizeof(float);
            SH_REQUIRE(smem <= 48 * 1024, "gnn_forward: embed_dim %d too large for the pooling stage", D);
            SH_LAUNCH("ln_relu_kernel", st, ln_relu_kernel<true><<<grid, 256, smem, st>>>(Hbuf, nullptr, G, n_fixed, sizes, n_fixed, D, p->ln_w[l],
                                                         p->ln_b[l], p->ln_eps, vertex_w, ld_v, partial, chunks));
        } else {
            SH_LAUNCH("ln_relu_kernel", st, ln_relu_kernel<false><<<grid, 256, 0, st>>>(Hbuf, Hbuf, G, n_fixed, sizes, n_fixed, D, p->ln_w[l], p->ln_b[l],
                                                        p->ln_eps, nullptr, 0, nullptr, chunks));
        }
        SH_CHECK_LAUNCH();
    }
    return launch_pool_fc(p, partial, G, chunks, n_fixed, mean_div, out, st);
}

extern "C" size_t sh_class_side_workspace_bytes(int K, int Vc, int D) { return sh_gnn_workspace_bytes(K, Vc, D); }

// Stage 3a + the class side of stage 3b in one call: class_vertices / class_edges (the tensors get_atlas() returns) and
// the [K, D] class embeddings.  On the tensor-core path the class graphs are compacted to their un-pruned vertices.
extern "C" int sh_dev_class_side(const sh_gnn_params *p, const float *vertex_weights, float *edge_weights,
                                 const int64_t *class_ingredients, int K, int Vc, float prune_threshold,
                                 int prune_in_place, int remove_self_loop, float *class_vertices, float *class_edges,
                                 float *feat_class, void *workspace, size_t workspace_bytes, sh_stream_t stream)
{
    SH_REQUIRE(p && K > 0 && Vc > 0 && class_vertices && feat_class, "class_side: bad arguments");
    const int D = p->embed_dim;
    SH_REQUIRE(workspace_bytes >= sh_class_side_workspace_bytes(K, Vc, D), "class_side: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const bool tensor_path = gnn_tc_supported(D, Vc) && getenv("SCHEMANET_GNN_SIMT") == nullptr;
    if (!tensor_path || getenv("SCHEMANET_CLASS_UNFUSED") != nullptr) {
        SH_REQUIRE(class_edges != nullptr, "class_side: class_edges may only be omitted on the tensor-core path "
                                           "(embed_dim %% 256 == 0, embed_dim <= 1024, Vc >= 32)");
        if (sh_dev_class_atlas(vertex_weights, edge_weights, K, Vc, prune_threshold, prune_in_place, remove_self_loop,
                               class_vertices, class_edges, stream)) return 1;
        return sh_dev_gnn_forward_class(p, K, Vc, class_vertices, class_edges, class_ingredients, prune_threshold, feat_class,
                                        workspace, workspace_bytes, stream);
    }
    if (launch_class_vertices(vertex_weights, K, Vc, class_vertices, st)) return 1;
    const size_t slab = align_up((size_t)K * Vc * D * sizeof(float), 256);
    const int chunks = pool_chunks(Vc);
    char *ws = (char *)workspace;
    float *partial = (float *)(ws + 2 * slab);
    char *tc_ws = ws + 2 * slab + align_up((size_t)K * chunks * D * sizeof(float), 256) + align_up((size_t)K * D * sizeof(float), 256);
    TcFinal fin{feat_class, nullptr, false};
    if (gnn_class_side_tc(p, edge_weights, K, Vc, prune_threshold, prune_in_place, remove_self_loop, class_vertices, class_edges,
                          class_ingredients, chunks, partial, tc_ws, st, &fin))
        return 1;
    if (fin.done) return 0;
    return launch_pool_fc(p, partial, K, chunks, Vc, nullptr, feat_class, st);
}

// GNN.forward on the K class graphs produced by get_atlas(): like sh_dev_gnn_forward, but knows that vertices with
// class_vertices <= prune_threshold have all-zero edge rows/columns and compacts them away on the tensor-core path.
extern "C" int sh_dev_gnn_forward_class(const sh_gnn_params *p, int K, int Vc, const float *class_vertices,
                                        const float *class_edges, const int64_t *class_ingredients, float prune_threshold,
                                        float *feat_class, void *workspace, size_t workspace_bytes, sh_stream_t stream)
{
    SH_REQUIRE(p && K > 0 && Vc > 0 && class_vertices && class_edges && feat_class, "gnn_forward_class: bad arguments");
    const int D = p->embed_dim;
    SH_REQUIRE(workspace_bytes >= sh_class_side_workspace_bytes(K, Vc, D), "gnn_forward_class: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const bool tensor_path = gnn_tc_supported(D, Vc) && getenv("SCHEMANET_GNN_SIMT") == nullptr;
    if (!tensor_path)
        return sh_dev_gnn_forward(p, K, Vc, nullptr, class_ingredients, class_vertices, Vc, class_edges, (int64_t)Vc * Vc, Vc,
                                  nullptr, feat_class, workspace, workspace_bytes, stream);
    const size_t slab = align_up((size_t)K * Vc * D * sizeof(float), 256);
    const int chunks = pool_chunks(Vc);
    char *ws = (char *)workspace;
    float *partial = (float *)(ws + 2 * slab);
    char *tc_ws = ws + 2 * slab + align_up((size_t)K * chunks * D * sizeof(float), 256) + align_up((size_t)K * D * sizeof(float), 256);
    TcFinal fin{feat_class, nullptr, false};
    if (gnn_class_forward_tc(p, K, Vc, class_vertices, class_edges, class_ingredients, prune_threshold, chunks, partial, tc_ws, st, &fin))
        return 1;
    if (fin.done) return 0;
    return launch_pool_fc(p, partial, K, chunks, Vc, nullptr, feat_class, st);
}

extern "C" int sh_dev_similarity(const float *feat_instance, const float *feat_class, int B, int K, int D, int kind,
                                 float *logits, sh_stream_t stream)
{
    SH_REQUIRE(B > 0 && K > 0 && D > 0 && kind >= 0 && kind <= 2, "similarity: bad arguments");
    if (kind == SH_SIM_INNER_PRODUCT && K <= 1024 && similarity_tc_supported(B, K, D) && getenv("SCHEMANET_GNN_SIMT") == nullptr) {
        // ImageNet scale (1024 x 1000 x 1024): the logits are a real GEMM -> tensor cores (one warp per pair costs 0.48 ms)
        static unsigned *scratch[64] = {nullptr};
        int dev = 0;
        SH_CHECK_CUDA(cudaGetDevice(&dev));
        SH_REQUIRE(dev >= 0 && dev < 64, "similarity: device index out of range");
        if (!scratch[dev]) SH_CHECK_CUDA(cudaMalloc(&scratch[dev], 256));
        return similarity_tc(feat_instance, feat_class, B, K, D, logits, scratch[dev], (cudaStream_t)stream);
    }
    const int64_t pairs = (int64_t)B * K;
    const int grid = (int)min(ceil_div64(pairs, 8), (int64_t)sm_count() * 16);
    SH_LAUNCH("similarity_kernel", (cudaStream_t)stream, similarity_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feat_instance, feat_class, B, K, D, kind, logits));
    SH_CHECK_LAUNCH();
    return 0;
}
