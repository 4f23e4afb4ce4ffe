// discretize.cu -- stage 1: nearest visual word of every patch token (exact fp32 CUDA-core path + shared helpers).
//
// Replaces Discretization.encode (discretization/discretization.py:58-70):
//     ingredients = torch.cdist(seq, vocabulary.weight).argmin(dim=1);  seq = vocabulary(ingredients)
// torch.cdist (p=2, more than 25 rows) is ATen's `_euclidean_dist`: sqrt(clamp_min(|x|^2 + |c|^2 - 2 x.c, 0)) with the
// cross term from an fp32 GEMM; argmin returns the LOWEST index among equal distances.  Both the clamp and the sqrt
// can create ties that the squared distances do not have, so the comparison below is made on the sqrt'ed value.
//
// This file holds
//   * the exact path: an fp32 FMA GEMM (128x128 tile, 8x8 per thread) with the distance + argmin fused in the
//     epilogue, so the [R, M] distance matrix the reference materialises (cfg4: 6.4 GB) never exists;
//   * norms / gather helpers shared with the tensor-core path (discretize_tc.cu), which uses the exact scoring
//     function below to re-check its near-tie candidates.
#include "common.cuh"
#include "discretize.cuh"

namespace sh {

// |row|^2 of a row-major [rows, d] matrix, one warp per row; optionally also the squared norm of what truncating every
// element to tf32 (10 mantissa bits) drops -- the operand residual the tensor-core path's candidate band is built from
// (resid per row, and/or the running maximum of its bit pattern over the rows)
__global__ void __launch_bounds__(256) row_sqnorm_kernel(const float *__restrict__ x, int64_t rows, int d,
                                                         float *__restrict__ out, float *__restrict__ resid,
                                                         unsigned *max_resid_bits)
{
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        const float *p = x + r * d;
        float s = 0.0f, e = 0.0f;
        for (int k = lane; k < d; k += kWarp) {
            const float v = p[k];
            s = fmaf(v, v, s);
            const float t = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            e = fmaf(t, t, e);
        }
        s = warp_sum(s);
        e = warp_sum(e) * 1.0000005f;
        if (lane == 0) {
            if (out) out[r] = s;
            if (resid) resid[r] = e;
            if (max_resid_bits) atomicMax(max_resid_bits, __float_as_uint(fabsf(e)));
        }
    }
}

// out_seq[r, :] = vocab[idx[r], :]   (discretization.py:66-67)
__global__ void __launch_bounds__(256)
gather_codewords_kernel(const float *__restrict__ vocab, const int64_t *__restrict__ idx, int64_t idx_rows,
                        int64_t idx_row_stride, int64_t idx_col_stride, int64_t R, int d, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < R; r += (int64_t)gridDim.x * 8) {
        const int64_t code = idx[(r / idx_rows) * idx_col_stride + (r % idx_rows) * idx_row_stride];
        const float *src = vocab + code * d;
        float *dst = out + r * d;
        for (int k = lane; k < d; k += kWarp) dst[k] = src[k];
    }
}

constexpr int DB = 128;   // rows / codewords per tile
constexpr int DK = 8;
constexpr int DT = 8;     // 8x8 outputs per thread, 256 threads

__global__ void __launch_bounds__(256)
discretize_exact_kernel(const float *__restrict__ X, const float *__restrict__ C, const float *__restrict__ cn,
                        int64_t R, int d, int M, int64_t *__restrict__ out_idx, int64_t idx_rows,
                        int64_t idx_row_stride, int64_t idx_col_stride)
{
    __shared__ float Xs[DK][DB + 4];
    __shared__ float Cs[DK][DB + 4];
    __shared__ float xn_s[DB];
    __shared__ float red_d[DB][17];
    __shared__ int red_i[DB][17];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid % 16, ty = tid / 16;
    const int64_t r0 = (int64_t)blockIdx.x * DB;

    // |x|^2 of the tile's rows
    for (int rr = warp; rr < DB; rr += 8) {
        const int64_t r = r0 + rr;
        float s = 0.0f;
        if (r < R)
            for (int k = lane; k < d; k += kWarp) { const float v = X[r * d + k]; s = fmaf(v, v, s); }
        s = warp_sum(s);
        if (lane == 0) xn_s[rr] = s;
    }
    __syncthreads();

    float best_d[DT];
    int best_i[DT];
#pragma unroll
    for (int i = 0; i < DT; ++i) { best_d[i] = INFINITY; best_i[i] = 0x7fffffff; }

    for (int n0 = 0; n0 < M; n0 += DB) {
        float acc[DT][DT];
#pragma unroll
        for (int i = 0; i < DT; ++i)
#pragma unroll
            for (int j = 0; j < DT; ++j) acc[i][j] = 0.0f;
        for (int k0 = 0; k0 < d; k0 += DK) {
            for (int e = tid; e < DB * DK; e += 256) {
                const int mm = e / DK, kk = e % DK;
                const int64_t r = r0 + mm;
                const int k = k0 + kk;
                Xs[kk][mm] = (r < R && k < d) ? X[r * d + k] : 0.0f;
                const int n = n0 + mm;
                Cs[kk][mm] = (n < M && k < d) ? C[(size_t)n * d + k] : 0.0f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < DK; ++kk) {
                float a[DT], b[DT];
#pragma unroll
                for (int i = 0; i < DT; ++i) a[i] = Xs[kk][ty * DT + i];
#pragma unroll
                for (int j = 0; j < DT; ++j) b[j] = Cs[kk][tx * DT + j];
#pragma unroll
                for (int i = 0; i < DT; ++i)
#pragma unroll
                    for (int j = 0; j < DT; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < DT; ++j) {
            const int n = n0 + tx * DT + j;
            if (n < M) {
                const float cnj = cn[n];
#pragma unroll
                for (int i = 0; i < DT; ++i) {
                    const float dist = exact_distance(xn_s[ty * DT + i], cnj, acc[i][j]);
                    if (dist < best_d[i]) { best_d[i] = dist; best_i[i] = n; }   // strict <: lowest index wins
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < DT; ++i) { red_d[ty * DT + i][tx] = best_d[i]; red_i[ty * DT + i][tx] = best_i[i]; }
    __syncthreads();
    if (tid < DB && r0 + tid < R) {
        float bd = red_d[tid][0];
        int bi = red_i[tid][0];
        for (int t = 1; t < 16; ++t) {
            const float dd = red_d[tid][t];
            const int ii = red_i[tid][t];
            if (dd < bd || (dd == bd && ii < bi)) { bd = dd; bi = ii; }
        }
        if (bi == 0x7fffffff) bi = 0;   // every distance was NaN: torch.argmin would return the first NaN's index
        const int64_t r = r0 + tid;
        out_idx[(r / idx_rows) * idx_col_stride + (r % idx_rows) * idx_row_stride] = bi;
    }
}

int launch_row_sqnorm(const float *x, int64_t rows, int d, float *out, cudaStream_t st, float *resid, unsigned *max_resid_bits)
{
    const int grid = (int)min(ceil_div64(rows, 8), (int64_t)sm_count() * 16);
    SH_LAUNCH("row_sqnorm_kernel", st, row_sqnorm_kernel<<<grid, 256, 0, st>>>(x, rows, d, out, resid, max_resid_bits));
    SH_CHECK_LAUNCH();
    return 0;
}

int launch_gather(const float *vocab, const int64_t *idx, int64_t idx_rows, int64_t idx_row_stride,
                  int64_t idx_col_stride, int64_t R, int d, float *out, cudaStream_t st)
{
    const int grid = (int)min(ceil_div64(R, 8), (int64_t)sm_count() * 16);
    SH_LAUNCH("gather_codewords_kernel", st, gather_codewords_kernel<<<grid, 256, 0, st>>>(vocab, idx, idx_rows, idx_row_stride, idx_col_stride, R, d, out));
    SH_CHECK_LAUNCH();
    return 0;
}

int launch_discretize_exact(const float *X, const float *C, const float *cn, int64_t R, int d, int M, int64_t *out_idx,
                            int64_t idx_rows, int64_t idx_row_stride, int64_t idx_col_stride, cudaStream_t st)
{
    const int64_t grid = ceil_div64(R, DB);
    SH_REQUIRE(grid < 2147483647LL, "discretize: too many rows");
    SH_LAUNCH("discretize_exact_kernel", st, discretize_exact_kernel<<<(int)grid, 256, 0, st>>>(X, C, cn, R, d, M, out_idx, idx_rows, idx_row_stride,
                                                       idx_col_stride));
    SH_CHECK_LAUNCH();
    return 0;
}

}  // namespace sh
