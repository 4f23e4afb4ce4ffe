// discretize.cuh -- declarations shared by the exact and the tensor-core discretization paths.
#pragma once
#include "common.cuh"

namespace sh {

// The reference distance for one (token, codeword) pair, evaluated the way ATen's _euclidean_dist does:
// sqrt(clamp_min(|x|^2 + |c|^2 - 2 x.c, 0)).  Comparisons for the argmin are made on THIS value (sqrt and clamp
// create ties; torch.argmin breaks ties towards the lowest index).
__device__ __forceinline__ float exact_distance(float xn, float cn, float dot)
{
    return sqrtf(fmaxf(fmaf(-2.0f, dot, xn + cn), 0.0f));
}

constexpr int kCandSlots = 16;   // near-tie candidates kept per row by the tensor-core path

struct DiscWorkspace {
    unsigned long long *counters;   // [0] rows re-checked, [1] rows whose candidate list overflowed,
                                    // [2] two 32-bit words: bits of max_j |c_j|^2 and of max_j |c_j - fp16(c_j)|^2
    float *cn;                      // [M]  |c_j|^2
    float *xn;                      // [R]  |x_r|^2
    float *xe;                      // [R]  |x_r - fp16(x_r)|^2 (fp16 tensor-core path)
    int *cand_count;                // [R]
    int *cand_idx;                  // [R, kCandSlots]
    unsigned short *xb;             // [R, d]  tokens rounded to fp16 (fp16 tensor-core path)
    unsigned short *cb;             // [M, d]  codebook rounded to fp16
    size_t bytes;
};

inline size_t ws_align(size_t x) { return (x + 255) / 256 * 256; }

inline DiscWorkspace carve_disc_workspace(void *base, int64_t R, int M, int d)
{
    DiscWorkspace w{};
    char *p = (char *)base;
    size_t off = 0;
    w.counters = (unsigned long long *)(p + off); off += 256;
    w.cn = (float *)(p + off); off += ws_align(sizeof(float) * ((size_t)M + 256));   // +inf padded to the N tile
    w.xn = (float *)(p + off); off += ws_align(sizeof(float) * (size_t)R);
    w.xe = (float *)(p + off); off += ws_align(sizeof(float) * (size_t)R);
    w.cand_count = (int *)(p + off); off += ws_align(sizeof(int) * (size_t)R);
    w.cand_idx = (int *)(p + off); off += ws_align(sizeof(int) * (size_t)R * kCandSlots);
    w.xb = (unsigned short *)(p + off); off += ws_align(sizeof(unsigned short) * (size_t)R * d);
    w.cb = (unsigned short *)(p + off); off += ws_align(sizeof(unsigned short) * (size_t)M * d);
    w.bytes = off;
    return w;
}

int launch_row_sqnorm(const float *x, int64_t rows, int d, float *out, cudaStream_t st, float *resid = nullptr,
                      unsigned *max_resid_bits = nullptr);
// |c_j|^2 into ws.cn (+inf padding up to a multiple of 256) and max_j |c_j|^2 into ws.counters[2]
int launch_codebook_norms(const float *C, int M, int d, const DiscWorkspace &ws, cudaStream_t st, bool half_copy = false);
int launch_gather(const float *vocab, const int64_t *idx, int64_t idx_rows, int64_t idx_row_stride,
                  int64_t idx_col_stride, int64_t R, int d, float *out, cudaStream_t st);
int launch_discretize_exact(const float *X, const float *C, const float *cn, int64_t R, int d, int M, int64_t *out_idx,
                            int64_t idx_rows, int64_t idx_row_stride, int64_t idx_col_stride, cudaStream_t st);
// tensor-core path (discretize_tc.cu); returns -1 if the shape is not supported by it
bool discretize_tc_supported(int64_t R, int d, int M, bool half);
int launch_discretize_tc(const float *X, const float *C, int64_t R, int d, int M, int64_t *out_idx, int64_t idx_rows,
                         int64_t idx_row_stride, int64_t idx_col_stride, const DiscWorkspace &ws, bool half, cudaStream_t st);

}  // namespace sh
