// common.cuh -- shared helpers of libschemahead (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <float.h>
#include <math.h>

#include "../../include/schemahead.h"

namespace sh {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

// Optional per-kernel timing (sh_profile_enable): CUDA events recorded on the launching stream around each launch.
bool prof_on();   // per-kernel event timing requested (sh_profile_enable): kernels then run one at a time
void prof_begin(const char *name, cudaStream_t st);
void prof_end(cudaStream_t st);
#define SH_LAUNCH(name, stream, ...)          \
    do {                                       \
        sh::prof_begin(name, stream);          \
        __VA_ARGS__;                           \
        sh::prof_end(stream);                  \
    } while (0)

#define SH_CHECK_CUDA(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            sh::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));         \
            return 1;                                                                                    \
        }                                                                                                \
    } while (0)

#define SH_CHECK_LAUNCH()                                                                                \
    do {                                                                                                 \
        cudaError_t _e = cudaGetLastError();                                                             \
        if (_e != cudaSuccess) {                                                                         \
            sh::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e));     \
            return 1;                                                                                    \
        }                                                                                                \
        sh::count_launch();                                                                              \
    } while (0)

#define SH_REQUIRE(cond, ...)                                                                            \
    do {                                                                                                 \
        if (!(cond)) {                                                                                   \
            sh::set_error(__VA_ARGS__);                                                                  \
            return 2;                                                                                    \
        }                                                                                                \
    } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// torch.nan_to_num(x, nan=0): NaN -> 0, +inf -> FLT_MAX, -inf -> -FLT_MAX
__device__ __forceinline__ float nan_to_num0(float x)
{
    if (isnan(x)) return 0.0f;
    if (isinf(x)) return x > 0.0f ? FLT_MAX : -FLT_MAX;
    return x;
}

// torch.max semantics: NaN propagates
__device__ __forceinline__ float max_nan(float a, float b)
{
    return (isnan(a) || isnan(b)) ? NAN : fmaxf(a, b);
}

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

int sm_count();

}  // namespace sh
