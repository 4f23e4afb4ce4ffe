// atlas.cu -- stage 3a: class IR-atlas normalisation for sm_100a.
//
// Replaces SchemaNet.get_class_vertices / get_class_edges (schema_inference/graph/schema_net.py:144-175) and
// normalize_sum_clamp (schema_inference/graph/utils.py:25-52).  The reference makes ~12 full passes over the
// [K, Vc, Vc] edge tensor (mask outer product by bmm, masked_fill_, multiply, clamp, sum, divide, nan_to_num);
// here every row is read once into registers, reduced with warp shuffles and written once (HBM-bound:
// 4*Vc B read + 4*Vc B written per row, plus the in-place prune stores the reference also performs).
#include "common.cuh"

namespace sh {

// class_vertices[k, :] = nan_to_num(clamp_min(vw[k, :], 1e-5) / sum)      (schema_net.py:144-150)
__global__ void __launch_bounds__(256)
class_vertices_kernel(const float *__restrict__ vw, int K, int Vc, float *__restrict__ cv)
{
    __shared__ float red[8];
    __shared__ float total;
    const int k = blockIdx.x;
    const float *src = vw + (size_t)k * Vc;
    float acc = 0.0f;
    for (int i = threadIdx.x; i < Vc; i += blockDim.x) acc += fmaxf(src[i], 1.0e-5f);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        total = t;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Vc; i += blockDim.x)
        cv[(size_t)k * Vc + i] = nan_to_num0(fmaxf(src[i], 1.0e-5f) / total);
}

// One warp per row (k, i) of the edge tensor.
//   keep(i, j) = cv[k,i] > thr && cv[k,j] > thr          (:157-163, the bmm of the 0/1 vertex mask with itself)
//   x = keep ? ew : 0 ; in place: ew = 0 where !keep      (:164-166)
//   ce = nan_to_num(clamp_min(x, 0) / sum_j clamp_min(x, 0))   (:168)
template <int kChunks>   // kChunks > 0: row cached in registers (Vc <= 128*kChunks, Vc % 4 == 0); 0: generic two-pass
__global__ void __launch_bounds__(256)
class_edges_kernel(float *__restrict__ ew, const float *__restrict__ cv, int K, int Vc, float thr, int prune,
                   int prune_in_place, int remove_self_loop, float *__restrict__ ce)
{
    const int lane = threadIdx.x & 31;
    const int64_t rows = (int64_t)K * Vc;
    const int wpb = blockDim.x >> 5;
    for (int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * wpb) {
        const int k = (int)(row / Vc), i = (int)(row % Vc);
        float *src = ew + row * Vc;
        float *dst = ce + row * Vc;
        const float *cvk = cv + (size_t)k * Vc;
        const bool keep_i = !prune || cvk[i] > thr;
        if constexpr (kChunks > 0) {
            float4 v[kChunks > 0 ? kChunks : 1];
            float acc = 0.0f;
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int j = (c * kWarp + lane) * 4;
                v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < Vc) {
                    float4 x = *reinterpret_cast<const float4 *>(src + j);
                    if (prune) {
                        const float4 m = *reinterpret_cast<const float4 *>(cvk + j);
                        const bool k0 = keep_i && m.x > thr, k1 = keep_i && m.y > thr, k2 = keep_i && m.z > thr,
                                   k3 = keep_i && m.w > thr;
                        // in-place prune of the parameter (schema_net.py:164); entries that are already zero (every
                        // call after the first) are not rewritten -- same memory image, half the HBM writes
                        const bool dirty = (!k0 && x.x != 0.f) || (!k1 && x.y != 0.f) || (!k2 && x.z != 0.f) || (!k3 && x.w != 0.f);
                        if (prune_in_place && dirty) {
                            float4 z = make_float4(k0 ? x.x : 0.f, k1 ? x.y : 0.f, k2 ? x.z : 0.f, k3 ? x.w : 0.f);
                            *reinterpret_cast<float4 *>(src + j) = z;
                        }
                        x.x = k0 ? x.x : 0.f; x.y = k1 ? x.y : 0.f; x.z = k2 ? x.z : 0.f; x.w = k3 ? x.w : 0.f;
                    }
                    x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
                    acc += (x.x + x.y) + (x.z + x.w);
                    v[c] = x;
                }
            }
            acc = warp_sum(acc);
            // one reciprocal per row: pruned entries are exact zeros and 0 / acc would take the IEEE-division slow path
            // for most of the tensor; x * (1 / acc) differs from x / acc by at most 1 ulp and keeps 0 * inf = NaN -> 0
            const float inv = 1.0f / acc;
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int j = (c * kWarp + lane) * 4;
                if (j < Vc) {
                    float4 o = make_float4(nan_to_num0(v[c].x * inv), nan_to_num0(v[c].y * inv),
                                           nan_to_num0(v[c].z * inv), nan_to_num0(v[c].w * inv));
                    if (remove_self_loop && i >= j && i < j + 4) {
                        if (i == j) o.x = 0.f; else if (i == j + 1) o.y = 0.f; else if (i == j + 2) o.z = 0.f; else o.w = 0.f;
                    }
                    __stcs(reinterpret_cast<float4 *>(dst + j), o);
                }
            }
        } else {
            float acc = 0.0f;
            for (int j = lane; j < Vc; j += kWarp) {
                float x = src[j];
                if (prune) {
                    const bool keep = keep_i && cvk[j] > thr;
                    if (!keep) { if (prune_in_place && x != 0.f) src[j] = 0.f; x = 0.f; }
                }
                acc += fmaxf(x, 0.f);
            }
            acc = warp_sum(acc);
            const float inv = 1.0f / acc;
            for (int j = lane; j < Vc; j += kWarp) {
                float x = src[j];   // pruned entries were just zeroed in place or are re-masked here
                if (prune && !(keep_i && cvk[j] > thr)) x = 0.f;
                float o = nan_to_num0(fmaxf(x, 0.f) * inv);
                if (remove_self_loop && i == j) o = 0.f;
                dst[j] = o;
            }
        }
    }
}

}  // namespace sh

using namespace sh;

extern "C" int sh_dev_class_atlas(const float *vertex_weights, float *edge_weights, int K, int Vc, float prune_threshold,
                                  int prune_in_place, int remove_self_loop, float *class_vertices, float *class_edges,
                                  sh_stream_t stream)
{
    SH_REQUIRE(K > 0 && Vc > 0, "class_atlas: bad shape K=%d Vc=%d", K, Vc);
    cudaStream_t st = (cudaStream_t)stream;
    SH_LAUNCH("class_vertices_kernel", st, class_vertices_kernel<<<K, 256, 0, st>>>(vertex_weights, K, Vc, class_vertices));
    SH_CHECK_LAUNCH();
    if (class_edges == nullptr) return 0;
    const int prune = prune_threshold >= 0.0f ? 1 : 0;
    const int64_t rows = (int64_t)K * Vc;
    const int grid = (int)min(ceil_div64(rows, 8), (int64_t)sm_count() * 32);
    const bool aligned = (Vc % 4 == 0) && ((reinterpret_cast<uintptr_t>(edge_weights) | reinterpret_cast<uintptr_t>(class_edges) |
                                           reinterpret_cast<uintptr_t>(class_vertices)) % 16 == 0);
    if (aligned && Vc <= 512)
        SH_LAUNCH("class_edges_kernel", st, class_edges_kernel<4><<<grid, 256, 0, st>>>(edge_weights, class_vertices, K, Vc, prune_threshold, prune,
                                                    prune_in_place, remove_self_loop, class_edges));
    else if (aligned && Vc <= 1024)
        SH_LAUNCH("class_edges_kernel", st, class_edges_kernel<8><<<grid, 256, 0, st>>>(edge_weights, class_vertices, K, Vc, prune_threshold, prune,
                                                    prune_in_place, remove_self_loop, class_edges));
    else
        SH_LAUNCH("class_edges_kernel", st, class_edges_kernel<0><<<grid, 256, 0, st>>>(edge_weights, class_vertices, K, Vc, prune_threshold, prune,
                                                    prune_in_place, remove_self_loop, class_edges));
    SH_CHECK_LAUNCH();
    return 0;
}
