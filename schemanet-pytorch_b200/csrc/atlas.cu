// atlas.cu -- stage 3a: class IR-atlas normalisation for sm_100a.
//
// Replaces SchemaNet.get_class_vertices / get_class_edges (schema_inference/graph/schema_net.py:144-175) and
// normalize_sum_clamp (schema_inference/graph/utils.py:25-52).  The reference makes ~12 full passes over the
// [K, Vc, Vc] edge tensor (mask outer product by bmm, masked_fill_, multiply, clamp, sum, divide, nan_to_num);
// here every row is read once into registers, reduced with warp shuffles and written once (HBM-bound:
// 4*Vc B read + 4*Vc B written per row, plus the in-place prune stores the reference also performs).
#include "common.cuh"
#include "atlas_rows.cuh"

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
// Fast path (Vc <= 128*kChunks, Vc % 8 == 0, 16-byte aligned): a warp owns runs of 8 consecutive rows of one class, so the
// keep-mask of its 32 columns per chunk (identical for every row of the class) is built once and held in ONE register
// per chunk group; the next row's loads are issued before the current row is reduced; rows whose sum is an ordinary
// positive number skip the per-element nan_to_num.
template <int kChunks>
__global__ void __launch_bounds__(256)
class_edges_fast_kernel(float *__restrict__ ew, const float *__restrict__ cv, int K, int Vc, float thr, int prune,
                        int prune_in_place, int remove_self_loop, float *__restrict__ ce, float *__restrict__ rowinv)
{
    constexpr int RUN = 8;
    const int units_per_class = (Vc + RUN - 1) / RUN;
    const int64_t units = (int64_t)K * units_per_class;
    const int wpb = blockDim.x >> 5;
    for (int64_t u = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); u < units; u += (int64_t)gridDim.x * wpb)
        class_edges_run<kChunks>(ew, cv, (int)(u / units_per_class), (int)(u % units_per_class) * RUN, Vc, thr, prune,
                                 prune_in_place, remove_self_loop, ce, rowinv);
}

// Generic path: any Vc / alignment, two passes over the row (the second one hits L1/L2).
__global__ void __launch_bounds__(256)
class_edges_kernel(float *__restrict__ ew, const float *__restrict__ cv, int K, int Vc, float thr, int prune,
                   int prune_in_place, int remove_self_loop, float *__restrict__ ce, float *__restrict__ rowinv)
{
    const int lane = threadIdx.x & 31;
    const int64_t rows = (int64_t)K * Vc;
    const int wpb = blockDim.x >> 5;
    for (int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * wpb) {
        const int k = (int)(row / Vc), i = (int)(row % Vc);
        float *src = ew + row * Vc;
        float *dst = ce ? ce + row * Vc : nullptr;
        const float *cvk = cv + (size_t)k * Vc;
        const bool keep_i = !prune || cvk[i] > thr;
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
        if (rowinv && lane == 0) rowinv[row] = inv;
        if (dst == nullptr) continue;
        for (int j = lane; j < Vc; j += kWarp) {
            float x = src[j];   // pruned entries were just zeroed in place or are re-masked here
            if (prune && !(keep_i && cvk[j] > thr)) x = 0.f;
            float o = nan_to_num0(fmaxf(x, 0.f) * inv);
            if (remove_self_loop && i == j) o = 0.f;
            dst[j] = o;
        }
    }
}

}  // namespace sh

using namespace sh;

namespace sh {

int launch_class_vertices(const float *vertex_weights, int K, int Vc, float *class_vertices, cudaStream_t st)
{
    SH_LAUNCH("class_vertices_kernel", st, class_vertices_kernel<<<K, 256, 0, st>>>(vertex_weights, K, Vc, class_vertices));
    SH_CHECK_LAUNCH();
    return 0;
}

// One pass over the edge parameter: in-place prune, class_edges [K, Vc, Vc] (may be null: not materialised) and/or the
// per-row normalisers rowinv [K, Vc] (may be null) from which a consumer rebuilds any class_edges entry exactly.
int launch_class_edges(float *edge_weights, const float *class_vertices, int K, int Vc, float prune_threshold,
                       int prune_in_place, int remove_self_loop, float *class_edges, float *rowinv, cudaStream_t st)
{
    const int prune = prune_threshold >= 0.0f ? 1 : 0;
    const int64_t rows = (int64_t)K * Vc;
    // float4 accesses: every row must start on a 16-byte boundary (Vc % 4 == 0; ImageNet's Vc = 500 qualifies)
    const bool aligned = (Vc % 4 == 0) && ((reinterpret_cast<uintptr_t>(edge_weights) | reinterpret_cast<uintptr_t>(class_edges) |
                                           reinterpret_cast<uintptr_t>(class_vertices)) % 16 == 0);
    if (aligned && Vc <= 1024) {
        const int grid = (int)min(ceil_div64((int64_t)K * ((Vc + 7) / 8), 8), (int64_t)sm_count() * 16);
        if (Vc <= 512)
            SH_LAUNCH("class_edges_kernel", st, class_edges_fast_kernel<4><<<grid, 256, 0, st>>>(edge_weights, class_vertices, K, Vc, prune_threshold, prune,
                                                        prune_in_place, remove_self_loop, class_edges, rowinv));
        else
            SH_LAUNCH("class_edges_kernel", st, class_edges_fast_kernel<8><<<grid, 256, 0, st>>>(edge_weights, class_vertices, K, Vc, prune_threshold, prune,
                                                        prune_in_place, remove_self_loop, class_edges, rowinv));
    } else {
        const int grid = (int)min(ceil_div64(rows, 8), (int64_t)sm_count() * 32);
        SH_LAUNCH("class_edges_kernel", st, class_edges_kernel<<<grid, 256, 0, st>>>(edge_weights, class_vertices, K, Vc, prune_threshold, prune,
                                                    prune_in_place, remove_self_loop, class_edges, rowinv));
    }
    SH_CHECK_LAUNCH();
    return 0;
}

}  // namespace sh

extern "C" int sh_dev_class_atlas(const float *vertex_weights, float *edge_weights, int K, int Vc, float prune_threshold,
                                  int prune_in_place, int remove_self_loop, float *class_vertices, float *class_edges,
                                  sh_stream_t stream)
{
    SH_REQUIRE(K > 0 && Vc > 0, "class_atlas: bad shape K=%d Vc=%d", K, Vc);
    cudaStream_t st = (cudaStream_t)stream;
    if (launch_class_vertices(vertex_weights, K, Vc, class_vertices, st)) return 1;
    if (class_edges == nullptr) return 0;
    return launch_class_edges(edge_weights, class_vertices, K, Vc, prune_threshold, prune_in_place, remove_self_loop,
                              class_edges, nullptr, st);
}
