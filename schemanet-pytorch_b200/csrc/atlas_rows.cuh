// atlas_rows.cuh -- the row-run body of stage 3a (class_edges_fast_kernel, atlas.cu), shared with the fused
// atlas + adjacency kernel of the tensor-core class side (gnn_tc.cu).
#pragma once
#include "common.cuh"

namespace sh {

// One warp, RUN = 8 consecutive rows [i0, i0 + 8) of class k:
//   keep(i, j) = cv[k,i] > thr && cv[k,j] > thr          (schema_net.py:157-163)
//   x = keep ? ew : 0 ; in place: ew = 0 where !keep      (:164-166)
//   ce = nan_to_num(clamp_min(x, 0) / sum_j clamp_min(x, 0))   (:168)
// The keep-mask of the warp's 32 columns per chunk is built once and held in ONE register per chunk group; the next row's
// loads are issued before the current row is reduced; rows whose sum is an ordinary positive number skip the per-element
// nan_to_num.  ce == nullptr: only the in-place prune and the per-row normalisers rowinv are produced.
template <int kChunks>
__device__ __forceinline__ void class_edges_run(float *__restrict__ ew, const float *__restrict__ cv, int k, int i0, int Vc,
                                                float thr, int prune, int prune_in_place, int remove_self_loop,
                                                float *__restrict__ ce, float *__restrict__ rowinv)
{
    constexpr int RUN = 8;
    const int lane = threadIdx.x & 31;
    const float *cvk = cv + (size_t)k * Vc;
    // bit (4*c + e) of `mask`: column (c*32 + lane)*4 + e survives the prune
    unsigned mask = 0xffffffffu;
    if (prune) {
        mask = 0;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const int j = (c * kWarp + lane) * 4;
            if (j < Vc) {
                const float4 m = *reinterpret_cast<const float4 *>(cvk + j);
                mask |= ((m.x > thr ? 1u : 0u) | (m.y > thr ? 2u : 0u) | (m.z > thr ? 4u : 0u) | (m.w > thr ? 8u : 0u)) << (4 * c);
            }
        }
    }
    float4 cur[kChunks], nxt[kChunks];
    float *src = ew + ((size_t)k * Vc + i0) * Vc;
    float *dst = ce ? ce + ((size_t)k * Vc + i0) * Vc : nullptr;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const int j = (c * kWarp + lane) * 4;
        nxt[c] = (j < Vc) ? *reinterpret_cast<const float4 *>(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int run = min(RUN, Vc - i0);          // (the last run of a class is short when Vc % 8 != 0)
#pragma unroll 1
    for (int r = 0; r < run; ++r) {
        const int i = i0 + r;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) cur[c] = nxt[c];
        if (r + 1 < run) {
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int j = (c * kWarp + lane) * 4;
                if (j < Vc) nxt[c] = *reinterpret_cast<const float4 *>(src + (size_t)(r + 1) * Vc + j);
            }
        }
        const bool keep_i = !prune || cvk[i] > thr;
        const unsigned rowmask = keep_i ? mask : 0u;
        float acc = 0.0f;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const int j = (c * kWarp + lane) * 4;
            if (j < Vc) {
                float4 x = cur[c];
                const unsigned mb = (rowmask >> (4 * c)) & 15u;
                if (mb != 15u) {
                    // in-place prune of the parameter (schema_net.py:164); entries that are already zero (every
                    // call after the first) are not rewritten -- same memory image, half the HBM writes
                    const float4 z = make_float4((mb & 1u) ? x.x : 0.f, (mb & 2u) ? x.y : 0.f, (mb & 4u) ? x.z : 0.f, (mb & 8u) ? x.w : 0.f);
                    if (prune_in_place && (z.x != x.x || z.y != x.y || z.z != x.z || z.w != x.w))
                        *reinterpret_cast<float4 *>(src + (size_t)r * Vc + j) = z;
                    x = z;
                }
                x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
                acc += (x.x + x.y) + (x.z + x.w);
                cur[c] = x;
            }
        }
        acc = warp_sum(acc);
        // one reciprocal per row (<= 1 ulp from x / acc).  acc == 0 means every entry is 0 -> 0/0 = NaN -> 0 in the
        // reference: emit zeros.  Only non-finite sums need the element-wise nan_to_num.
        const bool ordinary = acc < INFINITY && acc >= 0.0f;
        const float inv = (acc == 0.0f) ? 0.0f : 1.0f / acc;
        // rowinv: everything a consumer needs to rebuild this row of class_edges from the (pruned) parameter:
        // ce[i][j] = nan_to_num(max(ew[i][j], 0) * rowinv[i]) -- bit-identical to the values stored below
        if (rowinv && lane == 0) rowinv[(size_t)k * Vc + i] = inv;
        if (dst == nullptr) continue;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const int j = (c * kWarp + lane) * 4;
            if (j < Vc) {
                float4 o = make_float4(cur[c].x * inv, cur[c].y * inv, cur[c].z * inv, cur[c].w * inv);
                if (!ordinary) o = make_float4(nan_to_num0(o.x), nan_to_num0(o.y), nan_to_num0(o.z), nan_to_num0(o.w));
                if (remove_self_loop && i >= j && i < j + 4) {
                    if (i == j) o.x = 0.f; else if (i == j + 1) o.y = 0.f; else if (i == j + 2) o.z = 0.f; else o.w = 0.f;
                }
                __stcs(reinterpret_cast<float4 *>(dst + (size_t)r * Vc + j), o);
            }
        }
    }
}

}  // namespace sh
