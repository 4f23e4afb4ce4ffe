// graph_build.cu -- stage 0 (attention prologue) and stage 2 (instance IR-graphs) for sm_100a.
//
// Replaces, per image, the reference's CPU loops
//   ext::feat_to_instance_v  (cpp_extension/src/large_scale_feat_to_v.cpp:41-143)
//   ext::feat_to_instance_e  (cpp_extension/src/large_scale_feat_to_e.cpp:33-150)
//   ext::feat_to_v_attr      (cpp_extension/src/feat_to_v_attr.cpp:74-148)
//   ext::feat_to_e           (cpp_extension/src/feat_to_e.cpp:31-127)
// and the clamp + softmax the Python callers run first (schema_inference/graph/schema_net.py:295-297,334-336),
// optionally also the head-mean/slicing prologue (schema_inference/utils/ingredient_model_wrapper.py:57-69).
//
// Design (HBM-bound stage: 153,664 B of attention per image are read exactly once, nothing else is large):
//   * one CTA per image, 8 warps, several CTAs resident per SM so that >= 32 attention rows are in flight per SM;
//   * codes are ranked in shared memory (sorted-unique order == the reference's std::map iteration order);
//   * a warp owns one OUTPUT row r1 (= one distinct code) at a time and walks that code's positions p in ascending
//     order; each attention row is read with fully coalesced 128 B warp loads, soft-maxed in registers
//     (warp-shuffle max/sum), staged in shared memory, then every lane gathers the columns of the codes r2 it owns
//     and adds them, in ascending column order, into register accumulators.  The fp32 summation order is therefore
//     EXACTLY the reference's (p ascending, q ascending, one scalar accumulator from 0.0f; utils.cpp:9) and there
//     are no atomics on the accumulators or the output;
//   * block means, row normalisation (warp-shuffle row sum), nan_to_num and the 2->1 attribute mix are fused into
//     the epilogue; the output row is written once, coalesced.
#include "common.cuh"

namespace sh {

constexpr int kMaxL = 256;           // tokens per image supported by the shared-memory layout (reference: 196)
constexpr int kGraphThreads = 256;   // 8 warps
constexpr int kGraphWarps = kGraphThreads / kWarp;
constexpr int kMaxLaneCols = kMaxL / kWarp;   // columns / output codes owned by one lane (template LC <= this)

struct GraphArgs {
    const int64_t *ingredients;   // [B, L]
    float *attn;                  // [B, L, L] or extracted [B*H, T, T]
    float *attn_cls;              // [B, L]
    const float *geo;             // [L, L]
    int B, L, H;
    float clamp_v, clamp_e;
    const float *w_v, *w_e;       // [2]
    int flags;
    int64_t *ids;
    float *vertex_w;
    float *edges;
    int32_t *num_vertices;
    int32_t *max_vertices;
    // dense (init-time) variants
    const int64_t *class_ingredients;  // [K, n_max]
    const int64_t *label;              // [B]
    int n_max;
    float *dense_out;                  // [B, n_max, n_max, 2]
};

struct GraphSmem {
    int64_t code[kMaxL];
    int rank[kMaxL];      // rank of the code at position p among the image's sorted distinct codes
    int first[kMaxL];     // 1 if p is the first position holding its code
    int cnt[kMaxL];       // positions per distinct code
    int start[kMaxL + 1]; // CSR offsets into pos[]
    int pos[kMaxL];       // positions grouped by code rank, ascending inside a group
    int loc[kMaxL];       // output index of a distinct code (== rank, or class-local index / -1 for feat_to_e)
    float acls[kMaxL];
    float red0[kMaxL];
    float red1[kMaxL];
    float row[kGraphWarps][4][kMaxL];   // per warp: row by position (A, G) or by rank (A, G) + duplicate-position values (A, G)
    int didx[kMaxL];      // position -> index in the duplicate buffer (-1: first occurrence of its code)
    int multi[kMaxL];     // ranks of the codes that occur more than once
    int nmulti;
    int n;
    int next_row;
    float max0, max1;
};

// Ranks the codes of image b.  After this call (and the trailing barrier) rank/cnt/start/pos/n are valid.
__device__ __forceinline__ void rank_codes(GraphSmem &s, const int64_t *codes, int L)
{
    const int tid = threadIdx.x;
    if (tid < L) s.code[tid] = codes[tid];
    if (tid < kMaxL) s.cnt[tid] = 0;
    if (tid == 0) { s.n = 0; s.next_row = 0; s.nmulti = 0; }
    __syncthreads();
    int occ = 0, first = 1;
    int64_t c = 0;
    if (tid < L) {
        c = s.code[tid];
        for (int q = 0; q < tid; ++q)
            if (s.code[q] == c) { first = 0; ++occ; }
        s.first[tid] = first;
    }
    __syncthreads();
    int rank = 0;
    if (tid < L) {
        for (int q = 0; q < L; ++q) rank += (s.first[q] && s.code[q] < c) ? 1 : 0;
        s.rank[tid] = rank;
        atomicAdd(&s.cnt[rank], 1);
        if (first) atomicAdd(&s.n, 1);
    }
    __syncthreads();
    const int n = s.n;
    for (int r = tid; r <= n; r += blockDim.x) {
        int acc = 0;
        for (int k = 0; k < r; ++k) acc += s.cnt[k];
        s.start[r] = acc;
    }
    __syncthreads();
    if (tid < L) {
        s.pos[s.start[rank] + occ] = tid;
        // duplicates are numbered in (rank, occurrence) order: start[r] - r entries precede rank r's chain
        s.didx[tid] = first ? -1 : s.start[rank] - rank + occ - 1;
    }
    if (tid < n) {
        s.loc[tid] = tid;
        if (s.cnt[tid] > 1) s.multi[atomicAdd(&s.nmulti, 1)] = tid;
    }
    __syncthreads();
}

// One row of (optionally head-averaged) attention logits: lane holds columns q = lane + 32 t.
template <bool kFromHeads, int LC>
__device__ __forceinline__ void load_row(const GraphArgs &a, int b, int p, int lane, float (&x)[LC])
{
    const int L = a.L;
    if (kFromHeads) {
        const int T = L + 1;
        const float *base = a.attn + ((size_t)b * a.H * T + (size_t)(p + 1)) * T + 1;
#pragma unroll
        for (int t = 0; t < LC; ++t) {
            const int q = lane + kWarp * t;
            float acc = 0.0f;
            if (q < L) {
                for (int h = 0; h < a.H; ++h) acc += __ldg(base + (size_t)h * T * T + q);
                acc = acc / (float)a.H;   // torch.mean on CPU: sum over heads, then divide
            }
            x[t] = acc;
        }
    } else {
        const float *base = a.attn + ((size_t)b * L + p) * L;
#pragma unroll
        for (int t = 0; t < LC; ++t) {
            const int q = lane + kWarp * t;
            x[t] = (q < L) ? __ldg(base + q) : 0.0f;
        }
    }
}

// masked_fill(x < clamp, -inf) + softmax over the L valid columns held by the warp (schema_net.py:334-336).
template <int LC>
__device__ __forceinline__ void warp_softmax(float (&x)[LC], int L, int lane, float clamp, bool use_clamp)
{
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < LC; ++t) {
        const int q = lane + kWarp * t;
        if (q < L) {
            if (use_clamp && x[t] < clamp) x[t] = -INFINITY;
            m = fmaxf(m, x[t]);
        }
    }
    m = warp_max(m);
    float sum = 0.0f;
#pragma unroll
    for (int t = 0; t < LC; ++t) {
        const int q = lane + kWarp * t;
        // all-masked row: x - m = (-inf) - (-inf) = NaN, exactly like torch.softmax
        x[t] = (q < L) ? __expf(x[t] - m) : 0.0f;   // ex2.approx path: ~2 ulp, NaN/-inf semantics as expf
        sum += x[t];
    }
    sum = warp_sum(sum);
    // one reciprocal per row instead of 8 IEEE divisions per lane: the masked entries (exact zeros) would send every
    // division down the slow path (ncu: 55% of this kernel's instructions); the result differs from x / sum by at
    // most 1 ulp, far inside the 1e-5 tolerance, and 0 * (1/0) = NaN keeps the all-masked-row semantics
    const float inv = 1.0f / sum;
#pragma unroll
    for (int t = 0; t < LC; ++t) x[t] = x[t] * inv;
}

// ---------------------------------------------------------------------------------------------------------------
// vertices of one image (large_scale_feat_to_v.cpp:78-125)
// ---------------------------------------------------------------------------------------------------------------
template <bool kFromHeads, int LC>
__device__ __forceinline__ void build_vertices(const GraphArgs &a, GraphSmem &s, int b)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = a.L, n = s.n;
    const bool raw = (a.flags & SH_G_RAW_LOGITS) != 0;
    const bool use_clamp = raw && a.clamp_v != SH_NO_CLAMP;
    if (warp == 0) {
        float x[LC];
        if (kFromHeads) {
            GraphArgs a2 = a;   // row "-1" of the sliced map == the cls row (token 0) of the full map
            load_row<true, LC>(a2, b, -1, lane, x);
        } else {
#pragma unroll
            for (int t = 0; t < LC; ++t) {
                const int q = lane + kWarp * t;
                x[t] = (q < L) ? a.attn_cls[(size_t)b * L + q] : 0.0f;
            }
        }
        if (raw) {
            if (!kFromHeads && use_clamp && (a.flags & SH_G_WRITE_BACK_CLAMP)) {
#pragma unroll
                for (int t = 0; t < LC; ++t) {
                    const int q = lane + kWarp * t;
                    if (q < L && x[t] < a.clamp_v) a.attn_cls[(size_t)b * L + q] = -INFINITY;   // schema_net.py:296
                }
            }
            warp_softmax(x, L, lane, a.clamp_v, use_clamp);
#pragma unroll
            for (int t = 0; t < LC; ++t) x[t] = nan_to_num0(x[t]);                        // :297
        }
#pragma unroll
        for (int t = 0; t < LC; ++t) {
            const int q = lane + kWarp * t;
            if (q < L) s.acls[q] = x[t];
        }
    }
    __syncthreads();
    float a0 = 0.0f, a1 = 0.0f;
    if (tid < n) {
        float acc = 0.0f;   // sequential, position order, from 0.0f (utils.cpp:9)
        for (int k = s.start[tid]; k < s.start[tid + 1]; ++k) acc = acc + s.acls[s.pos[k]];
        a0 = (float)s.cnt[tid];
        a1 = (a.flags & SH_G_SUM) ? acc : acc / a0;
        s.red0[tid] = a0;
        s.red1[tid] = a1;
    }
    __syncthreads();
    if (warp == 0) {   // attrs.max(0): NaN propagates like torch.max
        float m0 = -INFINITY, m1 = -INFINITY;
        for (int k = lane; k < n; k += kWarp) { m0 = max_nan(m0, s.red0[k]); m1 = max_nan(m1, s.red1[k]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m0 = max_nan(m0, __shfl_xor_sync(kFull, m0, o));
            m1 = max_nan(m1, __shfl_xor_sync(kFull, m1, o));
        }
        if (lane == 0) { s.max0 = m0; s.max1 = m1; }
    }
    __syncthreads();
    if (tid < n) {
        const float w0 = __ldg(a.w_v), w1 = __ldg(a.w_v + 1);
        const float v0 = nan_to_num0(a0 / s.max0);          // large_scale_feat_to_v.cpp:124
        const float v1 = nan_to_num0(a1 / s.max1);
        a.vertex_w[(size_t)b * L + tid] = v0 * w0 + v1 * w1;   // :125
        a.ids[(size_t)b * L + tid] = s.code[s.pos[s.start[tid]]];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// edges of one image (large_scale_feat_to_e.cpp:99-140); kDense selects the feat_to_e.cpp output convention
// ---------------------------------------------------------------------------------------------------------------
template <bool kFromHeads, bool kDense, int LC>
__device__ __forceinline__ void build_edges(const GraphArgs &a, GraphSmem &s, int b)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = a.L, n = s.n;
    const bool raw = (a.flags & SH_G_RAW_LOGITS) != 0;
    const bool use_clamp = raw && a.clamp_e != SH_NO_CLAMP;
    const bool write_back = !kFromHeads && use_clamp && (a.flags & SH_G_WRITE_BACK_CLAMP);
    const bool mean = (a.flags & SH_G_SUM) == 0;
    float *rowA = s.row[warp][0];
    float *rowG = s.row[warp][1];
    float w0 = 0.f, w1 = 0.f;
    if (!kDense) { w0 = __ldg(a.w_e); w1 = __ldg(a.w_e + 1); }

    // The output codes r2 = lane + 32 t a lane owns are the same for every row of the image: keep each code's first
    // position, its position count and its CSR offset in registers (most codes occur once -> one shared-memory read
    // per (row, code) pair and no loop).
    int q0[LC], qn[LC], qs[LC];
#pragma unroll
    for (int t = 0; t < LC; ++t) {
        const int r2 = lane + kWarp * t;
        const bool live = r2 < n && (!kDense || s.loc[r2] >= 0);
        qs[t] = live ? s.start[r2] : 0;
        qn[t] = live ? s.cnt[r2] : 0;
        q0[t] = live ? s.pos[qs[t]] : 0;
    }
    const int nt = (n + kWarp - 1) / kWarp;   // lane slots in use (warp-uniform)

    for (;;) {
        int r1 = 0;
        if (lane == 0) r1 = atomicAdd(&s.next_row, 1);
        r1 = __shfl_sync(kFull, r1, 0);
        if (r1 >= n) break;
        if (kDense && s.loc[r1] < 0) continue;   // code not in the label's class (feat_to_e.cpp:62-77)

        float acc_a[LC], acc_g[LC];
#pragma unroll
        for (int t = 0; t < LC; ++t) { acc_a[t] = 0.0f; acc_g[t] = 0.0f; }

        const int k_begin = s.start[r1], k_end = s.start[r1 + 1];
        float x[LC], xn[LC];
        load_row<kFromHeads, LC>(a, b, s.pos[k_begin], lane, xn);
        for (int k = k_begin; k < k_end; ++k) {
            const int p = s.pos[k];
#pragma unroll
            for (int t = 0; t < LC; ++t) x[t] = xn[t];
            if (k + 1 < k_end) load_row<kFromHeads, LC>(a, b, s.pos[k + 1], lane, xn);   // prefetch the next row
            float g[LC];
#pragma unroll
            for (int t = 0; t < LC; ++t) {
                const int q = lane + kWarp * t;
                g[t] = (q < L) ? __ldg(a.geo + (size_t)p * L + q) : 0.0f;
            }
            if (raw) {
                if (write_back) {
#pragma unroll
                    for (int t = 0; t < LC; ++t) {
                        const int q = lane + kWarp * t;
                        if (q < L && x[t] < a.clamp_e) a.attn[((size_t)b * L + p) * L + q] = -INFINITY;  // :335
                    }
                }
                warp_softmax<LC>(x, L, lane, a.clamp_e, use_clamp);
            }
            __syncwarp();
#pragma unroll
            for (int t = 0; t < LC; ++t) {
                const int q = lane + kWarp * t;
                if (q < L) { rowA[q] = x[t]; rowG[q] = g[t]; }
            }
            __syncwarp();
            // same fp32 order as the reference: positions ascending, one scalar accumulator per (r1, r2)
#pragma unroll
            for (int t = 0; t < LC; ++t) {
                if (t < nt && qn[t] > 0) {
                    acc_a[t] = acc_a[t] + rowA[q0[t]];
                    acc_g[t] = acc_g[t] + rowG[q0[t]];
                    for (int kk = 1; kk < qn[t]; ++kk) {
                        const int q = s.pos[qs[t] + kk];
                        acc_a[t] = acc_a[t] + rowA[q];
                        acc_g[t] = acc_g[t] + rowG[q];
                    }
                }
            }
        }

        // epilogue: block mean, row normalisation, nan_to_num, 2->1 mix
        const float c1 = (float)(k_end - k_begin);
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int t = 0; t < LC; ++t) {
            if (t < nt && qn[t] > 0) {
                const float denom = c1 * (float)qn[t];          // container.size() (utils.cpp:12)
                if (mean && denom != 1.0f) {                     // x / 1 == x: skip the IEEE division for single pairs
                    acc_g[t] = acc_g[t] / denom;
                    acc_a[t] = acc_a[t] / denom;
                }
                s0 += acc_g[t];
                s1 += acc_a[t];
            }
        }
        if (kDense) {
            const int l1 = s.loc[r1];
            float *o = a.dense_out + (size_t)b * a.n_max * a.n_max * 2;
#pragma unroll
            for (int t = 0; t < LC; ++t) {
                const int r2 = lane + kWarp * t;
                if (t < nt && qn[t] > 0)
                    *reinterpret_cast<float2 *>(o + ((size_t)l1 * a.n_max + s.loc[r2]) * 2) = make_float2(acc_g[t], acc_a[t]);
            }
        } else {
            s0 = warp_sum(s0);
            s1 = warp_sum(s1);
            // one reciprocal per row and channel (<= 1 ulp from x / s); a row whose sums are finite and non-zero has
            // only finite entries, so nan_to_num (large_scale_feat_to_e.cpp:135) is only applied to the other rows
            const float inv0 = 1.0f / s0, inv1 = 1.0f / s1;
            const bool clean = isfinite(inv0) && isfinite(inv1) && isfinite(s0) && isfinite(s1);
            float *o = a.edges + (size_t)b * L * L + (size_t)r1 * L;
#pragma unroll
            for (int t = 0; t < LC; ++t) {
                const int r2 = lane + kWarp * t;
                if (r2 < n) {
                    float v0 = acc_g[t] * inv0, v1 = acc_a[t] * inv1;
                    if (!clean) { v0 = nan_to_num0(v0); v1 = nan_to_num0(v1); }
                    o[r2] = v0 * w0 + v1 * w1;                     // :140
                } else if (r2 < L && (a.flags & SH_G_ZERO_PAD)) {
                    o[r2] = 0.0f;                                  // match.py:54 padding, produced in place
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// instance edges, scatter formulation (the hot path; the gather version above serves the dense init-time API)
// ---------------------------------------------------------------------------------------------------------------
// Most codes of an image occur once, so the [n, n] block-sum matrix is almost a row/column permutation of the attention
// map.  A warp owns one output row r1 and keeps its n partial sums in a shared-memory row indexed by RANK:
//   * a column q that is the first occurrence of its code stores (first row of r1) or adds (later rows) its value
//     straight from registers into slot rank[q] -- one writer per slot, no conflicts, no gather loop;
//   * the few columns that repeat a code park their value in a side buffer, and one lane per repeated code folds its
//     chain into the slot afterwards, in ascending position order.
// The fp32 order per (r1, r2) is again exactly the reference's: rows ascending, columns ascending, one running sum.
template <bool kFromHeads, int LC>
__device__ __forceinline__ void build_edges_scatter(const GraphArgs &a, GraphSmem &s, int b)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = a.L, n = s.n;
    const bool raw = (a.flags & SH_G_RAW_LOGITS) != 0;
    const bool use_clamp = raw && a.clamp_e != SH_NO_CLAMP;
    const bool write_back = !kFromHeads && use_clamp && (a.flags & SH_G_WRITE_BACK_CLAMP);
    const bool mean = (a.flags & SH_G_SUM) == 0;
    float *bufA = s.row[warp][0], *bufG = s.row[warp][1], *dupA = s.row[warp][2], *dupG = s.row[warp][3];
    const float w0 = __ldg(a.w_e), w1 = __ldg(a.w_e + 1);
    const int nmulti = s.nmulti;

    // image-constant per-lane column info: destination slot of column q = lane + 32 t (>= 0: rank slot, < 0: ~dup index)
    int slot[LC];
    float cnt_inv[LC];  // 1 / (positions of the output code r2 = lane + 32 t), for the block mean
#pragma unroll
    for (int t = 0; t < LC; ++t) {
        const int q = lane + kWarp * t;
        slot[t] = (q < L) ? (s.didx[q] < 0 ? s.rank[q] : ~s.didx[q]) : 0x40000000;   // 0x40000000: no column
        cnt_inv[t] = (q < n) ? 1.0f / (float)s.cnt[q] : 1.0f;
    }
    const int nt = (n + kWarp - 1) / kWarp;

    // Software pipeline over (output row, position) pairs: the attention row and the geometry row of the NEXT pair are
    // requested before the current pair is reduced, including across output rows (most codes have a single position,
    // so without this every row would expose a full DRAM round trip).
    auto grab = [&]() {
        int r = 0;
        if (lane == 0) r = atomicAdd(&s.next_row, 1);
        return __shfl_sync(kFull, r, 0);
    };
    auto load_geo = [&](int p, float (&g)[LC]) {
#pragma unroll
        for (int t = 0; t < LC; ++t) {
            const int q = lane + kWarp * t;
            g[t] = (q < L) ? __ldg(a.geo + (size_t)p * L + q) : 0.0f;
        }
    };
    float x[LC], xn[LC], g[LC], gn[LC];
    int r1 = grab();
    if (r1 < n) {
        load_row<kFromHeads, LC>(a, b, s.pos[s.start[r1]], lane, xn);
        load_geo(s.pos[s.start[r1]], gn);
    }
    while (r1 < n) {
        const int r1_next = grab();
        const int k_begin = s.start[r1], k_end = s.start[r1 + 1];
        for (int k = k_begin; k < k_end; ++k) {
            const int p = s.pos[k];
#pragma unroll
            for (int t = 0; t < LC; ++t) { x[t] = xn[t]; g[t] = gn[t]; }
            const int p_next = (k + 1 < k_end) ? s.pos[k + 1] : (r1_next < n ? s.pos[s.start[r1_next]] : -1);
            if (p_next >= 0) {
                load_row<kFromHeads, LC>(a, b, p_next, lane, xn);
                load_geo(p_next, gn);
            }
            if (raw) {
                if (write_back) {
#pragma unroll
                    for (int t = 0; t < LC; ++t) {
                        const int q = lane + kWarp * t;
                        if (q < L && x[t] < a.clamp_e) a.attn[((size_t)b * L + p) * L + q] = -INFINITY;  // :335
                    }
                }
                warp_softmax<LC>(x, L, lane, a.clamp_e, use_clamp);
            }
            __syncwarp();
            if (k == k_begin) {
#pragma unroll
                for (int t = 0; t < LC; ++t) {
                    if (slot[t] >= 0) { if (slot[t] < kMaxL) { bufA[slot[t]] = x[t]; bufG[slot[t]] = g[t]; } }
                    else { dupA[~slot[t]] = x[t]; dupG[~slot[t]] = g[t]; }
                }
            } else {
#pragma unroll
                for (int t = 0; t < LC; ++t) {
                    if (slot[t] >= 0) { if (slot[t] < kMaxL) { bufA[slot[t]] = bufA[slot[t]] + x[t]; bufG[slot[t]] = bufG[slot[t]] + g[t]; } }
                    else { dupA[~slot[t]] = x[t]; dupG[~slot[t]] = g[t]; }
                }
            }
            __syncwarp();
            for (int m = lane; m < nmulti; m += kWarp) {     // fold the repeated codes' chains, ascending positions
                const int r = s.multi[m];
                const int base = s.start[r] - r, len = s.cnt[r] - 1;
                float va = bufA[r], vg = bufG[r];
                for (int kk = 0; kk < len; ++kk) { va = va + dupA[base + kk]; vg = vg + dupG[base + kk]; }
                bufA[r] = va;
                bufG[r] = vg;
            }
        }
        __syncwarp();

        // epilogue: block mean, row normalisation, nan_to_num, 2->1 mix
        // block mean = sum / (cnt1 * cnt2) (utils.cpp:12) as a multiplication by the two precomputed reciprocals: at most
        // 2 ulp from the reference's division (it is exactly 1.0 for the usual single-occurrence codes) and it keeps 14
        // IEEE-division sequences per output row out of an issue-bound kernel
        const float c1_inv = mean ? 1.0f / (float)(k_end - k_begin) : 1.0f;
        float ea[LC], eg[LC];
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int t = 0; t < LC; ++t) {
            const int r2 = lane + kWarp * t;
            ea[t] = 0.0f; eg[t] = 0.0f;
            if (t < nt && r2 < n) {
                const float sc = mean ? c1_inv * cnt_inv[t] : 1.0f;
                ea[t] = bufA[r2] * sc;
                eg[t] = bufG[r2] * sc;
                s0 += eg[t];
                s1 += ea[t];
            }
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        // one reciprocal per row and channel (<= 1 ulp from x / s); a row whose sums are finite and non-zero has only
        // finite entries, so nan_to_num (large_scale_feat_to_e.cpp:135) is only applied to the other rows
        const float inv0 = 1.0f / s0, inv1 = 1.0f / s1;
        const bool clean = isfinite(inv0) && isfinite(inv1) && isfinite(s0) && isfinite(s1);
        float *o = a.edges + (size_t)b * L * L + (size_t)r1 * L;
#pragma unroll
        for (int t = 0; t < LC; ++t) {
            const int r2 = lane + kWarp * t;
            if (r2 < n) {
                float v0 = eg[t] * inv0, v1 = ea[t] * inv1;
                if (!clean) { v0 = nan_to_num0(v0); v1 = nan_to_num0(v1); }
                o[r2] = v0 * w0 + v1 * w1;                     // :140
            } else if (r2 < L && (a.flags & SH_G_ZERO_PAD)) {
                o[r2] = 0.0f;                                  // match.py:54 padding, produced in place
            }
        }
        r1 = r1_next;
    }
}

template <bool kFromHeads, int LC>
__global__ void __launch_bounds__(kGraphThreads) instance_graph_kernel(GraphArgs a)
{
    __shared__ GraphSmem s;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        rank_codes(s, a.ingredients + (size_t)b * a.L, a.L);
        if (threadIdx.x == 0) {
            if (a.num_vertices) a.num_vertices[b] = s.n;
            if (a.max_vertices) atomicMax(a.max_vertices, s.n);
        }
        if (a.vertex_w) build_vertices<kFromHeads, LC>(a, s, b);
        if (a.edges) {
            build_edges_scatter<kFromHeads, LC>(a, s, b);
            if (a.flags & SH_G_ZERO_PAD) {   // rows n..L-1 of the [L, L] slot
                float *o = a.edges + (size_t)b * a.L * a.L;
                for (int i = s.n * a.L + threadIdx.x; i < a.L * a.L; i += blockDim.x) o[i] = 0.0f;
            }
        }
        __syncthreads();
    }
}

// feat_to_e.cpp:31-127 -- only codes of the label's class, written at class-local indices, no normalisation.
template <int LC>
__global__ void __launch_bounds__(kGraphThreads) dense_edges_kernel(GraphArgs a)
{
    __shared__ GraphSmem s;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        rank_codes(s, a.ingredients + (size_t)b * a.L, a.L);
        const int n = s.n;
        if (threadIdx.x < n) {
            const int64_t code = s.code[s.pos[s.start[threadIdx.x]]];
            const int64_t *cls = a.class_ingredients + (size_t)a.label[b] * a.n_max;
            int found = -1;   // a later duplicate key overwrites an earlier one (schema_net.py:124)
            for (int j = 0; j < a.n_max; ++j)
                if (cls[j] == code) found = j;
            s.loc[threadIdx.x] = found;
        }
        __syncthreads();
        build_edges<false, true, LC>(a, s, b);
        __syncthreads();
    }
}

// feat_to_v_attr.cpp:19-63,74-148 -- (count, sum-or-mean attention) scattered at the code id.
__global__ void __launch_bounds__(kGraphThreads)
dense_vertices_kernel(const int64_t *ingredients, const float *attn_cls, int B, int L, int n_vertices, int mean,
                      int ingredients_only, float *out)
{
    __shared__ GraphSmem s;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        float *o = out + (size_t)b * n_vertices * 2;
        for (int i = threadIdx.x; i < n_vertices * 2; i += blockDim.x) o[i] = 0.0f;
        rank_codes(s, ingredients + (size_t)b * L, L);   // contains barriers: the zero fill above is ordered
        const int tid = threadIdx.x;
        if (tid < s.n) {
            const int64_t code = s.code[s.pos[s.start[tid]]];
            float acc = 0.0f;
            if (!ingredients_only) {
                for (int k = s.start[tid]; k < s.start[tid + 1]; ++k) acc = acc + attn_cls[(size_t)b * L + s.pos[k]];
                if (mean) acc = acc / (float)s.cnt[tid];
            }
            if (code >= 0 && code < n_vertices) {
                o[code * 2 + 0] = (float)s.cnt[tid];
                o[code * 2 + 1] = acc;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// stage 0 as a stand-alone kernel (ingredient_model_wrapper.py:57-69): one warp per output row
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attention_prologue_kernel(const float *__restrict__ extracted, int B, int H, int T, float *__restrict__ attn,
                          float *__restrict__ attn_cls)
{
    const int L = T - 1;
    const int64_t rows = (int64_t)B * T;   // (b, p) with p = 0 the cls row
    const int lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
         row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int b = (int)(row / T), p = (int)(row % T);
        const float *src = extracted + ((size_t)b * H * T + p) * T + 1;
        float *dst = (p == 0) ? attn_cls + (size_t)b * L : attn + ((size_t)b * L + (p - 1)) * L;
        for (int q = lane; q < L; q += kWarp) {
            float acc = 0.0f;
            for (int h = 0; h < H; ++h) acc += __ldg(src + (size_t)h * T * T + q);
            dst[q] = acc / (float)H;
        }
    }
}

}  // namespace sh

using namespace sh;

extern "C" int sh_dev_attention_prologue(const float *extracted, int B, int H, int T, float *attn, float *attn_cls,
                                         sh_stream_t stream)
{
    SH_REQUIRE(B > 0 && H > 0 && T > 1, "attention_prologue: bad shape B=%d H=%d T=%d", B, H, T);
    const int64_t rows = (int64_t)B * T;
    const int grid = (int)min((int64_t)sm_count() * 16, ceil_div64(rows, 8));
    SH_LAUNCH("attention_prologue_kernel", (cudaStream_t)stream, attention_prologue_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(extracted, B, H, T, attn, attn_cls));
    SH_CHECK_LAUNCH();
    return 0;
}

extern "C" int sh_dev_instance_graphs(const int64_t *ingredients, float *attn, float *attn_cls, const float *geo_sim,
                                      int B, int L, int H, float clamp_vertex, float clamp_edge, const float *w_vertex,
                                      const float *w_edge, int flags, int64_t *ids, float *vertex_w, float *edges,
                                      int32_t *num_vertices, int32_t *max_vertices, sh_stream_t stream)
{
    SH_REQUIRE(B > 0 && L > 0 && L <= kMaxL, "instance_graphs: need 0 < L <= %d (got B=%d L=%d)", kMaxL, B, L);
    SH_REQUIRE((ids == nullptr) == (vertex_w == nullptr), "instance_graphs: ids and vertex_w go together");
    SH_REQUIRE(!(vertex_w && !w_vertex) && !(edges && !w_edge), "instance_graphs: attribute weights missing");
    SH_REQUIRE(!(edges && !geo_sim), "instance_graphs: geo_sim missing");
    const bool heads = (flags & SH_G_FROM_HEADS) != 0;
    SH_REQUIRE(!heads || (H > 0 && (flags & SH_G_RAW_LOGITS)), "instance_graphs: FROM_HEADS needs H > 0 and RAW_LOGITS");
    GraphArgs a{};
    a.ingredients = ingredients; a.attn = attn; a.attn_cls = attn_cls; a.geo = geo_sim;
    a.B = B; a.L = L; a.H = H; a.clamp_v = clamp_vertex; a.clamp_e = clamp_edge;
    a.w_v = w_vertex; a.w_e = w_edge; a.flags = flags;
    a.ids = ids; a.vertex_w = vertex_w; a.edges = edges; a.num_vertices = num_vertices; a.max_vertices = max_vertices;
    const int grid = B;
    cudaStream_t st = (cudaStream_t)stream;
    const bool narrow = L <= 7 * kWarp;   // 196 tokens: 7 columns per lane instead of 8
    if (heads && narrow) SH_LAUNCH("instance_graph_kernel", st, instance_graph_kernel<true, 7><<<grid, kGraphThreads, 0, st>>>(a));
    else if (heads) SH_LAUNCH("instance_graph_kernel", st, instance_graph_kernel<true, 8><<<grid, kGraphThreads, 0, st>>>(a));
    else if (narrow) SH_LAUNCH("instance_graph_kernel", st, instance_graph_kernel<false, 7><<<grid, kGraphThreads, 0, st>>>(a));
    else SH_LAUNCH("instance_graph_kernel", st, instance_graph_kernel<false, 8><<<grid, kGraphThreads, 0, st>>>(a));
    SH_CHECK_LAUNCH();
    return 0;
}

extern "C" int sh_dev_feat_to_v_attr(const int64_t *ingredients, const float *attn_cls, int B, int L, int n_vertices,
                                     int mean, int ingredients_only, float *out, sh_stream_t stream)
{
    SH_REQUIRE(B > 0 && L > 0 && L <= kMaxL && n_vertices > 0, "feat_to_v_attr: bad shape");
    SH_REQUIRE(ingredients_only || attn_cls, "feat_to_v_attr: attn_cls missing");
    SH_LAUNCH("dense_vertices_kernel", (cudaStream_t)stream, dense_vertices_kernel<<<B, kGraphThreads, 0, (cudaStream_t)stream>>>(ingredients, attn_cls, B, L, n_vertices, mean,
                                                                       ingredients_only, out));
    SH_CHECK_LAUNCH();
    return 0;
}

extern "C" int sh_dev_feat_to_e(const int64_t *ingredients, const float *attn, const float *geo_sim,
                                const int64_t *class_ingredients, const int64_t *label, int B, int L, int K, int n_max,
                                int mean, float *out, sh_stream_t stream)
{
    SH_REQUIRE(B > 0 && L > 0 && L <= kMaxL && K > 0 && n_max > 0, "feat_to_e: bad shape");
    SH_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * n_max * n_max * 2, (cudaStream_t)stream));
    GraphArgs a{};
    a.ingredients = ingredients; a.attn = const_cast<float *>(attn); a.geo = geo_sim;
    a.B = B; a.L = L; a.flags = mean ? 0 : SH_G_SUM; a.clamp_v = a.clamp_e = SH_NO_CLAMP;
    a.class_ingredients = class_ingredients; a.label = label; a.n_max = n_max; a.dense_out = out;
    if (L <= 7 * kWarp) SH_LAUNCH("dense_edges_kernel", (cudaStream_t)stream, dense_edges_kernel<7><<<B, kGraphThreads, 0, (cudaStream_t)stream>>>(a));
    else SH_LAUNCH("dense_edges_kernel", (cudaStream_t)stream, dense_edges_kernel<8><<<B, kGraphThreads, 0, (cudaStream_t)stream>>>(a));
    SH_CHECK_LAUNCH();
    return 0;
}
