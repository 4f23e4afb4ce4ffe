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
// Design.  153,664 B of attention per image are read exactly once and nothing else is large, but the stage is bound by
// instruction issue and dependent-instruction latency, not by HBM: ~20 thread-instructions per attention element
// (clamp, exp, two normalisations, scatter, mix) against the ~23 an SM can issue per element at HBM speed.
//   * one CTA per image (small batches: per (image, split)), 12 warps, two CTAs per SM;
//   * codes are ranked by a bitonic sort of (code, position) keys, 30 of 36 steps in registers by warp shuffle
//     (sorted-unique order == the reference's std::map iteration order);
//   * a warp owns a range of OUTPUT rows r1 (= distinct codes) and walks their positions p in ascending order; each
//     attention row is read with fully coalesced 128 B warp loads (the next row is requested before the current one
//     is reduced), soft-maxed in registers (warp-shuffle max/sum) and scattered to shared-memory slots indexed by the
//     rank of the column's code.  The fp32 summation order is EXACTLY the reference's (p ascending, q ascending, one
//     scalar accumulator from 0.0f; utils.cpp:9) and there are no atomics on the accumulators or the output;
//   * block means, row normalisation (warp-shuffle row sum), nan_to_num and the 2->1 attribute mix are fused into
//     the epilogue; the output row is written once, coalesced.
#include "common.cuh"

namespace sh {

constexpr int kMaxL = 256;           // tokens per image supported by the shared-memory layout (reference: 196)
constexpr int kGraphThreads = 256;   // 8 warps
constexpr int kGraphWarps = kGraphThreads / kWarp;
constexpr int kMaxLaneCols = kMaxL / kWarp;
static_assert(kGraphThreads == kMaxL, "rank_codes sorts one (code, position) pair per thread");   // columns / output codes owned by one lane (template LC <= this)

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

template <int kWarps>
struct GraphSmemT {
    int64_t code[kMaxL];
    int rank[kMaxL];      // rank of the code at position p among the image's sorted distinct codes
    float cinvm[kMaxL];   // by rank: 1 / cnt (1 when block sums are requested), 0 for ranks >= n -- the epilogue's scale + mask
    int cnt[kMaxL];       // positions per distinct code
    int start[kMaxL + 1]; // CSR offsets into pos[]
    int pos[kMaxL];       // positions grouped by code rank, ascending inside a group
    int loc[kMaxL];       // output index of a distinct code (== rank, or class-local index / -1 for feat_to_e)
    float acls[kMaxL];
    float red0[kMaxL];
    float red1[kMaxL];
    float row[kWarps][2][kMaxL];   // per warp and channel (A, G): the row by position (gather kernels) or by slot (scatter)
    int didx[kMaxL];      // position -> index in the duplicate buffer (-1: first occurrence of its code)
    int multi[kMaxL];     // ranks of the codes that occur more than once
    int krow[kMaxL];      // CSR entry -> position | rank << 8 | first << 16 | last << 17 (instance kernel)
    int nmulti;
    int n;
    int next_row;
    float max0, max1;
};
using GraphSmem = GraphSmemT<kGraphWarps>;   // the 256-thread kernels


// Ranks the codes of image b.  After this call (and the trailing barrier) rank/first/cnt/start/pos/didx/multi/n are valid.
// A bitonic sort of the (code, position) pairs (one pair per thread, 30 of the 36 compare-exchange steps by warp shuffle)
// replaces the O(L^2) rank loops: sorted order == the reference's std::map iteration order, equal codes stay in ascending
// position order, so the sorted array IS the CSR position list.  Requires blockDim.x >= kMaxL.
template <class S>
__device__ __forceinline__ void rank_codes(S &s, const int64_t *codes, int L)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // scratch aliased onto the per-warp row buffers (not in use while ranking)
    // (blocks of up to 512 threads: threads >= kMaxL sort padding among themselves and are ignored afterwards)
    int64_t *xc = reinterpret_cast<int64_t *>((reinterpret_cast<uintptr_t>(&s.row[0][0][0]) + 7) & ~(uintptr_t)7);  // [512] codes
    int *xp = reinterpret_cast<int *>(&s.row[3][0][0]);                            // [512] positions / 32-bit keys
    int *wtot = reinterpret_cast<int *>(&s.row[4][0][0]);                          // [16] heads per warp
    int *wlast = wtot + 16;                                                        // [16] last head index per warp
    int64_t c = (tid < L) ? codes[tid] : INT64_MAX;   // padding sorts behind every real pair (ties broken by position)
    int p = tid;
    if (tid < L) s.code[tid] = c;
    if (tid == 0) { s.next_row = 0; s.nmulti = 0; }
    constexpr int64_t kSmall = (1 << 23) - 1;
    if (!__syncthreads_or(tid < L && (c < 0 || c >= kSmall))) {
        // usual case (codebook indices): one 32-bit key code << 8 | position per thread, compare-exchange = shuffle + min/max
        int key = (tid < L) ? ((int)c << 8) | tid : (int)(kSmall << 8) | tid;
        int *xk = xp;
#pragma unroll
        for (int k = 2; k <= kMaxL; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {     // fully unrolled: the 36 (k, j) pairs are compile-time masks
                int ok;
                if (j >= kWarp) {
                    __syncthreads();
                    xk[tid] = key;
                    __syncthreads();
                    ok = xk[tid ^ j];
                } else {
                    ok = __shfl_xor_sync(kFull, key, j);
                }
                const bool want_min = ((tid & k) == 0) == ((tid & j) == 0);
                key = want_min ? min(key, ok) : max(key, ok);
            }
        }
        c = key >> 8;
        p = key & 255;
    } else {
#pragma unroll 1
        for (int k = 2; k <= kMaxL; k <<= 1) {
#pragma unroll 1
            for (int j = k >> 1; j > 0; j >>= 1) {
                int64_t oc;
                int op;
                if (j >= kWarp) {
                    __syncthreads();
                    xc[tid] = c; xp[tid] = p;
                    __syncthreads();
                    oc = xc[tid ^ j]; op = xp[tid ^ j];
                } else {
                    oc = __shfl_xor_sync(kFull, c, j);
                    op = __shfl_xor_sync(kFull, p, j);
                }
                const bool other_less = oc < c || (oc == c && op < p);
                const bool want_min = ((tid & k) == 0) == ((tid & j) == 0);
                if (want_min == other_less) { c = oc; p = op; }
            }
        }
    }
    __syncthreads();
    xc[tid] = c;
    __syncthreads();
    // run heads (first pair of each distinct code), their ranks by a ballot scan
    const bool live = tid < L;
    const bool head = live && (tid == 0 || xc[tid - 1] != c);
    const unsigned hm = __ballot_sync(kFull, head);
    if (lane == 0) { wtot[warp] = __popc(hm); wlast[warp] = hm ? warp * kWarp + 31 - __clz(hm) : -1; }
    __syncthreads();
    int before = 0, n = 0, run0 = -1;
#pragma unroll
    for (int w = 0; w < kGraphWarps; ++w) {
        const int tw = wtot[w];
        if (w < warp) { before += tw; if (wlast[w] >= 0) run0 = wlast[w]; }
        n += tw;
    }
    const unsigned le = hm & (0xffffffffu >> (31 - lane));
    const int rank = before + __popc(le) - 1;
    if (le) run0 = warp * kWarp + 31 - __clz(le);
    if (tid == 0) s.n = n;
    if (live) {
        const int occ = tid - run0;
        s.rank[p] = rank;
        s.pos[tid] = p;
        s.didx[p] = occ == 0 ? -1 : tid - rank - 1;   // duplicates are numbered in (rank, occurrence) order
        if (head) { s.start[rank] = tid; s.loc[rank] = rank; }
    }
    if (tid == 0) s.start[n] = L;
    __syncthreads();
    if (tid < n) {
        const int len = s.start[tid + 1] - s.start[tid];
        s.cnt[tid] = len;
        if (len > 1) s.multi[atomicAdd(&s.nmulti, 1)] = tid;
    }
    __syncthreads();
}

// One row of (optionally head-averaged) attention logits: lane holds columns q = lane + 32 t; `fill` for q >= L.
template <bool kFromHeads, int LC>
__device__ __forceinline__ void load_row(const GraphArgs &a, int b, int p, int lane, float fill, float (&x)[LC])
{
    const int L = a.L;
    if (kFromHeads) {
        const int T = L + 1;
        const float *base = a.attn + ((size_t)b * a.H * T + (size_t)(p + 1)) * T + 1 + lane;
#pragma unroll
        for (int t = 0; t < LC; ++t) {
            float acc = fill;
            if (lane + kWarp * t < L) {
                acc = 0.0f;
                for (int h = 0; h < a.H; ++h) acc += __ldg(base + (size_t)h * T * T + kWarp * t);
                acc = acc / (float)a.H;   // torch.mean on CPU: sum over heads, then divide
            }
            x[t] = acc;
        }
    } else {
        const float *base = a.attn + ((size_t)b * L + p) * L + lane;
#pragma unroll
        for (int t = 0; t < LC; ++t) x[t] = (lane + kWarp * t < L) ? __ldg(base + kWarp * t) : fill;
    }
}

__device__ __forceinline__ float rcp_approx(float x)   // MUFU.RCP: <= 1 ulp for normal x
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float ex2_approx(float x)   // MUFU.EX2: ~2 ulp
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// masked_fill(x < clamp, -inf) + softmax over the L valid columns held by the warp (schema_net.py:334-336).
// Columns q >= L must hold -inf on entry.  exp(x - m) is evaluated as 2^((x - m) log2 e + 64): the 2^64 bias keeps every
// term a normal number down to x - m = -131 (below the reference's own denormal floor), so nothing is flushed to zero
// that the reference keeps; it cancels in the normalisation, whose final multiply rounds into the denormal range like
// the reference's division does.  An all-masked row gives (-inf) - (-inf) = NaN everywhere, exactly like torch.softmax.
template <int LC>
__device__ __forceinline__ void warp_softmax(float (&x)[LC], float clamp, bool use_clamp)
{
    if (use_clamp) {
#pragma unroll
        for (int t = 0; t < LC; ++t) x[t] = x[t] < clamp ? -INFINITY : x[t];
    }
    float m = x[0];
#pragma unroll
    for (int t = 1; t < LC; ++t) m = fmaxf(m, x[t]);
    m = warp_max(m);
    float sum = 0.0f;
#pragma unroll
    for (int t = 0; t < LC; ++t) {
        x[t] = ex2_approx(fmaf(x[t] - m, 1.4426950408889634f, 64.0f));
        sum += x[t];
    }
    sum = warp_sum(sum);
    // one reciprocal per row instead of an IEEE division per element (ncu: the masked entries, exact zeros, sent every
    // division down the slow path -- 55% of this kernel's instructions); <= 2 ulp from x / sum, inside the 1e-5 bar
    const float inv = rcp_approx(sum);      // sum in [2^64, 2^72): normal
#pragma unroll
    for (int t = 0; t < LC; ++t) x[t] = x[t] * inv;
}

// ---------------------------------------------------------------------------------------------------------------
// vertices of one image (large_scale_feat_to_v.cpp:78-125)
// ---------------------------------------------------------------------------------------------------------------
// The cls-attention row of image b (warp 0 only).  Issued before the codes are ranked so that its DRAM latency is hidden.
template <bool kFromHeads, int LC>
__device__ __forceinline__ void load_cls_row(const GraphArgs &a, int b, float (&x)[LC])
{
    const int lane = threadIdx.x & 31, L = a.L;
    const bool raw = (a.flags & SH_G_RAW_LOGITS) != 0;
    if (kFromHeads) {
        load_row<true, LC>(a, b, -1, lane, -INFINITY, x);   // row "-1" of the sliced map == the cls row (token 0)
    } else {
#pragma unroll
        for (int t = 0; t < LC; ++t) {
            const int q = lane + kWarp * t;
            x[t] = (q < L) ? a.attn_cls[(size_t)b * L + q] : (raw ? -INFINITY : 0.0f);
        }
    }
}

template <bool kFromHeads, int LC, class S>
__device__ __forceinline__ void build_vertices(const GraphArgs &a, S &s, int b, float (&x)[LC])
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = a.L, n = s.n;
    const bool raw = (a.flags & SH_G_RAW_LOGITS) != 0;
    const bool use_clamp = raw && a.clamp_v != SH_NO_CLAMP;
    if (warp == 0) {
        if (raw) {
            if (!kFromHeads && use_clamp && (a.flags & SH_G_WRITE_BACK_CLAMP)) {
#pragma unroll
                for (int t = 0; t < LC; ++t) {
                    const int q = lane + kWarp * t;
                    if (q < L && x[t] < a.clamp_v) a.attn_cls[(size_t)b * L + q] = -INFINITY;   // schema_net.py:296
                }
            }
            warp_softmax<LC>(x, a.clamp_v, use_clamp);
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
    const float fill = raw ? -INFINITY : 0.0f;   // columns q >= L: neutral for the soft-max / for the sums
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
        load_row<kFromHeads, LC>(a, b, s.pos[k_begin], lane, fill, xn);
        for (int k = k_begin; k < k_end; ++k) {
            const int p = s.pos[k];
#pragma unroll
            for (int t = 0; t < LC; ++t) x[t] = xn[t];
            if (k + 1 < k_end) load_row<kFromHeads, LC>(a, b, s.pos[k + 1], lane, fill, xn);   // prefetch the next row
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
                warp_softmax<LC>(x, a.clamp_e, use_clamp);
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
// map.  A warp owns a contiguous range of OUTPUT rows r1 (= of CSR entries: the positions of those codes, ascending) and
// keeps the n partial sums of the current row in a shared-memory buffer indexed by RANK:
//   * the attention row is read with coalesced 128 B warp loads, soft-maxed in registers, and every column is stored
//     straight to its destination: rank[q] when q is the first occurrence of its code, n + (duplicate index) otherwise.
//     The destinations are image constants held in registers as byte offsets, so the scatter is 2 x LC plain STS, no
//     branches; one writer per slot, no atomics;
//   * one lane per repeated code then folds its chain of duplicates into the rank slot, in ascending position order.
// The fp32 order per (r1, r2) is exactly the reference's: rows ascending, columns ascending, one running sum
// (utils.cpp:9).  The epilogue reads all 32 LC slots; those >= n are masked by a zero in the per-lane reciprocal counts.

template <bool kFromHeads, int LC, class S>
__device__ __forceinline__ void build_edges_scatter(const GraphArgs &a, S &s, int b, int split, int nsplit)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = a.L, n = s.n;
    const bool raw = (a.flags & SH_G_RAW_LOGITS) != 0;
    const bool use_clamp = raw && a.clamp_e != SH_NO_CLAMP;
    const bool write_back = !kFromHeads && use_clamp && (a.flags & SH_G_WRITE_BACK_CLAMP);
    const bool mean = (a.flags & SH_G_SUM) == 0;
    const float fill = raw ? -INFINITY : 0.0f;   // columns q >= L: neutral for the soft-max / for the sums
    float *bufA = &s.row[warp][0][0], *bufG = &s.row[warp][1][0];
    const float w0 = __ldg(a.w_e), w1 = __ldg(a.w_e + 1);
    const int nmulti = s.nmulti;
    const int n_store = (a.flags & SH_G_ZERO_PAD) ? L : n;          // columns written per output row

    // image constants per lane: destination slot of column q = lane + 32 t
    int dst[LC];
    unsigned dupmask = 0;   // bit t: column t of this lane repeats an earlier code (its value is parked, not summed)
    unsigned colmask = 0;   // bit t: column lane + 32 t exists (q < L)
#pragma unroll
    for (int t = 0; t < LC; ++t) {
        const int q = lane + kWarp * t;
        int d = kMaxL - 1;      // sink for the columns q >= L (no real slot: n + duplicates = L <= 255 when it is used)
        if (q < L) {
            const int di = s.didx[q];
            d = di < 0 ? s.rank[q] : n + di;
            if (di >= 0) dupmask |= 1u << t;
            colmask |= 1u << t;
        }
        dst[t] = d;
    }
    // one lane per repeated code (lane m -> multi[m]; more than 32 repeated codes fall to the loop below)
    const bool has_m = lane < nmulti;
    const int m_r = has_m ? s.multi[lane] : 0;
    const int m_len = has_m ? s.cnt[m_r] - 1 : 0;
    const int m_base = n + s.start[m_r] - m_r;
    for (int j = L + lane; j < kMaxL; j += kWarp) { bufA[j] = 0.0f; bufG[j] = 0.0f; }   // never written, read by the epilogue

    // rows of this CTA (split of the image), then of this warp; CSR entries [k, k_hi)
    const int r_lo = n * split / nsplit, r_hi = n * (split + 1) / nsplit;
    const int rows = r_hi - r_lo;
    const int nwarps = blockDim.x >> 5;
    int k = s.start[r_lo + rows * warp / nwarps];
    const int k_hi = s.start[r_lo + rows * (warp + 1) / nwarps];
    __syncwarp();

    auto load_geo = [&](int p, float (&g)[LC]) {
        const float *base = a.geo + (size_t)p * L + lane;
#pragma unroll
        for (int t = 0; t < LC; ++t) g[t] = (lane + kWarp * t < L) ? __ldg(base + kWarp * t) : 0.0f;
    };
    // Software pipeline over the CSR entries: the attention and geometry rows of the NEXT entry are requested before the
    // current one is reduced (most codes have a single position: without this every row would expose a DRAM round trip).
    float x[LC], xn[LC], g[LC], gn[LC];
    int info_n = 0;
    if (k < k_hi) {
        info_n = s.krow[k];
        load_row<kFromHeads, LC>(a, b, info_n & 255, lane, fill, xn);
        load_geo(info_n & 255, gn);
    }
    while (k < k_hi) {
        const int info = info_n;
        const int p = info & 255, r1 = (info >> 8) & 255;
#pragma unroll
        for (int t = 0; t < LC; ++t) { x[t] = xn[t]; g[t] = gn[t]; }
        ++k;
        if (k < k_hi) {
            info_n = s.krow[k];
            load_row<kFromHeads, LC>(a, b, info_n & 255, lane, fill, xn);
            load_geo(info_n & 255, gn);
        }
        if (raw) {
            if (write_back) {
                float *wb = a.attn + ((size_t)b * L + p) * L + lane;
#pragma unroll
                for (int t = 0; t < LC; ++t)
                    if (lane + kWarp * t < L && x[t] < a.clamp_e) wb[kWarp * t] = -INFINITY;              // :335
            }
            warp_softmax<LC>(x, a.clamp_e, use_clamp);
        }
        if (info & 0x10000) {          // first position of output row r1: plain stores
#pragma unroll
            for (int t = 0; t < LC; ++t) { bufA[dst[t]] = x[t]; bufG[dst[t]] = g[t]; }
        } else {                        // later positions: running sums continue; duplicates are parked again
#pragma unroll
            for (int t = 0; t < LC; ++t) {
                if (!((colmask >> t) & 1u)) continue;   // the shared sink slot is write-only
                const bool dup = (dupmask >> t) & 1u;
                const float oa = bufA[dst[t]], og = bufG[dst[t]];
                bufA[dst[t]] = dup ? x[t] : oa + x[t];
                bufG[dst[t]] = dup ? g[t] : og + g[t];
            }
        }
        __syncwarp();
        if (has_m) {                    // fold the repeated codes' chains, ascending positions
            float va = bufA[m_r], vg = bufG[m_r];
            for (int kk = 0; kk < m_len; ++kk) { va = va + bufA[m_base + kk]; vg = vg + bufG[m_base + kk]; }
            bufA[m_r] = va;
            bufG[m_r] = vg;
        }
        for (int m = lane + kWarp; m < nmulti; m += kWarp) {
            const int r = s.multi[m];
            const int base = n + s.start[r] - r, len = s.cnt[r] - 1;
            float va = bufA[r], vg = bufG[r];
            for (int kk = 0; kk < len; ++kk) { va = va + bufA[base + kk]; vg = vg + bufG[base + kk]; }
            bufA[r] = va;
            bufG[r] = vg;
        }
        __syncwarp();
        if (!(info & 0x20000)) continue;   // more positions of r1 follow

        // epilogue: block mean, row normalisation, nan_to_num, 2->1 mix.
        // block mean = sum / (cnt1 * cnt2) (utils.cpp:12) as a multiplication by the two precomputed reciprocals: at most
        // 2 ulp from the reference's division (exactly 1.0 for the usual single-occurrence codes)
        const float c1_inv = s.cinvm[r1];
        float ea[LC], eg[LC];
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int t = 0; t < LC; ++t) {
            const float sc = c1_inv * s.cinvm[lane + kWarp * t];   // 0 for the slots >= n (duplicates, padding)
            ea[t] = bufA[lane + kWarp * t] * sc;
            eg[t] = bufG[lane + kWarp * t] * sc;
            s0 += eg[t];
            s1 += ea[t];
        }
        __syncwarp();                   // the buffer is rewritten by the next row
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        float *o = a.edges + (size_t)b * L * L + (size_t)r1 * L + lane;
        // a row whose sums are ordinary positive numbers has only finite entries: one reciprocal per channel (<= 1 ulp from
        // x / s), folded with the mix weights; every other row takes the reference's division + nan_to_num
        // (large_scale_feat_to_e.cpp:135) element by element
        if (s0 > 1.0e-30f && s0 < 1.0e30f && s1 > 1.0e-30f && s1 < 1.0e30f) {
            const float f0 = w0 * rcp_approx(s0), f1 = w1 * rcp_approx(s1);
#pragma unroll
            for (int t = 0; t < LC; ++t)
                if (lane + kWarp * t < n_store) o[kWarp * t] = eg[t] * f0 + ea[t] * f1;               // :140
        } else {
#pragma unroll
            for (int t = 0; t < LC; ++t)
                if (lane + kWarp * t < n_store)
                    o[kWarp * t] = (lane + kWarp * t < n) ? nan_to_num0(eg[t] / s0) * w0 + nan_to_num0(ea[t] / s1) * w1 : 0.0f;
        }
    }
}

// grid: (image, split) pairs; every CTA ranks the codes of its image and builds 1/nsplit of the edge rows (split 0 also
// the vertices), so that B images spread evenly over the 148 SMs.
template <bool kFromHeads, int LC, int kWarps>
__global__ void __launch_bounds__(kWarps * kWarp, kWarps == 8 ? 3 : 2) instance_graph_kernel(GraphArgs a, int nsplit)
{
    __shared__ GraphSmemT<kWarps> s;
    for (int u = blockIdx.x; u < a.B * nsplit; u += gridDim.x) {
        const int b = u / nsplit, split = u % nsplit;
        // the cls row is requested before the (latency-bound) ranking, into warp 0's registers
        float cls[LC];
        const bool do_vertices = a.vertex_w && split == 0;
        if (do_vertices && threadIdx.x < kWarp) load_cls_row<kFromHeads, LC>(a, b, cls);
        rank_codes(s, a.ingredients + (size_t)b * a.L, a.L);
        const int n = s.n;
        if (threadIdx.x == 0 && split == 0) {
            if (a.num_vertices) a.num_vertices[b] = n;
            if (a.max_vertices) atomicMax(a.max_vertices, n);
        }
        if (do_vertices) build_vertices<kFromHeads, LC>(a, s, b, cls);
        if (a.edges) {
            if (threadIdx.x < kMaxL)
                s.cinvm[threadIdx.x] = (int)threadIdx.x < n ? ((a.flags & SH_G_SUM) ? 1.0f : 1.0f / (float)s.cnt[threadIdx.x]) : 0.0f;
            if (threadIdx.x < a.L) {   // CSR entry -> position | rank << 8 | first-of-row << 16 | last-of-row << 17
                const int p = s.pos[threadIdx.x], r = s.rank[p];
                s.krow[threadIdx.x] = p | (r << 8) | (threadIdx.x == s.start[r] ? 0x10000 : 0) |
                                      (threadIdx.x + 1 == s.start[r + 1] ? 0x20000 : 0);
            }
            __syncthreads();
            build_edges_scatter<kFromHeads, LC>(a, s, b, split, nsplit);
            if (a.flags & SH_G_ZERO_PAD) {   // rows n..L-1 of the [L, L] slot
                float *o = a.edges + (size_t)b * a.L * a.L;
                for (int i = n * a.L + split * blockDim.x + threadIdx.x; i < a.L * a.L; i += blockDim.x * nsplit) o[i] = 0.0f;
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
    // (image, split) CTAs.  A CTA's fixed part (rank the codes, vertices, per-lane constants) is ~1/5 of an image, and
    // r01 measurements at B = 256 (profiles/r01_graph_variants.md) have 1 split fastest once every SM has a CTA; smaller
    // batches are split so that no SM idles.
    int nsplit = 1;
    if (edges) {
        static const int forced = [] { const char *e = getenv("SCHEMANET_GRAPH_SPLIT"); return e ? atoi(e) : 0; }();
        nsplit = forced > 0 ? forced : (int)max((int64_t)1, min((int64_t)4, ceil_div64((int64_t)sm_count(), B)));
    }
    const int grid = B * nsplit;
    cudaStream_t st = (cudaStream_t)stream;
    const bool narrow = L <= 7 * kWarp;   // 196 tokens: 7 columns per lane instead of 8
    // 12 warps per image, two CTAs per SM at 80 registers (r01: 8 warps x 3 CTAs is ~4-12 % slower, 16 warps at 64
    // registers spills and is ~25 % slower; profiles/r01_graph_variants.md).  SCHEMANET_GRAPH_WARPS=8 selects the former.
    static const int warps = [] { const char *e = getenv("SCHEMANET_GRAPH_WARPS"); return e ? atoi(e) : 12; }();
#define SH_GRAPH_LAUNCH(HEADS, LCOLS, W)                                                                                \
    SH_LAUNCH("instance_graph_kernel", st, instance_graph_kernel<HEADS, LCOLS, W><<<grid, W * kWarp, 0, st>>>(a, nsplit))
#define SH_GRAPH_LAUNCH_OCC(HEADS, LCOLS)                                                                               \
    do {                                                                                                                \
        if (warps == 8) SH_GRAPH_LAUNCH(HEADS, LCOLS, 8);                                                               \
        else SH_GRAPH_LAUNCH(HEADS, LCOLS, 12);                                                                         \
    } while (0)
    if (heads && narrow) SH_GRAPH_LAUNCH_OCC(true, 7);
    else if (heads) SH_GRAPH_LAUNCH_OCC(true, 8);
    else if (narrow) SH_GRAPH_LAUNCH_OCC(false, 7);
    else SH_GRAPH_LAUNCH_OCC(false, 8);
    SH_CHECK_LAUNCH();
    return 0;
}

// Per-class running sums of the atlas initialisation (scripts/init_schema_net.py:31-34, 57-59):
//     for cls_id, x_b in zip(label, x): acc[cls_id] += x_b;  n_tracked[cls_id] += 1
// One thread per element of a sample (N elements, coalesced), the batch walked IN ORDER, so every accumulator sees the
// reference's sequential fp32 additions (bit-exact) and no atomics are needed.  Labels outside [0, K) are skipped.
namespace sh {
__global__ void __launch_bounds__(256)
class_accumulate_kernel(const float *__restrict__ x, const int64_t *__restrict__ label, int B, int64_t N, int K,
                        float *__restrict__ acc, float *__restrict__ n_tracked)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < N)
        for (int b = 0; b < B; ++b) {
            const int64_t k = label[b];
            if (k >= 0 && k < K) acc[k * N + e] += x[(int64_t)b * N + e];
        }
    if (e == 0 && n_tracked)
        for (int b = 0; b < B; ++b) {
            const int64_t k = label[b];
            if (k >= 0 && k < K) n_tracked[k] += 1.0f;
        }
}
}  // namespace sh

extern "C" int sh_dev_class_accumulate(const float *x, const int64_t *label, int B, int64_t N, int K, float *acc,
                                       float *n_tracked, sh_stream_t stream)
{
    SH_REQUIRE(B > 0 && N > 0 && K > 0, "class_accumulate: bad shape B=%d N=%lld K=%d", B, (long long)N, K);
    const int64_t grid = ceil_div64(N, 256);
    SH_REQUIRE(grid < 2147483647LL, "class_accumulate: sample too large");
    SH_LAUNCH("class_accumulate_kernel", (cudaStream_t)stream,
              sh::class_accumulate_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(x, label, B, N, K, acc, n_tracked));
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
