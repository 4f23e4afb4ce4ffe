// discretize_tc.cu -- stage 1 on the 5th-generation tensor cores: TMA-fed tcgen05 GEMM (kind::f16 on fp16 copies of the
// operands by default, kind::tf32 on the fp32 tensors themselves) with the distance + argmin fused into the TMEM epilogue,
// and an exact fp32 re-check of near ties.
//
// Replaces `torch.cdist(seq, vocabulary.weight).argmin(dim=1)` (discretization/discretization.py:65).
//
//   scores  s[r, j] = |c_j|^2 - 2 x_r.c_j            (|x_r|^2 is constant per row: irrelevant for the argmin)
//   coarse  x_r.c_j from tcgen05.mma: 128-byte-swizzled tiles go from HBM to shared memory by TMA, fp32 accumulation in
//           tensor memory
//   epilogue one thread per token row (TMEM lane) and column half: running minimum over all N tiles + the list of every
//           codeword whose coarse score is within `band` of it.  The band is a worst-case bound, not a statistical one
//           (DESIGN.md 4.1): with dx = x - x^ and dc_j = c_j - c^_j the operand rounding residuals (measured by the
//           conversion pass) and g the fp32 accumulation error coefficient of a length-d dot product,
//               |x.c_j - tc(x^, c^_j)| <= |dx| |c_j| + |x^| |dc_j| + g |x^| |c^_j|  =: e      (Cauchy-Schwarz)
//           so the exact-score winner lies within 4e of the coarse minimum (2e per score, two scores).
//   recheck rows with more than one candidate are re-scored in exact fp32 with the reference formula
//           sqrt(max(|x|^2 + |c|^2 - 2 x.c, 0)), lowest index on ties (the clamp and the sqrt create ties)
//
// Warp roles (320 threads, one CTA per SM, persistent over 128-row blocks):
//   warp 0: TMA producer      warp 1: TMEM allocator + MMA issuer (one elected lane)      warps 2-9: epilogue (kSplit = 2 per
//   TMEM lane quarter, each scanning 1/kSplit of the accumulator's columns; merged through shared memory)
// Pipelines: a 4-stage shared-memory ring (full/empty mbarriers, slots freed by tcgen05.commit) and a 2-stage TMEM
// accumulator ring (tmem_full/tmem_empty), so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "discretize.cuh"
#include "tc_common.cuh"

namespace sh {

using namespace tc;

constexpr int TC_BM = 128;       // token rows per tile (UMMA M)
constexpr int TC_BK = 32;        // fp32 per k-block = one 128-byte swizzle row
constexpr int TC_STAGES = 4;
constexpr int TC_EPI_WARPS = 8;    // kSplit warps per TMEM lane quarter: each takes 1/kSplit of a tile's column chunks (16 warps
                                   // measured r02: cfg2 50.0 vs 50.4 us, ImageNet shape 1034 vs 1004 us -- the epilogue is not latency-bound)
constexpr int kSplit = TC_EPI_WARPS / 4;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int kOverflowMark = -1;
constexpr int kListSlots = kSplit == 2 ? 14 : 6;   // per-split shared-memory list capacity (4 stages + the lists fit in 227 KB)
constexpr int kCandStride = kListSlots + 1;   // slot kListSlots absorbs the stores of a full list

struct DiscTcArgs {
    int64_t R;
    int d, M;
    int num_m_blocks, num_n_blocks, num_k_blocks;
    const float *cn;       // [M padded to the N tile] |c_j|^2, +inf beyond M
    const float *xn;       // [R]  |x_r|^2
    const float *xe;       // [R]  |x_r - fp16(x_r)|^2 (half operands) or |x_r - trunc_tf32(x_r)|^2 (tf32 operands)
    const unsigned *cmax_bits;   // bit patterns of max_j |c_j|^2 and ([1]) max_j |c_j - fp16(c_j)|^2
    float gamma;           // accumulation error coefficient: |tc dot - exact dot of the rounded operands| <= gamma |x^| |c^|
    int debug;             // bit 0: epilogue skips its compute, bit 1: MMA issuer skips the MMAs (timing experiments)
    int64_t *out_idx;
    int64_t idx_rows, idx_row_stride, idx_col_stride;
    int *cand_count;       // [R]
    int *cand_idx;         // [R, kCandSlots]
};

template <int BN, int CTAS = 1>
struct DiscTcSmem {
    static constexpr int kABytes = TC_BM * TC_BK * 4;
    static constexpr int kBBytes = (BN / CTAS) * TC_BK * 4;     // a CTA pair stages half of the codebook tile each
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = TC_STAGES * kStageBytes;
    static constexpr int kCandOffset = kBarOffset + 256;
    // kSplit candidate lists per row (one per column group) + the groups' running minima / counts
    static constexpr int kHalfOffset = kCandOffset + kSplit * TC_BM * kCandStride * 8;
    static constexpr int kTotal = kHalfOffset + kSplit * TC_BM * 8 + 1024;   // + slack for 1024-B alignment
};

// kHalf: operands are fp16 copies (64 elements per 128-byte swizzle row, kind::f16, UMMA K = 16) instead of the fp32
// tensors themselves (32 elements per row, kind::tf32, UMMA K = 8).  Same tile bytes and MMA count per k-block; half the
// k-blocks, i.e. half the L2->SM operand traffic, at twice the tensor rate.  fp16 (11 significant bits) rather than bf16
// (8): the rigorous candidate band scales with the operand rounding error, and patch tokens / codewords are far inside
// the fp16 range (values beyond +-65504 saturate; their residual then blows the band up and the row is rescanned exactly).
// CTAS = 2 (opt-in, SCHEMANET_DISC_CTAS=2): CTA pairs (cta_group::2).  One UMMA of M = 256 covers the 128-row blocks of both
// CTAs; each CTA stages its own A rows and HALF of the codebook tile, which cuts the L2->SM operand traffic per flop by a
// third.  Measured on B200 (r01): parity-green but no faster (cfg2 100 vs 94 us per call, B=512 d=768 M=1024 246 vs 236 us,
// ImageNet shape 1272 vs 1291 TFLOP/s) -- the 9 TB/s of operand traffic at cfg2 is not the limiter; the sampled stalls
// are the MMA <-> epilogue hand-offs of 3 K-cycle tiles (two accumulators).  Same protocol as gemm3x_kernel (gnn_tc.cu):
// TMA bytes of both CTAs complete on the leader's `full` barrier, tcgen05.commit multicasts to both CTAs' `empty` /
// `tmem_full` barriers, the peer's epilogue arrives remotely on the leader's `tmem_empty`.
template <int BN, bool kHalf, int CTAS>
__global__ void __launch_bounds__(TC_THREADS, 1)
discretize_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, DiscTcArgs a)
{
    constexpr int KB_ELEMS = kHalf ? 64 : 32;      // elements per k-block (one 128-byte row)
    using S = DiscTcSmem<BN, CTAS>;
    extern __shared__ uint8_t smem_raw[];
    // 1 KB alignment for the 128-byte-swizzled TMA tiles, as an OFFSET into the shared array: going through uintptr_t makes
    // the compiler lose the address space and emit 64-bit generic LD/ST for every shared-memory access of the epilogue
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = (uint64_t *)(smem + S::kBarOffset);
    uint64_t *empty = full + TC_STAGES;
    uint64_t *tmem_full = empty + TC_STAGES;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_ptr = (uint32_t *)(tmem_empty + 2);
    float *cand_s = (float *)(smem + S::kCandOffset);
    int *cand_i = (int *)(cand_s + kSplit * TC_BM * kCandStride);
    float *half_m = (float *)(smem + S::kHalfOffset);
    int *half_c = (int *)(half_m + kSplit * TC_BM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = CTAS == 2 ? (int)cluster_ctarank() : 0;      // 0 = the CTA that issues the MMAs
    const int unit = (int)blockIdx.x / CTAS, num_units = (int)gridDim.x / CTAS;
    const int num_p_blocks = (a.num_m_blocks + CTAS - 1) / CTAS;  // work items: one 128-row block per CTA of the unit

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], TC_EPI_WARPS * CTAS); }
        fence_barrier_init();
    }
    if (CTAS == 2) cluster_sync();       // the peer's barriers must exist before anything signals them
    if (warp == 1) { if (CTAS == 2) tmem_alloc_pair(tmem_ptr, 2 * BN); else tmem_alloc(tmem_ptr, 2 * BN); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int pb = unit; pb < num_p_blocks; pb += num_units) {
                const int mb = pb * CTAS + rank;           // (a row block past the end loads zero rows: TMA OOB fill)
                for (int nb = 0; nb < a.num_n_blocks; ++nb)
                    for (int kb = 0; kb < a.num_k_blocks; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t *sa = smem + stage * S::kStageBytes;
                        if (CTAS == 1) {
                            mbar_arrive_expect_tx(&full[stage], S::kStageBytes);
                            tma_load_2d(sa, &tmA, &full[stage], kb * KB_ELEMS, mb * TC_BM);
                            tma_load_2d(sa + S::kABytes, &tmB, &full[stage], kb * KB_ELEMS, nb * BN);
                        } else {
                            // both CTAs' bytes complete on the LEADER's barrier (the only one the MMA issuer waits on)
                            const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
                            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * S::kStageBytes);
                            tma_load_2d_pair(sa, &tmA, lead_full, kb * KB_ELEMS, mb * TC_BM);
                            tma_load_2d_pair(sa + S::kABytes, &tmB, lead_full, kb * KB_ELEMS, nb * BN + rank * (BN / 2));
                        }
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = kHalf ? make_idesc_f16(TC_BM * CTAS, BN) : make_idesc_tf32(TC_BM * CTAS, BN);
        int stage = 0, as = 0;
        uint32_t phase = 0, aphase = 0;
        for (int pb = unit; pb < num_p_blocks && rank == 0; pb += num_units)
            for (int nb = 0; nb < a.num_n_blocks; ++nb) {
                mbar_wait(&tmem_empty[as], aphase ^ 1);      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < a.num_k_blocks; ++kb) {
                    mbar_wait(&full[stage], phase);          // TMA bytes have landed
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
                        const uint64_t da = make_desc_k_sw128(sa), db = make_desc_k_sw128(sa + S::kABytes);
#pragma unroll
                        for (int k = 0; k < ((a.debug & 2) ? 0 : 4); ++k) {   // one UMMA consumes 32 bytes of K (8 tf32 / 16 fp16) of the 128-byte row
                            const uint64_t ak = da + (uint64_t)(2 * k), bk = db + (uint64_t)(2 * k);
                            if (CTAS == 1) {
                                if (kHalf) umma_f16(tmem_d, ak, bk, idesc, (kb | k) != 0);
                                else umma_tf32(tmem_d, ak, bk, idesc, (kb | k) != 0);
                            } else {
                                if (kHalf) umma_f16_pair(tmem_d, ak, bk, idesc, (kb | k) != 0);
                                else umma_tf32_pair(tmem_d, ak, bk, idesc, (kb | k) != 0);
                            }
                        }
                        if (CTAS == 1) {
                            umma_commit(&empty[stage]);          // slot is free once these MMAs have read it
                            if (kb == a.num_k_blocks - 1) umma_commit(&tmem_full[as]);
                        } else {
                            umma_commit_pair(&empty[stage], 3);  // frees the slot in both CTAs
                            if (kb == a.num_k_blocks - 1) umma_commit_pair(&tmem_full[as], 3);
                        }
                    }
                    __syncwarp();
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
    } else {
        // ===================== epilogue: fused score + running argmin + near-tie candidates =====================
        const int wq = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;              // which group of every tile's column chunks this warp scans (0 .. kSplit-1)
        const int row_in_tile = wq * 32 + lane;
        float *my_s = cand_s + (half * TC_BM + row_in_tile) * kCandStride;
        int *my_i = cand_i + (half * TC_BM + row_in_tile) * kCandStride;
        const uint32_t my_base = smem_u32(my_s), my_end = my_base + 4u * kListSlots;
        constexpr uint32_t kIdxDelta = (uint32_t)kSplit * TC_BM * kCandStride * 4u;   // byte distance cand_s -> cand_i
        (void)my_i;
        constexpr int kChunks = BN / 32;
        constexpr int kChunksPerHalf = kChunks >= kSplit ? kChunks / kSplit : 1;   // (narrow tiles leave the last groups idle)
        // (tf32 operands: the residuals are those of truncation to 10 mantissa bits, an upper bound element by element of what
        // the hardware drops whether it truncates or rounds)
        const float cmax2 = __uint_as_float(a.cmax_bits[0]), dcmax2 = __uint_as_float(a.cmax_bits[1]);
        const float cmax = sqrtf(cmax2), dcmax = sqrtf(dcmax2);
        int as = 0;
        uint32_t aphase = 0;
        for (int pb = unit; pb < num_p_blocks; pb += num_units) {
            const int mb = pb * CTAS + rank;
            const int64_t row = (int64_t)mb * TC_BM + row_in_tile;
            const bool valid = row < a.R;
            // band = 4 e + fp32 slack of the re-check itself (its own rounding, and ties created by sqrt / clamp: two squared
            // distances closer than 2^-21 (|x|^2 + |c|^2) can round to the same fp32 distance).  Non-finite rows give a NaN
            // or infinite band: no candidate or every candidate is kept, and either way the row is rescanned exactly.
            float band = 0.0f;
            if (valid) {
                const float xn2 = a.xn[row], xe2 = a.xe[row];
                const float nx = sqrtf(xn2), dx = sqrtf(xe2), nxh = nx + dx;
                const float e = dx * cmax + nxh * dcmax + a.gamma * nxh * (cmax + dcmax);
                band = 4.0f * e + 1.9073486328125e-6f * (xn2 + cmax2);
            }
            float m_run = INFINITY;
            int cnt = 0;                  // candidates stored; kListSlots means "full: some may have been lost"
            for (int nb = 0; nb < a.num_n_blocks; ++nb) {
                mbar_wait(&tmem_full[as], aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
                for (int c = half * kChunksPerHalf; c < min((half + 1) * kChunksPerHalf, kChunks); ++c) {
                    const int n_base = nb * BN + c * 32;
                    if (n_base >= a.M || (a.debug & 1)) break;
                    float v[32];
                    tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
                    const float4 *cn4 = reinterpret_cast<const float4 *>(a.cn + n_base);
                    float cmin = INFINITY;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 cc = __ldg(cn4 + q);      // +inf beyond M: out-of-range columns never win
                        v[4 * q + 0] = fmaf(-2.0f, v[4 * q + 0], cc.x);
                        v[4 * q + 1] = fmaf(-2.0f, v[4 * q + 1], cc.y);
                        v[4 * q + 2] = fmaf(-2.0f, v[4 * q + 2], cc.z);
                        v[4 * q + 3] = fmaf(-2.0f, v[4 * q + 3], cc.w);
                        cmin = fminf(cmin, fminf(fminf(v[4 * q], v[4 * q + 1]), fminf(v[4 * q + 2], v[4 * q + 3])));
                    }
                    // A chunk matters to a row only if its minimum comes within `band` of the row's running minimum
                    // (~ln(#chunks) times per row, but with 32 rows per warp some lane hits on most chunks, so this
                    // path must stay short: straight-line, branch-free appends from the registers already loaded).
                    if (cmin <= m_run + band) {
                        if (cmin < m_run - band) cnt = 0;          // every older candidate is now out of reach
                        m_run = fminf(m_run, cmin);
                        const float thr = m_run + band;
                        // straight-line predicated appends with explicit shared-space stores (the pointers are carved
                        // from an aligned dynamic-smem base, which the compiler only knows as generic: it emitted 64-bit
                        // generic ST.E, unconditionally, 64 per chunk and thread).  A store is only issued for an
                        // actual candidate; slot kListSlots absorbs the stores of a full list.
                        uint32_t cur = my_base + 4u * (uint32_t)cnt;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const bool hit = v[j] <= thr;
                            if (hit) {
                                asm volatile("st.shared.f32 [%0], %1;" ::"r"(cur), "f"(v[j]) : "memory");
                                asm volatile("st.shared.s32 [%0], %1;" ::"r"(cur + kIdxDelta), "r"(n_base + j) : "memory");
                            }
                            cur = min(cur + (hit ? 4u : 0u), my_end);       // my_end: list is full
                        }
                        cnt = (int)((cur - my_base) >> 2);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CTAS == 1) mbar_arrive(&tmem_empty[as]);
                    else mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0));   // the leader owns the accumulator ring
                }
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
            // merge the two column halves of each row: half 1 publishes its minimum / count, half 0 finalises
            half_m[half * TC_BM + row_in_tile] = m_run;
            half_c[half * TC_BM + row_in_tile] = cnt;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
            if (half == 0 && valid) {
                float m_all = m_run;
#pragma unroll
                for (int h = 1; h < kSplit; ++h) m_all = fminf(m_all, half_m[h * TC_BM + row_in_tile]);
                const float thr = m_all + band;
                bool overflow = false;
                int keep = 0, first = 0;
#pragma unroll 1
                for (int h = 0; h < kSplit; ++h) {
                    const int cnt_h = half_c[h * TC_BM + row_in_tile];
                    const float *ls = cand_s + (h * TC_BM + row_in_tile) * kCandStride;
                    const int *li = cand_i + (h * TC_BM + row_in_tile) * kCandStride;
                    // a full list may have dropped candidates -- harmless if even its minimum is out of reach
                    if (cnt_h >= kListSlots && half_m[h * TC_BM + row_in_tile] <= thr) overflow = true;
                    for (int t = 0; t < min(cnt_h, kListSlots) && !overflow; ++t)
                        if (ls[t] <= thr) {
                            if (keep == 0) first = li[t];
                            if (keep < kCandSlots) a.cand_idx[row * kCandSlots + keep] = li[t];
                            else overflow = true;
                            ++keep;
                        }
                }
                if (overflow) a.cand_count[row] = kOverflowMark;   // exact rescan of the whole row
                else {
                    a.cand_count[row] = keep;
                    if (keep == 1)
                        a.out_idx[(row % a.idx_rows) * a.idx_row_stride + (row / a.idx_rows) * a.idx_col_stride] = first;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");   // lists are reused by the next row block
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CTAS == 2) cluster_sync();       // the peer may still be reading this CTA's shared memory / signalling its barriers
    if (warp == 1) { if (CTAS == 2) tmem_dealloc_pair(tmem_base, 2 * BN); else tmem_dealloc(tmem_base, 2 * BN); }
}

// Exact fp32 re-check of the rows the tensor-core pass could not decide (one warp per row).
__global__ void __launch_bounds__(256)
discretize_recheck_kernel(const float *__restrict__ X, const float *__restrict__ C, const float *__restrict__ cn,
                          const float *__restrict__ xn, int64_t R, int d, int M, const int *__restrict__ cand_count,
                          const int *__restrict__ cand_idx, int64_t *__restrict__ out_idx, int64_t idx_rows,
                          int64_t idx_row_stride, int64_t idx_col_stride, unsigned long long *counters)
{
    const int lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < R; row += (int64_t)gridDim.x * 8) {
        const int cnt = cand_count[row];
        if (cnt == 1) continue;
        const float *x = X + row * d;
        const float xr = xn[row];
        float best_d = INFINITY;
        int best_i = 0x7fffffff;
        const bool full_scan = (cnt == kOverflowMark || cnt <= 0);
        const int total = full_scan ? M : cnt;
        for (int t = 0; t < total; ++t) {
            const int n = full_scan ? t : cand_idx[row * kCandSlots + t];
            const float *c = C + (size_t)n * d;
            float dot = 0.0f;
            for (int k = lane; k < d; k += kWarp) dot = fmaf(x[k], c[k], dot);
            dot = warp_sum(dot);
            const float dist = exact_distance(xr, cn[n], dot);
            if (dist < best_d || (dist == best_d && n < best_i)) { best_d = dist; best_i = n; }
        }
        if (lane == 0) {
            if (best_i == 0x7fffffff) best_i = 0;
            out_idx[(row % idx_rows) * idx_row_stride + (row / idx_rows) * idx_col_stride] = best_i;
            atomicAdd(&counters[0], 1ULL);
            if (full_scan) atomicAdd(&counters[1], 1ULL);
        }
    }
}

__device__ __forceinline__ void warp_row_to_half(const float *__restrict__ row, int d, unsigned short *__restrict__ out, float &norm,
                                                 float &resid);

// |c_j|^2 (+inf padding up to `padded`) and the bit pattern of max_j |c_j|^2; with `ch` also the fp16 copy of the codebook and
// the running maximum of its rounding residuals (what rows_to_half_kernel would produce: one launch less on the stage's chain)
__global__ void __launch_bounds__(256)
codebook_norms_kernel(const float *__restrict__ C, int M, int padded, int d, float *__restrict__ cn, unsigned *cmax_bits,
                      unsigned short *__restrict__ ch, unsigned *max_resid_bits)
{
    const int lane = threadIdx.x & 31;
    for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < padded; r += gridDim.x * 8) {
        if (r >= M) { if (lane == 0) cn[r] = INFINITY; continue; }
        const float *p = C + (size_t)r * d;
        float s = 0.0f;
        for (int k0 = lane; k0 < d; k0 += 8 * kWarp) {      // 8 loads in flight per lane, folded in the order of the plain loop
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (k0 + u * kWarp < d) ? p[k0 + u * kWarp] : 0.0f;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (k0 + u * kWarp < d) s = fmaf(v[u], v[u], s);
        }
        s = warp_sum(s);
        if (lane == 0) { cn[r] = s; atomicMax(cmax_bits, __float_as_uint(s)); }
        if (ch != nullptr) {
            float s2, e;
            warp_row_to_half(p, d, ch + (size_t)r * d, s2, e);            // (the row is in L1 by now)
            if (lane == 0) atomicMax(max_resid_bits, __float_as_uint(fabsf(e)));
        }
    }
}

// One warp converts one fp32 row to fp16 (round to nearest even, saturated to the fp16 range) and returns, on every lane,
// |row|^2 of the ORIGINAL fp32 row and an upper bound of the squared norm of the rounding residual |row - fp16(row)|^2 (the
// measured quantity the candidate band is built from).
__device__ __forceinline__ void warp_row_to_half(const float *__restrict__ row, int d, unsigned short *__restrict__ out, float &norm,
                                                 float &resid)
{
    const int lane = threadIdx.x & 31;
    const float2 *p = reinterpret_cast<const float2 *>(row);
    unsigned *q = reinterpret_cast<unsigned *>(out);
    float s = 0.0f, e = 0.0f;
    // batches of 8 independent 8-byte loads per lane (a whole d = 512 row in flight per warp); the running sums keep
    // the element order of the plain loop, so the norms do not depend on the batching
    for (int k0 = lane; k0 < d / 2; k0 += 8 * kWarp) {
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * kWarp;
            v[u] = (k < d / 2) ? __ldcs(p + k) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * kWarp;
            if (k >= d / 2) break;
            s = fmaf(v[u].x, v[u].x, s);
            s = fmaf(v[u].y, v[u].y, s);
            const __half h0 = __float2half_rn(fminf(fmaxf(v[u].x, -65504.0f), 65504.0f));
            const __half h1 = __float2half_rn(fminf(fmaxf(v[u].y, -65504.0f), 65504.0f));
            const float d0 = v[u].x - __half2float(h0), d1 = v[u].y - __half2float(h1);    // exact in fp32 (NaN / inf stay NaN / inf)
            e = fmaf(d0, d0, e);
            e = fmaf(d1, d1, e);
            q[k] = (unsigned)__half_as_ushort(h0) | ((unsigned)__half_as_ushort(h1) << 16);
        }
    }
    norm = warp_sum(s);
    resid = warp_sum(e) * 1.0000005f;          // the sum itself is rounded: keep the residual norm an upper bound
}

// fp32 rows -> fp16 copy, |row|^2 and the residual norms; max_resid_bits: running maximum of the residuals' bit patterns.
// One warp per row.
__global__ void __launch_bounds__(256)
rows_to_half_kernel(const float *__restrict__ x, int64_t rows, int d, unsigned short *__restrict__ xh, float *__restrict__ norm,
                    float *__restrict__ resid, unsigned *max_resid_bits)
{
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        float s, e;
        warp_row_to_half(x + r * d, d, xh + r * d, s, e);
        if (lane == 0) {
            if (norm) norm[r] = s;
            if (resid) resid[r] = e;
            if (max_resid_bits) atomicMax(max_resid_bits, __float_as_uint(fabsf(e)));   // NaN bit patterns compare above +inf
        }
    }
}

bool discretize_tc_supported(int64_t R, int d, int M, bool half)
{
    // TMA needs 16-byte aligned row strides; tiny problems are not worth a tensor-core launch
    return d % (half ? 8 : 4) == 0 && d >= 32 && M >= 16 && R >= 1 && encode_tiled_fn() != nullptr;
}

int launch_codebook_norms(const float *C, int M, int d, const DiscWorkspace &ws, cudaStream_t st, bool half_copy)
{
    const int padded = (M + 255) / 256 * 256;
    SH_LAUNCH("codebook_norms_kernel", st,
              codebook_norms_kernel<<<ceil_div(padded, 8), 256, 0, st>>>(C, M, padded, d, ws.cn, (unsigned *)(ws.counters + 2),
                                                                         half_copy ? ws.cb : nullptr, (unsigned *)(ws.counters + 2) + 1));
    SH_CHECK_LAUNCH();
    return 0;
}

template <int BN, bool kHalf, int CTAS>
static int launch_tc_n(const CUtensorMap &tmA, const CUtensorMap &tmB, const DiscTcArgs &a, cudaStream_t st)
{
    using S = DiscTcSmem<BN, CTAS>;
    static bool configured = false;
    if (!configured) {
        SH_CHECK_CUDA(cudaFuncSetAttribute(discretize_tc_kernel<BN, kHalf, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
        configured = true;
    }
    const int units = (a.num_m_blocks + CTAS - 1) / CTAS;
    const int num_units = min(units, sm_count() / CTAS);
    const char *name = kHalf ? "discretize_tc_f16_kernel" : "discretize_tc_kernel";
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(num_units * CTAS);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = S::kTotal;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    prof_begin(name, st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, discretize_tc_kernel<BN, kHalf, CTAS>, tmA, tmB, a);
    prof_end(st);
    if (e != cudaSuccess) { set_error("%s launch -> %s", name, cudaGetErrorString(e)); return 1; }
    SH_CHECK_LAUNCH();
    return 0;
}

// SCHEMANET_DISC_CTAS=2 selects CTA pairs (the codebook tile map must then have a box of BN / 2 rows)
static int disc_ctas()
{
    static const int c = [] { const char *e = getenv("SCHEMANET_DISC_CTAS"); return (e && atoi(e) == 2) ? 2 : 1; }();
    return c;
}

template <int BN, bool kHalf>
static int launch_tc(const CUtensorMap &tmA, const CUtensorMap &tmB, const DiscTcArgs &a, cudaStream_t st)
{
    if (disc_ctas() == 2 && BN >= 128) return launch_tc_n<BN, kHalf, 2>(tmA, tmB, a, st);
    return launch_tc_n<BN, kHalf, 1>(tmA, tmB, a, st);
}

// fp16 row-major [rows, cols]: box {64 halves (128 B), box_rows}
static int make_tmap_f16(CUtensorMap *map, const unsigned short *base, uint64_t cols, uint64_t rows, uint32_t box_rows)
{
    EncodeTiledFn fn = encode_tiled_fn();
    SH_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<unsigned short *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SH_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp16) failed with CUresult %d", (int)r);
    return 0;
}

int launch_discretize_tc(const float *X, const float *C, int64_t R, int d, int M, int64_t *out_idx, int64_t idx_rows,
                         int64_t idx_row_stride, int64_t idx_col_stride, const DiscWorkspace &ws, bool half, cudaStream_t st)
{
    SH_REQUIRE(((uintptr_t)X % 16 == 0) && ((uintptr_t)C % 16 == 0), "discretize: tensor-core path needs 16-byte aligned inputs");
    const int BN = M > 128 ? 256 : (M > 64 ? 128 : 64);
    const int b_split = (disc_ctas() == 2 && BN >= 128) ? 2 : 1;   // CTA pairs stage half of the codebook tile each
    CUtensorMap tmA, tmB;
    if (half) {
        // one pass over the tokens produces the fp16 copy, |x|^2 and the residual norms (the fp32 path needs a pass for |x|^2 anyway)
        // (the codebook's fp16 copy and residual maximum come from launch_codebook_norms(..., half_copy = true))
        const int g1 = (int)min(ceil_div64(R, 8), (int64_t)sm_count() * 16);
        SH_LAUNCH("rows_to_half_kernel", st, rows_to_half_kernel<<<g1, 256, 0, st>>>(X, R, d, ws.xb, ws.xn, ws.xe, nullptr));
        SH_CHECK_LAUNCH();
        if (make_tmap_f16(&tmA, ws.xb, (uint64_t)d, (uint64_t)R, TC_BM)) return 1;
        if (make_tmap_f16(&tmB, ws.cb, (uint64_t)d, (uint64_t)M, (uint32_t)(BN / b_split))) return 1;
    } else {
        if (make_tmap_f32(&tmA, X, (uint64_t)d, (uint64_t)R, 1, (uint64_t)d, 0, TC_BM)) return 1;
        if (make_tmap_f32(&tmB, C, (uint64_t)d, (uint64_t)M, 1, (uint64_t)d, 0, (uint32_t)(BN / b_split))) return 1;
        if (launch_row_sqnorm(X, R, d, ws.xn, st, ws.xe, nullptr)) return 1;
        if (launch_row_sqnorm(C, M, d, nullptr, st, nullptr, (unsigned *)(ws.counters + 2) + 1)) return 1;
    }
    DiscTcArgs a{};
    a.R = R; a.d = d; a.M = M;
    a.num_m_blocks = (int)ceil_div64(R, TC_BM);
    a.num_n_blocks = ceil_div(M, BN);
    a.num_k_blocks = ceil_div(d, half ? 64 : TC_BK);
    a.cn = ws.cn; a.xn = ws.xn; a.cmax_bits = (const unsigned *)(ws.counters + 2);
    a.xe = ws.xe;
    // fp32 accumulation of d exact products, in whatever order and with truncating adds (unit roundoff 2^-23): the error
    // is at most (d - 1) 2^-23 sum_k |x^_k c^_k| <= d 2^-23 |x^| |c^| (Higham, Accuracy and Stability, 4.2); doubled as a
    // margin for the alignment shifts inside one UMMA k-step
    a.gamma = 2.0f * (float)d * 1.1920928955078125e-7f;
    a.out_idx = out_idx; a.idx_rows = idx_rows; a.idx_row_stride = idx_row_stride; a.idx_col_stride = idx_col_stride;
    a.cand_count = ws.cand_count; a.cand_idx = ws.cand_idx;
    { const char *e = getenv("SCHEMANET_DISC_DEBUG"); a.debug = e ? atoi(e) : 0; }
    int rc;
    if (half) {
        if (BN == 256) rc = launch_tc<256, true>(tmA, tmB, a, st);
        else if (BN == 128) rc = launch_tc<128, true>(tmA, tmB, a, st);
        else rc = launch_tc<64, true>(tmA, tmB, a, st);
    } else {
        if (BN == 256) rc = launch_tc<256, false>(tmA, tmB, a, st);
        else if (BN == 128) rc = launch_tc<128, false>(tmA, tmB, a, st);
        else rc = launch_tc<64, false>(tmA, tmB, a, st);
    }
    if (rc) return rc;
    const int grid = (int)min(ceil_div64(R, 8), (int64_t)sm_count() * 8);
    SH_LAUNCH("discretize_recheck_kernel", st,
              discretize_recheck_kernel<<<grid, 256, 0, st>>>(X, C, ws.cn, ws.xn, R, d, M, ws.cand_count, ws.cand_idx, out_idx,
                                                              idx_rows, idx_row_stride, idx_col_stride, ws.counters));
    SH_CHECK_LAUNCH();
    return 0;
}

}  // namespace sh
