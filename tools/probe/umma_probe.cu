// umma_probe.cu -- which shared-memory operand layouts does tcgen05.mma kind::f16 accept when the tiles are written by
// ordinary stores (no TMA)?  One CTA, one MMA of M=128, N=64, K=32 (two K=16 steps) per variant; D is read back from
// TMEM and compared with a host reference.
//   variant 0: SWIZZLE_NONE, K-major "interleaved": core matrix = 8 rows x 16 B contiguous; LBO = stride between the two
//              16-byte k-chunks of one K=16 step, SBO = stride between 8-row groups  (cute::UMMA make_umma_desc<Major::K>)
//   variant 1: SWIZZLE_64B, K-major: rows of 64 B (32 halves), 16-byte chunk c of row r stored at c ^ ((r >> 1) & 3)
//   variant 2: A as variant 0; B MN-major "interleaved" (instruction descriptor bit 16 = 1): a 16-byte chunk holds 8
//              consecutive N elements of one k; 8 consecutive k are contiguous (128 B core matrix); SBO = stride between
//              N chunks, LBO = stride between groups of 8 k.  B is given as [K][N] (k rows) in global memory.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int M = 128, N = 64, K = 32;

__device__ __forceinline__ uint32_t idesc_f16(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout)
{
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}

// byte offset of the 16-byte chunk (row r, k-chunk c in 0..3) of a [rows x 32 halves] tile
__device__ __forceinline__ uint32_t chunk_off(int variant, int r, int c)
{
    if (variant == 0) return (uint32_t)((c >> 1) * 0 + 0) + (uint32_t)((r >> 3) * 512 + c * 128 + (r & 7) * 16);   // all 4 chunks of a row group contiguous
    return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) * 16));
}

__global__ void __launch_bounds__(128) probe(const __half *A, const __half *B, const __half *Bt, float *D, int variant)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = smem, *sB = smem + 16384;
    uint64_t *bar = (uint64_t *)(smem + 32768);
    uint32_t *tptr = (uint32_t *)(smem + 32768 + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * 4; i += 128) {
        const int r = i / 4, c = i % 4;
        *reinterpret_cast<uint4 *>(sA + chunk_off(variant, r, c)) = *reinterpret_cast<const uint4 *>(A + r * K + c * 8);
    }
    if (variant < 2) {
        for (int i = tid; i < N * 4; i += 128) {
            const int r = i / 4, c = i % 4;
            *reinterpret_cast<uint4 *>(sB + chunk_off(variant, r, c)) = *reinterpret_cast<const uint4 *>(B + r * K + c * 8);
        }
    } else {
        // Bt [K][N]: chunk (k, n8) = Bt[k][8 n8 .. 8 n8 + 7] -> (k % 8) * 16 + n8 * 128 + (k / 8) * (N / 8) * 128
        for (int i = tid; i < K * (N / 8); i += 128) {
            const int k = i % K, n8 = i / K;
            *reinterpret_cast<uint4 *>(sB + (k % 8) * 16 + n8 * 128 + (k / 8) * (N / 8) * 128) = *reinterpret_cast<const uint4 *>(Bt + k * N + n8 * 8);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tptr;
    if (tid == 0) {
        for (int k = 0; k < 2; ++k) {
            uint64_t da, db;
            if (variant == 0) {       // K step k = chunks 2k, 2k+1: start at +256 B per step, LBO = 128, SBO = 512
                da = desc(smem_u32(sA) + k * 256, 128, 512, 0);
                db = desc(smem_u32(sB) + k * 256, 128, 512, 0);
            } else if (variant == 2) { // B MN-major: a K = 16 step = 2 groups of 8 k, LBO = (N / 8) * 128 apart; SBO = 128
                da = desc(smem_u32(sA) + k * 256, 128, 512, 0);
                db = desc(smem_u32(sB) + k * 2 * (N / 8) * 128, (N / 8) * 128, 128, 0);
            } else if (variant == 3) { // same layout, LBO / SBO exchanged
                da = desc(smem_u32(sA) + k * 256, 128, 512, 0);
                db = desc(smem_u32(sB) + k * 2 * (N / 8) * 128, 128, (N / 8) * 128, 0);
            } else {                  // (variant 1) 64-byte swizzle: +32 B per K step, SBO = 8 rows x 64 B
                da = desc(smem_u32(sA) + k * 32, 16, 512, 4);
                db = desc(smem_u32(sB) + k * 32, 16, 512, 4);
            }
            const uint32_t acc = k != 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc_f16(M, N) | (variant >= 2 ? (1u << 16) : 0u)), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    // wait for the MMAs
    {
        uint32_t ok = 0;
        long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
            if (clock64() - t0 > 2000000000LL) { if (lane == 0) printf("probe: timeout\n"); __trap(); }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N / 32; ++c) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c * 32 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

int main()
{
    __half *hA = (__half *)malloc(M * K * 2), *hB = (__half *)malloc(N * K * 2);
    float *ref = (float *)malloc(M * N * 4), *out = (float *)malloc(M * N * 4);
    srand(1);
    for (int i = 0; i < M * K; ++i) hA[i] = __float2half((float)(rand() % 17 - 8) / 4.0f);
    for (int i = 0; i < N * K; ++i) hB[i] = __float2half((float)(rand() % 13 - 6) / 2.0f);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < K; ++k) s += __half2float(hA[m * K + k]) * __half2float(hB[n * K + k]);
            ref[m * N + n] = s;
        }
    __half *dA, *dB, *dBt;
    float *dD;
    __half *hBt = (__half *)malloc(N * K * 2);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hBt[k * N + n] = hB[n * K + k];
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dBt, N * K * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, N * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dBt, hBt, N * K * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    for (int variant = 0; variant < 4; ++variant) {
        cudaMemset(dD, 0, M * N * 4);
        probe<<<1, 128, 40960>>>(dA, dB, dBt, dD, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out, dD, M * N * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        int bad = 0;
        for (int i = 0; i < M * N; ++i) { double d = fabs(out[i] - ref[i]); if (d > worst) worst = d; if (d > 1e-3) ++bad; }
        printf("variant %d (%s): max |err| = %g, wrong entries = %d / %d\n", variant, variant == 0 ? "SWIZZLE_NONE LBO=128 SBO=512" : (variant == 1 ? "SWIZZLE_64B SBO=512" : (variant == 2 ? "B MN-major interleaved LBO=kgroup SBO=nchunk" : "B MN-major interleaved LBO=nchunk SBO=kgroup")), worst, bad, M * N);
    }
    return 0;
}
