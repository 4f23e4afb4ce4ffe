// capi.cu -- C-ABI plumbing of libschemahead: error state, launch accounting, the stage-1 dispatcher and the
// host-buffer entry points that mirror the reference's CPU-tensor pybind functions
// (cpp_extension/src/extension.cpp:6-12).
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "discretize.cuh"

namespace sh {

static thread_local char g_error[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional per-kernel profiling -------------------------------------------------------------------------
struct ProfRecord { const char *name; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRecord> g_prof;
static std::mutex g_prof_mu;

bool prof_on() { return g_prof_on; }

void prof_begin(const char *name, cudaStream_t st)
{
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRecord r{name, nullptr, nullptr};
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
}

void prof_end(cudaStream_t st)
{
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof.empty()) cudaEventRecord(g_prof.back().b, st);
}

int sm_count()
{
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}

// Scratch device buffers of the host entry points (grown on demand, reused across calls, one set per process).
struct HostScratch {
    std::mutex mu;
    std::vector<void *> bufs;
    std::vector<size_t> sizes;
    cudaStream_t stream = nullptr;
    void *get(size_t slot, size_t bytes)
    {
        if (bufs.size() <= slot) { bufs.resize(slot + 1, nullptr); sizes.resize(slot + 1, 0); }
        if (sizes[slot] < bytes) {
            if (bufs[slot]) cudaFree(bufs[slot]);
            bufs[slot] = nullptr;
            if (cudaMalloc(&bufs[slot], bytes) != cudaSuccess) { sizes[slot] = 0; return nullptr; }
            sizes[slot] = bytes;
        }
        return bufs[slot];
    }
};
static HostScratch g_scratch;

}  // namespace sh

using namespace sh;

extern "C" int sh_abi_version(void) { return SH_ABI_VERSION; }
extern "C" const char *sh_last_error(void) { return g_error; }
extern "C" int64_t sh_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int sh_profile_enable(int on)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
    return 0;
}

// Synchronises, then writes up to `cap` records "name\tcount\ttotal_ms\n" (aggregated by kernel name) into buf.
extern "C" int sh_profile_collect(char *buf, size_t cap)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    SH_CHECK_CUDA(cudaDeviceSynchronize());
    std::vector<const char *> names;
    std::vector<double> total;
    std::vector<long long> count;
    for (auto &r : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
        size_t i = 0;
        for (; i < names.size(); ++i)
            if (strcmp(names[i], r.name) == 0) break;
        if (i == names.size()) { names.push_back(r.name); total.push_back(0.0); count.push_back(0); }
        total[i] += ms;
        count[i] += 1;
    }
    g_prof.clear();
    size_t off = 0;
    if (cap) buf[0] = 0;
    for (size_t i = 0; i < names.size(); ++i) {
        int n = snprintf(buf + off, cap > off ? cap - off : 0, "%s\t%lld\t%.6f\n", names[i], count[i], total[i]);
        if (n < 0 || off + (size_t)n >= cap) break;
        off += (size_t)n;
    }
    return 0;
}

extern "C" int sh_device_info(int *sms, int *cc_major, int *cc_minor)
{
    int dev = 0;
    SH_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    SH_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sms) *sms = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// stage 1 dispatcher
// ---------------------------------------------------------------------------------------------------------------
extern "C" size_t sh_discretize_workspace_bytes(int64_t R, int d, int M)
{
    return carve_disc_workspace(nullptr, R, M, d).bytes;
}

extern "C" int sh_dev_discretize(const float *tokens, const float *vocab, int64_t R, int d, int M, int64_t *out_idx,
                                 int64_t idx_rows, int64_t idx_row_stride, int64_t idx_col_stride, float *out_seq,
                                 void *workspace, size_t workspace_bytes, int mode, sh_stream_t stream)
{
    SH_REQUIRE(R > 0 && d > 0 && M > 0 && idx_rows > 0, "discretize: bad shape R=%lld d=%d M=%d", (long long)R, d, M);
    SH_REQUIRE(workspace && workspace_bytes >= sh_discretize_workspace_bytes(R, d, M), "discretize: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    DiscWorkspace ws = carve_disc_workspace(workspace, R, M, d);
    bool tensor = false, half = false;
    if (mode == SH_DISC_TENSOR || mode == SH_DISC_TENSOR_F16) {
        half = mode == SH_DISC_TENSOR_F16;
        SH_REQUIRE(discretize_tc_supported(R, d, M, half), "discretize: tensor-core path needs d %% %d == 0, d >= 32, M >= 16 (d=%d M=%d)",
                   half ? 8 : 4, d, M);
        tensor = true;
    } else if (mode == SH_DISC_AUTO) {
        half = discretize_tc_supported(R, d, M, true);
        tensor = half || discretize_tc_supported(R, d, M, false);
    }
    SH_CHECK_CUDA(cudaMemsetAsync(ws.counters, 0, 256, st));
    if (launch_codebook_norms(vocab, M, d, ws, st, tensor && half)) return 1;     // (+ the fp16 codebook copy of the f16 coarse pass)
    if (tensor) {
        if (launch_discretize_tc(tokens, vocab, R, d, M, out_idx, idx_rows, idx_row_stride, idx_col_stride, ws, half, st)) return 1;
    } else {
        if (launch_discretize_exact(tokens, vocab, ws.cn, R, d, M, out_idx, idx_rows, idx_row_stride, idx_col_stride, st)) return 1;
    }
    if (out_seq)
        if (launch_gather(vocab, out_idx, idx_rows, idx_row_stride, idx_col_stride, R, d, out_seq, st)) return 1;
    return 0;
}

extern "C" int sh_discretize_stats(const void *workspace, int64_t *recheck_rows, int64_t *overflow_rows)
{
    unsigned long long c[2] = {0, 0};
    SH_CHECK_CUDA(cudaMemcpy(c, workspace, sizeof(c), cudaMemcpyDeviceToHost));
    if (recheck_rows) *recheck_rows = (int64_t)c[0];
    if (overflow_rows) *overflow_rows = (int64_t)c[1];
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// host-buffer entry points
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct ScratchLock {
    std::lock_guard<std::mutex> g;
    ScratchLock() : g(g_scratch.mu) {}
};
int ensure_stream()
{
    if (!g_scratch.stream) SH_CHECK_CUDA(cudaStreamCreateWithFlags(&g_scratch.stream, cudaStreamNonBlocking));
    return 0;
}
}  // namespace

#define SH_SCRATCH(ptr, type, slot, count)                                                 \
    type *ptr = (type *)g_scratch.get(slot, sizeof(type) * (size_t)(count));              \
    SH_REQUIRE(ptr != nullptr, "host entry: cudaMalloc of %zu bytes failed", sizeof(type) * (size_t)(count))

extern "C" int sh_host_feat_to_instance_v(const int64_t *ingredients, const float *attn_cls, int B, int L,
                                          const float *w_vertex2, int mean, int64_t *ids, float *vertex_w,
                                          int64_t *num_vertices)
{
    ScratchLock lock;
    if (ensure_stream()) return 1;
    cudaStream_t st = g_scratch.stream;
    const size_t BL = (size_t)B * L;
    SH_SCRATCH(d_ing, int64_t, 0, BL);
    SH_SCRATCH(d_cls, float, 1, BL);
    SH_SCRATCH(d_w, float, 2, 2);
    SH_SCRATCH(d_ids, int64_t, 3, BL);
    SH_SCRATCH(d_vw, float, 4, BL);
    SH_SCRATCH(d_nv, int32_t, 5, B);
    SH_CHECK_CUDA(cudaMemcpyAsync(d_ing, ingredients, sizeof(int64_t) * BL, cudaMemcpyHostToDevice, st));
    SH_CHECK_CUDA(cudaMemcpyAsync(d_cls, attn_cls, sizeof(float) * BL, cudaMemcpyHostToDevice, st));
    SH_CHECK_CUDA(cudaMemcpyAsync(d_w, w_vertex2, sizeof(float) * 2, cudaMemcpyHostToDevice, st));
    if (sh_dev_instance_graphs(d_ing, nullptr, d_cls, nullptr, B, L, 0, SH_NO_CLAMP, SH_NO_CLAMP, d_w, nullptr,
                               mean ? 0 : SH_G_SUM, d_ids, d_vw, nullptr, d_nv, nullptr, st)) return 1;
    std::vector<int32_t> nv(B);
    SH_CHECK_CUDA(cudaMemcpyAsync(ids, d_ids, sizeof(int64_t) * BL, cudaMemcpyDeviceToHost, st));
    SH_CHECK_CUDA(cudaMemcpyAsync(vertex_w, d_vw, sizeof(float) * BL, cudaMemcpyDeviceToHost, st));
    SH_CHECK_CUDA(cudaMemcpyAsync(nv.data(), d_nv, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st));
    SH_CHECK_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b) num_vertices[b] = nv[b];
    return 0;
}

extern "C" int sh_host_feat_to_instance_e(const int64_t *ingredients, const float *attn, const float *geo_sim, int B,
                                          int L, const float *w_edge2, int mean, float *edges, int64_t *num_vertices)
{
    ScratchLock lock;
    if (ensure_stream()) return 1;
    cudaStream_t st = g_scratch.stream;
    const size_t BL = (size_t)B * L, LL = (size_t)L * L;
    SH_SCRATCH(d_ing, int64_t, 0, BL);
    SH_SCRATCH(d_w, float, 2, 2);
    SH_SCRATCH(d_nv, int32_t, 5, B);
    SH_SCRATCH(d_attn, float, 6, B * LL);
    SH_SCRATCH(d_geo, float, 7, LL);
    SH_SCRATCH(d_e, float, 8, B * LL);
    SH_CHECK_CUDA(cudaMemcpyAsync(d_ing, ingredients, sizeof(int64_t) * BL, cudaMemcpyHostToDevice, st));
    SH_CHECK_CUDA(cudaMemcpyAsync(d_attn, attn, sizeof(float) * B * LL, cudaMemcpyHostToDevice, st));
    SH_CHECK_CUDA(cudaMemcpyAsync(d_geo, geo_sim, sizeof(float) * LL, cudaMemcpyHostToDevice, st));
    SH_CHECK_CUDA(cudaMemcpyAsync(d_w, w_edge2, sizeof(float) * 2, cudaMemcpyHostToDevice, st));
    if (sh_dev_instance_graphs(d_ing, d_attn, nullptr, d_geo, B, L, 0, SH_NO_CLAMP, SH_NO_CLAMP, nullptr, d_w,
                               mean ? 0 : SH_G_SUM, nullptr, nullptr, d_e, d_nv, nullptr, st)) return 1;
    std::vector<int32_t> nv(B);
    SH_CHECK_CUDA(cudaMemcpyAsync(edges, d_e, sizeof(float) * B * LL, cudaMemcpyDeviceToHost, st));
    SH_CHECK_CUDA(cudaMemcpyAsync(nv.data(), d_nv, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st));
    SH_CHECK_CUDA(cudaStreamSynchronize(st));
    if (num_vertices)
        for (int b = 0; b < B; ++b) num_vertices[b] = nv[b];
    return 0;
}

extern "C" int sh_host_feat_to_v_attr(const int64_t *ingredients, const float *attn_cls, int B, int L, int n_vertices,
                                      int mean, int ingredients_only, float *out)
{
    ScratchLock lock;
    if (ensure_stream()) return 1;
    cudaStream_t st = g_scratch.stream;
    const size_t BL = (size_t)B * L, O = (size_t)B * n_vertices * 2;
    SH_SCRATCH(d_ing, int64_t, 0, BL);
    SH_SCRATCH(d_cls, float, 1, BL);
    SH_SCRATCH(d_out, float, 8, O);
    SH_CHECK_CUDA(cudaMemcpyAsync(d_ing, ingredients, sizeof(int64_t) * BL, cudaMemcpyHostToDevice, st));
    if (attn_cls) SH_CHECK_CUDA(cudaMemcpyAsync(d_cls, attn_cls, sizeof(float) * BL, cudaMemcpyHostToDevice, st));
    if (sh_dev_feat_to_v_attr(d_ing, attn_cls ? d_cls : nullptr, B, L, n_vertices, mean, ingredients_only, d_out, st)) return 1;
    SH_CHECK_CUDA(cudaMemcpyAsync(out, d_out, sizeof(float) * O, cudaMemcpyDeviceToHost, st));
    SH_CHECK_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int sh_host_feat_to_e(const int64_t *ingredients, const float *attn, const float *geo_sim,
                                 const int64_t *class_ingredients, const int64_t *label, int B, int L, int K, int n_max,
                                 int mean, float *out)
{
    ScratchLock lock;
    if (ensure_stream()) return 1;
    cudaStream_t st = g_scratch.stream;
    const size_t BL = (size_t)B * L, LL = (size_t)L * L, O = (size_t)B * n_max * n_max * 2;
    SH_SCRATCH(d_ing, int64_t, 0, BL);
    SH_SCRATCH(d_attn, float, 6, B * LL);
    SH_SCRATCH(d_geo, float, 7, LL);
    SH_SCRATCH(d_out, float, 8, O);
    SH_SCRATCH(d_ci, int64_t, 9, (size_t)K * n_max);
    SH_SCRATCH(d_label, int64_t, 10, B);
    SH_CHECK_CUDA(cudaMemcpyAsync(d_ing, ingredients, sizeof(int64_t) * BL, cudaMemcpyHostToDevice, st));
    SH_CHECK_CUDA(cudaMemcpyAsync(d_attn, attn, sizeof(float) * B * LL, cudaMemcpyHostToDevice, st));
    SH_CHECK_CUDA(cudaMemcpyAsync(d_geo, geo_sim, sizeof(float) * LL, cudaMemcpyHostToDevice, st));
    SH_CHECK_CUDA(cudaMemcpyAsync(d_ci, class_ingredients, sizeof(int64_t) * K * n_max, cudaMemcpyHostToDevice, st));
    SH_CHECK_CUDA(cudaMemcpyAsync(d_label, label, sizeof(int64_t) * B, cudaMemcpyHostToDevice, st));
    for (int b = 0; b < B; ++b)
        SH_REQUIRE(label[b] >= 0 && label[b] < K, "feat_to_e: label[%d]=%lld out of range", b, (long long)label[b]);
    if (sh_dev_feat_to_e(d_ing, d_attn, d_geo, d_ci, d_label, B, L, K, n_max, mean, d_out, st)) return 1;
    SH_CHECK_CUDA(cudaMemcpyAsync(out, d_out, sizeof(float) * O, cudaMemcpyDeviceToHost, st));
    SH_CHECK_CUDA(cudaStreamSynchronize(st));
    return 0;
}
