/*
 * oracle/graph_oracle.c -- CPU restatement of the reference's native instance-graph builders.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under schemanet-pytorch_b200/ may link, import or execute this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Each function restates, in plain C with the same fp32 summation order, one pybind function of the
 * reference's `cpp_extension` (citations are relative to /root/reference):
 *   oracle_feat_to_instance_v  <- cpp_extension/src/large_scale_feat_to_v.cpp:41-143
 *   oracle_feat_to_instance_e  <- cpp_extension/src/large_scale_feat_to_e.cpp:33-150
 *   oracle_feat_to_v_attr      <- cpp_extension/src/feat_to_v_attr.cpp:19-63,74-148
 *   oracle_feat_to_e           <- cpp_extension/src/feat_to_e.cpp:31-127
 *   seq_sum                    <- cpp_extension/src/utils.cpp:6-15 (std::accumulate from 0.0f, then / size)
 * The reference finishes each image with a handful of ATen ops (max/sum, div_, nan_to_num_, matmul with a
 * [2,1] weight); those are restated here as scalar fp32 loops.
 *
 * Pinning: the reference ships no tests or golden vectors, so this restatement is pinned
 * against the reference's own C++ compiled into oracle/_ref (oracle/build_ref.py) -- see tests/test_oracle.py
 * and the fixtures written by oracle/gen_golden.py.
 *
 * Layout conventions (shared with the CUDA library so that outputs compare element for element):
 *   per-image slots: ids_out[b*L + k], w_out[b*L + k] for k < n_b; e_out[b*L*L + i*n_b + j] (compact rows).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* torch.nan_to_num(x, nan=0): NaN -> 0, +inf -> FLT_MAX, -inf -> -FLT_MAX */
static float nan_to_num0(float x)
{
    if (isnan(x)) return 0.0f;
    if (isinf(x)) return x > 0 ? FLT_MAX : -FLT_MAX;
    return x;
}

static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* sorted unique codes of one image (the iteration order of the reference's std::map<long, ...>) */
static int sorted_unique(const int64_t *codes, int L, int64_t *uniq)
{
    int n = 0, i;
    memcpy(uniq, codes, sizeof(int64_t) * (size_t)L);
    qsort(uniq, (size_t)L, sizeof(int64_t), cmp_i64);
    for (i = 0; i < L; ++i)
        if (i == 0 || uniq[i] != uniq[i - 1]) uniq[n++] = uniq[i];
    return n;
}

static int rank_of(const int64_t *uniq, int n, int64_t code)
{
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        int mid = (lo + hi) / 2;
        if (uniq[mid] == code) return mid;
        if (uniq[mid] < code) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

/* large_scale_feat_to_v.cpp:41-143.  attn_cls is already soft-maxed by the Python caller. */
void oracle_feat_to_instance_v(const int64_t *ingredients, const float *attn_cls, int B, int L,
                               const float *w /* [2] */, int mean,
                               int64_t *ids_out, float *w_out, int64_t *num_vertices)
{
    int64_t *uniq = (int64_t *)malloc(sizeof(int64_t) * (size_t)L);
    float *cnt = (float *)malloc(sizeof(float) * (size_t)L);
    float *att = (float *)malloc(sizeof(float) * (size_t)L);
    int *icnt = (int *)malloc(sizeof(int) * (size_t)L);
    int b, i, k;
    for (b = 0; b < B; ++b) {
        const int64_t *c = ingredients + (size_t)b * L;
        const float *a = attn_cls + (size_t)b * L;
        int n = sorted_unique(c, L, uniq);
        float max0 = -INFINITY, max1 = -INFINITY;
        int nan1 = 0;
        for (k = 0; k < n; ++k) { icnt[k] = 0; att[k] = 0.0f; }
        /* :78-96 -- positions are visited in order, so each code's attention list is in position order,
           and utils.cpp:9 sums it left to right starting from 0.0f */
        for (i = 0; i < L; ++i) {
            k = rank_of(uniq, n, c[i]);
            icnt[k] += 1;
            att[k] = att[k] + a[i];
        }
        for (k = 0; k < n; ++k) {
            cnt[k] = (float)icnt[k];
            if (mean) att[k] = att[k] / (float)icnt[k];
        }
        /* :124 attrs.div_(attrs.max(0, keepdim)).nan_to_num_(0); torch max propagates NaN */
        for (k = 0; k < n; ++k) {
            if (cnt[k] > max0) max0 = cnt[k];
            if (isnan(att[k])) nan1 = 1;
            if (att[k] > max1) max1 = att[k];
        }
        if (nan1) max1 = NAN;
        for (k = 0; k < n; ++k) {
            float a0 = nan_to_num0(cnt[k] / max0);
            float a1 = nan_to_num0(att[k] / max1);
            /* :125 attrs.matmul(W[2,1]) */
            w_out[(size_t)b * L + k] = a0 * w[0] + a1 * w[1];
            ids_out[(size_t)b * L + k] = uniq[k];
        }
        num_vertices[b] = n;
    }
    free(uniq); free(cnt); free(att); free(icnt);
}

/* helper: position lists per sorted code */
static void build_positions(const int64_t *c, int L, const int64_t *uniq, int n, int *start, int *pos)
{
    int i, k;
    int *fill = (int *)calloc((size_t)n + 1, sizeof(int));
    for (k = 0; k <= n; ++k) start[k] = 0;
    for (i = 0; i < L; ++i) { k = rank_of(uniq, n, c[i]); if (k >= 0) start[k + 1] += 1; }
    for (k = 0; k < n; ++k) start[k + 1] += start[k];
    for (i = 0; i < L; ++i) { k = rank_of(uniq, n, c[i]); if (k >= 0) pos[start[k] + fill[k]++] = i; }
    free(fill);
}

/* large_scale_feat_to_e.cpp:33-150.  attn is already soft-maxed; the code->rank dictionary of the reference
   is the sorted-unique rank (schema_net.py:345-348 builds it from feat_to_instance_v's ids). */
void oracle_feat_to_instance_e(const int64_t *ingredients, const float *attn, const float *geo_sim,
                               int B, int L, const float *w /* [2] */, int mean,
                               float *e_out, int64_t *num_vertices)
{
    int64_t *uniq = (int64_t *)malloc(sizeof(int64_t) * (size_t)L);
    int *start = (int *)malloc(sizeof(int) * ((size_t)L + 1));
    int *pos = (int *)malloc(sizeof(int) * (size_t)L);
    float *e0 = (float *)malloc(sizeof(float) * (size_t)L * L);
    float *e1 = (float *)malloc(sizeof(float) * (size_t)L * L);
    int b, ci, cj, pi, pj;
    for (b = 0; b < B; ++b) {
        const int64_t *c = ingredients + (size_t)b * L;
        const float *A = attn + (size_t)b * L * L;
        float *out = e_out + (size_t)b * L * L;
        int n = sorted_unique(c, L, uniq);
        build_positions(c, L, uniq, n, start, pos);
        /* :99-125 four-deep pair loop, sequential fp32 sums in (p ascending, q ascending) order */
        for (ci = 0; ci < n; ++ci)
            for (cj = 0; cj < n; ++cj) {
                float sg = 0.0f, sa = 0.0f;
                int cntp = 0;
                for (pi = start[ci]; pi < start[ci + 1]; ++pi)
                    for (pj = start[cj]; pj < start[cj + 1]; ++pj) {
                        sa = sa + A[(size_t)pos[pi] * L + pos[pj]];
                        sg = sg + geo_sim[(size_t)pos[pi] * L + pos[pj]];
                        ++cntp;
                    }
                if (mean) { sg = sg / (float)cntp; sa = sa / (float)cntp; }
                e0[ci * n + cj] = sg;
                e1[ci * n + cj] = sa;
            }
        /* :135 div_(sum(1, keepdim)).nan_to_num_(0); :140 matmul(W[2,1]) */
        for (ci = 0; ci < n; ++ci) {
            float s0 = 0.0f, s1 = 0.0f;
            for (cj = 0; cj < n; ++cj) { s0 += e0[ci * n + cj]; s1 += e1[ci * n + cj]; }
            for (cj = 0; cj < n; ++cj) {
                float a0 = nan_to_num0(e0[ci * n + cj] / s0);
                float a1 = nan_to_num0(e1[ci * n + cj] / s1);
                out[(size_t)ci * n + cj] = a0 * w[0] + a1 * w[1];
            }
        }
        if (num_vertices) num_vertices[b] = n;
    }
    free(uniq); free(start); free(pos); free(e0); free(e1);
}

/* feat_to_v_attr.cpp:19-63 (ingredients_only) and :74-148.  out: [B, n_vertices, 2], zero-initialised here. */
void oracle_feat_to_v_attr(const int64_t *ingredients, const float *attn_cls, int B, int L,
                           int n_vertices, int mean, int ingredients_only, float *out)
{
    int b, i;
    memset(out, 0, sizeof(float) * (size_t)B * n_vertices * 2);
    for (b = 0; b < B; ++b) {
        const int64_t *c = ingredients + (size_t)b * L;
        float *o = out + (size_t)b * n_vertices * 2;
        for (i = 0; i < L; ++i) {
            o[c[i] * 2 + 0] += 1.0f;                      /* counts <= L are exact in fp32 */
            if (!ingredients_only) o[c[i] * 2 + 1] = o[c[i] * 2 + 1] + attn_cls[(size_t)b * L + i];
        }
        if (!ingredients_only && mean)
            for (i = 0; i < n_vertices; ++i)
                if (o[i * 2] > 0.0f) o[i * 2 + 1] = o[i * 2 + 1] / o[i * 2];
    }
}

/* feat_to_e.cpp:31-127.  class_ingredients: [K, n_max] code ids of each class (the reference passes the
   equivalent list of {code: local index} dictionaries, schema_net.py:121-126); label: [B].
   out: [B, n_max, n_max, 2], zero-initialised here. */
void oracle_feat_to_e(const int64_t *ingredients, const float *attn, const float *geo_sim,
                      const int64_t *class_ingredients, const int64_t *label,
                      int B, int L, int K, int n_max, int mean, float *out)
{
    int64_t *uniq = (int64_t *)malloc(sizeof(int64_t) * (size_t)L);
    int64_t *kept = (int64_t *)malloc(sizeof(int64_t) * (size_t)L);
    int *local = (int *)malloc(sizeof(int) * (size_t)L);
    int *start = (int *)malloc(sizeof(int) * ((size_t)L + 1));
    int *pos = (int *)malloc(sizeof(int) * (size_t)L);
    int b, k, j, ci, cj, pi, pj;
    (void)K;
    memset(out, 0, sizeof(float) * (size_t)B * n_max * n_max * 2);
    for (b = 0; b < B; ++b) {
        const int64_t *c = ingredients + (size_t)b * L;
        const float *A = attn + (size_t)b * L * L;
        const int64_t *cls = class_ingredients + (size_t)label[b] * n_max;
        float *o = out + (size_t)b * n_max * n_max * 2;
        int n = sorted_unique(c, L, uniq), m = 0;
        /* :62-77 only codes that belong to the label's class take part */
        for (k = 0; k < n; ++k) {
            int found = -1;
            /* dictionary semantics: a later duplicate key overwrites an earlier one (schema_net.py:124) */
            for (j = 0; j < n_max; ++j) if (cls[j] == uniq[k]) found = j;
            if (found >= 0) { kept[m] = uniq[k]; local[m] = found; ++m; }
        }
        build_positions(c, L, kept, m, start, pos);
        for (ci = 0; ci < m; ++ci)
            for (cj = 0; cj < m; ++cj) {
                float sg = 0.0f, sa = 0.0f;
                int cntp = 0;
                for (pi = start[ci]; pi < start[ci + 1]; ++pi)
                    for (pj = start[cj]; pj < start[cj + 1]; ++pj) {
                        sa = sa + A[(size_t)pos[pi] * L + pos[pj]];
                        sg = sg + geo_sim[(size_t)pos[pi] * L + pos[pj]];
                        ++cntp;
                    }
                if (mean) { sg = sg / (float)cntp; sa = sa / (float)cntp; }
                o[((size_t)local[ci] * n_max + local[cj]) * 2 + 0] = sg;
                o[((size_t)local[ci] * n_max + local[cj]) * 2 + 1] = sa;
            }
    }
    free(uniq); free(kept); free(local); free(start); free(pos);
}
