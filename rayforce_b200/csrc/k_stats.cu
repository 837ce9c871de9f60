// k_stats.cu — the order-statistic and list-valued aggregates (sm_100a), built from the sort / gather / group kernels.
//
//   rfb_group_rows_dev    aggr_row / aggr_collect (reference core/aggr.c:3021-3136): the rows of every group, in row order
//   rfb_aggr_med_launch   aggr_med  (core/aggr.c:2136-2246): per-group median of the values sorted like ray_asc
//   rfb_aggr_stddev_launch aggr_dev (core/aggr.c:2250-2864, final formula :2893-2906): per-group population deviation
//   rfb_med_dev           ray_med   (core/math.c:2529-2626), ungrouped
//   rfb_stddev_dev        ray_dev   (core/math.c:2628-2700), ungrouped (two passes: mean, then squared deviations)
//
// "Rows of every group": a stable sort of the group ids IS the grouping — the device's LSD radix sort skips the digits that
// are constant over the column, so G groups cost ceil(log256 G) passes.  Medians need every group's values in ray_asc
// order: a stable sort by value followed by a stable sort by group id leaves the rows ordered by (group, value).
#include "rfb_common.cuh"
#include "rfb_moments.cuh"

namespace {

constexpr int THREADS = 256;

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

// out[i] = src[idx[i]] for 8-byte elements
__global__ void __launch_bounds__(THREADS) k_take8(const u64 *__restrict__ src, const i64 *__restrict__ idx, i64 n, u64 *__restrict__ out) {
    for (i64 i = (i64)blockIdx.x * THREADS + threadIdx.x; i < n; i += (i64)gridDim.x * THREADS) out[i] = __ldg(src + ld_stream(idx + i));
}

// offsets[g] = first position of the sorted group-id sequence sg[0..n) holding an id >= g, g in [0, groups]
__global__ void __launch_bounds__(THREADS) k_group_bounds(const i64 *__restrict__ sg, i64 n, i64 groups, i64 *__restrict__ offsets) {
    for (i64 g = (i64)blockIdx.x * THREADS + threadIdx.x; g <= groups; g += (i64)gridDim.x * THREADS) {
        i64 lo = 0, hi = n;
        while (lo < hi) {
            const i64 mid = (lo + hi) >> 1;
            if (__ldg(sg + mid) < g) lo = mid + 1; else hi = mid;
        }
        offsets[g] = lo;
    }
}

// median of the group's sorted values (core/aggr.c:2157-2181): integers are converted before they are added
template <typename T>
__global__ void __launch_bounds__(THREADS) k_group_median(const T *__restrict__ v, const i64 *__restrict__ order, const i64 *__restrict__ offsets,
                                                          i64 groups, f64 *__restrict__ out) {
    for (i64 g = (i64)blockIdx.x * THREADS + threadIdx.x; g < groups; g += (i64)gridDim.x * THREADS) {
        const i64 o = offsets[g], l = offsets[g + 1] - o;
        if (l == 0) { out[g] = null_f64(); continue; }
        const i64 mid = l / 2;
        const f64 hi = (f64)v[order[o + mid]];
        out[g] = (l % 2 == 0) ? __dmul_rn(__dadd_rn((f64)v[order[o + mid - 1]], hi), 0.5) : hi;
    }
}

__global__ void k_fill_f64(f64 *p, i64 n, f64 v) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) p[i] = v;
}

// ---- deviation: per-group sum, sum of squares (f64, as the reference accumulates them) and non-null count: rfb_moments.cuh
template <typename T> __device__ __forceinline__ bool stat_value(T x, f64 &v) { return moments::value_of<T>(x, v); }

// core/aggr.c:2893-2906: 0 rows -> null, 1 row -> 0, else sqrt(max(sumsq/n - mean^2, 0))
__global__ void __launch_bounds__(THREADS) k_group_stddev(const f64 *sum, const f64 *sumsq, const unsigned long long *cnt, i64 groups, f64 *out) {
    for (i64 g = (i64)blockIdx.x * THREADS + threadIdx.x; g < groups; g += (i64)gridDim.x * THREADS) {
        const unsigned long long c = cnt[g];
        if (c == 0) out[g] = null_f64();
        else if (c == 1) out[g] = 0.0;
        else {
            const f64 n = (f64)c, mean = __ddiv_rn(sum[g], n);
            const f64 var = __dsub_rn(__ddiv_rn(sumsq[g], n), __dmul_rn(mean, mean));
            out[g] = var < 0.0 ? 0.0 : __dsqrt_rn(var);
        }
    }
}

// ungrouped: sum over the non-null rows of (x - mean)^2 (core/math.c:2119-2174); per-CTA partials (fixed tree) folded by the
// last CTA in index order => deterministic
template <typename T>
__global__ void __launch_bounds__(THREADS, 4) k_sq_dev(const T *__restrict__ x, i64 n, f64 mean, f64 *partials, unsigned int *ticket, f64 *out) {
    __shared__ f64 red[32];
    __shared__ bool last;
    f64 acc = 0.0;
    for (i64 i = (i64)blockIdx.x * THREADS + threadIdx.x; i < n; i += (i64)gridDim.x * THREADS) {
        f64 v;
        if (stat_value<T>(ld_stream(x + i), v)) { const f64 t = __dsub_rn(v, mean); acc = __dadd_rn(acc, __dmul_rn(t, t)); }
    }
    acc = block_reduce<f64>(acc, OpAdd(), 0.0, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = acc;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    f64 t = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += THREADS) t = __dadd_rn(t, ((volatile f64 *)partials)[i]);
    t = block_reduce<f64>(t, OpAdd(), 0.0, red);
    if (threadIdx.x == 0) { *out = t; *ticket = 0; }
}

struct Temp {   // the context's auxiliary buffer, carved up
    rfb_ctx_t *ctx;
    char *base = nullptr;
    size_t used = 0;
    explicit Temp(rfb_ctx_t *c) : ctx(c) {}
    int reserve(size_t bytes) { void *p; const int rc = rfb_ensure_aux(ctx, bytes ? bytes : 256, &p); base = (char *)p; return rc; }
    template <typename T> T *take(i64 n) { T *p = (T *)(base + used); used += align256((size_t)(n > 0 ? n : 1) * sizeof(T)); return p; }
};

int take8(rfb_ctx_t *ctx, const void *src, const i64 *idx, i64 n, void *out) {
    if (n <= 0) return RFB_OK;
    k_take8<<<rfb_grid_for(ctx, n, THREADS * 4, 8), THREADS, 0, ctx->stream>>>((const u64 *)src, idx, n, (u64 *)out);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

}  // namespace

extern "C" int rfb_group_rows_dev(rfb_ctx_t *ctx, const int64_t *group_ids, const int64_t *filter, int64_t len, int64_t groups,
                                  int64_t *out_rows, int64_t *offsets) {
    RFB_ARG(ctx && len >= 0 && groups >= 0 && offsets && ((group_ids && out_rows) || len == 0), "rfb_group_rows_dev");
    Temp t(ctx);
    int rc = t.reserve(2 * align256((size_t)(len > 0 ? len : 1) * 8));
    if (rc) { rfb_set_error("rfb_group_rows_dev: out of device memory"); return rc; }
    i64 *pos = t.take<i64>(len), *sg = t.take<i64>(len);
    if (len > 0) {
        rc = rfb_sort_dev(ctx, RFB_I64, group_ids, len, 0, pos);          // stable: positions grouped by id, ascending inside a group
        if (rc) return rc;
        rc = take8(ctx, group_ids, pos, len, sg);
        if (rc) return rc;
        if (filter) rc = take8(ctx, filter, pos, len, out_rows);          // aggr_row pushes the row id ($x = filter[i])
        else RFB_CUDA(cudaMemcpyAsync(out_rows, pos, (size_t)len * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (rc) return rc;
    }
    k_group_bounds<<<rfb_grid_for(ctx, groups + 1, THREADS, 8), THREADS, 0, ctx->stream>>>(sg, len, groups, offsets);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

// grouped median; called by rfb_aggr_dev(RFB_A_MED).  Value types without a median in the reference (anything but
// I64/TIMESTAMP/F64, core/aggr.c:2182-2184) give an all-null result, not an error.
int rfb_aggr_med_launch(rfb_ctx_t *ctx, int val_type, const void *val, const i64 *filter, const i64 *group_ids, i64 len, i64 groups, f64 *out) {
    const bool is_f = val_type == RFB_F64, is_i = val_type == RFB_I64 || val_type == RFB_TIMESTAMP;
    if (!(is_f || is_i) || len == 0) {
        k_fill_f64<<<rfb_grid_for(ctx, groups, 256, 8), 256, 0, ctx->stream>>>(out, groups, null_f64());
        RFB_CHECK_LAUNCH(ctx);
        return RFB_OK;
    }
    Temp t(ctx);
    const size_t bn = align256((size_t)len * 8);
    int rc = t.reserve(5 * bn + align256((size_t)(groups + 1) * 8));
    if (rc) { rfb_set_error("aggr med: out of device memory"); return rc; }
    u64 *v = t.take<u64>(len);
    i64 *p1 = t.take<i64>(len), *g1 = t.take<i64>(len), *p2 = t.take<i64>(len), *order = t.take<i64>(len), *offsets = t.take<i64>(groups + 1);
    const void *vals = val;
    if (filter) { rc = take8(ctx, val, filter, len, v); if (rc) return rc; vals = v; }     // the group's values in position order
    rc = rfb_sort_dev(ctx, val_type, vals, len, 0, p1);                                  // by value (ray_asc order: nulls / NaN first)
    if (rc) return rc;
    rc = take8(ctx, group_ids, p1, len, g1);
    if (rc) return rc;
    rc = rfb_sort_dev(ctx, RFB_I64, g1, len, 0, p2);                                     // then, stably, by group
    if (rc) return rc;
    rc = take8(ctx, p1, p2, len, order);                                                 // positions ordered by (group, value)
    if (rc) return rc;
    rc = take8(ctx, g1, p2, len, p1);                                                    // (p1 reused) the sorted group ids
    if (rc) return rc;
    k_group_bounds<<<rfb_grid_for(ctx, groups + 1, THREADS, 8), THREADS, 0, ctx->stream>>>(p1, len, groups, offsets);
    RFB_CHECK_LAUNCH(ctx);
    if (is_f) k_group_median<f64><<<rfb_grid_for(ctx, groups, THREADS, 8), THREADS, 0, ctx->stream>>>((const f64 *)vals, order, offsets, groups, out);
    else k_group_median<i64><<<rfb_grid_for(ctx, groups, THREADS, 8), THREADS, 0, ctx->stream>>>((const i64 *)vals, order, offsets, groups, out);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

int rfb_aggr_stddev_launch(rfb_ctx_t *ctx, int val_type, const void *val, const i64 *filter, const i64 *group_ids, i64 len, i64 groups, f64 *out) {
    void *w;
    const size_t bg = align256((size_t)(groups > 0 ? groups : 1) * 8);
    int rc = rfb_ensure_work(ctx, 3 * bg, &w);
    if (rc) return rc;
    f64 *sum = (f64 *)w, *sumsq = (f64 *)((char *)w + bg);
    unsigned long long *cnt = (unsigned long long *)((char *)w + 2 * bg);
    RFB_CUDA(cudaMemsetAsync(w, 0, 3 * bg, ctx->stream));
    switch (rfb_kind_of(val_type)) {
        case K_I16: rc = moments::launch<i16, true>(ctx, val, filter, group_ids, len, groups, sum, sumsq, cnt, nullptr); break;
        case K_I32: rc = moments::launch<i32, true>(ctx, val, filter, group_ids, len, groups, sum, sumsq, cnt, nullptr); break;
        case K_I64: rc = moments::launch<i64, true>(ctx, val, filter, group_ids, len, groups, sum, sumsq, cnt, nullptr); break;
        default: rc = moments::launch<f64, true>(ctx, val, filter, group_ids, len, groups, sum, sumsq, cnt, nullptr); break;
    }
    if (rc) return rc;
    k_group_stddev<<<rfb_grid_for(ctx, groups, THREADS, 8), THREADS, 0, ctx->stream>>>(sum, sumsq, cnt, groups, out);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

extern "C" int rfb_med_dev(rfb_ctx_t *ctx, int type, const void *x, int64_t n, double *out) {
    RFB_ARG(ctx && out && n >= 0 && (x || n == 0), "rfb_med_dev");
    // ray_med has vector cases for U8, I16 and I64 only (core/math.c:2555-2590; the I32 and F64 bodies are commented out there)
    if (!(type == RFB_U8 || type == RFB_I16 || type == RFB_I64)) { rfb_set_error("med: unsupported type %d", type); return RFB_ERR_TYPE; }
    rfb_fold_t f;
    int rc = rfb_fold_dev(ctx, RFB_F_CNT, type, x, n, &f);
    if (rc) return rc;
    const i64 l = f.nonnull;     // the reference indexes the sorted column (nulls first) with the NON-NULL count (core/math.c:2530)
    if (l == 0) { *out = null_f64(); return RFB_OK; }
    Temp t(ctx);
    rc = t.reserve(align256((size_t)n * 8));
    if (rc) { rfb_set_error("med: out of device memory"); return rc; }
    i64 *perm = t.take<i64>(n);
    rc = rfb_sort_dev(ctx, type, x, n, 0, perm);
    if (rc) return rc;
    i64 at[2] = {0, 0};
    const i64 first = l % 2 == 0 ? l / 2 - 1 : l / 2;
    RFB_CUDA(cudaMemcpyAsync(at, perm + first, (l % 2 == 0 ? 2 : 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    const int w = rfb_type_size(type);
    i64 e[2] = {0, 0};
    for (int j = 0; j < (l % 2 == 0 ? 2 : 1); j++) {
        unsigned char raw[8] = {0};
        RFB_CUDA(cudaMemcpy(raw, (const char *)x + at[j] * w, (size_t)w, cudaMemcpyDeviceToHost));
        if (type == RFB_U8) e[j] = raw[0];
        else if (type == RFB_I16) { i16 s; memcpy(&s, raw, 2); e[j] = s; }
        else memcpy(&e[j], raw, 8);
    }
    // (x[l/2-1] + x[l/2]) / 2.0 : the two elements are added as integers (int for U8/I16, i64 for I64), then halved
    *out = l % 2 == 0 ? (f64)(i64)((u64)e[0] + (u64)e[1]) / 2.0 : (f64)e[0];
    return RFB_OK;
}

extern "C" int rfb_stddev_dev(rfb_ctx_t *ctx, int type, const void *x, int64_t n, double *out) {
    RFB_ARG(ctx && out && n >= 0 && (x || n == 0), "rfb_stddev_dev");
    const int k = rfb_kind_of(type);
    // the types ray_sum accepts (core/math.c:1850-1871): ray_dev takes its mean from ray_sum, and for DATE / TIMESTAMP (a type
    // error there) it dereferences the error object — a type error here
    if (!(type == RFB_U8 || type == RFB_I16 || type == RFB_I32 || type == RFB_TIME || type == RFB_I64 || type == RFB_F64)) {
        rfb_set_error("dev: unsupported type %d", type);
        return RFB_ERR_TYPE;
    }
    rfb_fold_t f;
    int rc = rfb_fold_dev(ctx, RFB_F_SUM | RFB_F_CNT, type, x, n, &f);
    if (rc) return rc;
    const i64 l = f.nonnull;
    if (l == 0) { *out = null_f64(); return RFB_OK; }
    if (l == 1) { *out = 0.0; return RFB_OK; }
    // mean = sum / count with the sum in the width ray_sum produces (I32/DATE/TIME sums wrap in 32 bits, core/math.c:2645-2651)
    f64 mean;
    if (k == K_F64) mean = f.sum_f64 / (f64)l;
    else if (k == K_I32) mean = (f64)(i32)(u32)(u64)f.sum_i64 / (f64)l;
    else mean = (f64)f.sum_i64 / (f64)l;
    const int grid = rfb_grid_for(ctx, n, THREADS * 4, 4);
    f64 *partials = (f64 *)((char *)ctx->d_scratch + 40960);           // [grid] + ticket + result
    unsigned int *ticket = (unsigned int *)((char *)ctx->d_scratch + 40960 + 8192);
    f64 *d_out = (f64 *)((char *)ctx->d_scratch + 40960 + 8192 + 64);
    RFB_CUDA(cudaMemsetAsync(ticket, 0, 4, ctx->stream));
    switch (k) {
        case K_U8: k_sq_dev<u8><<<grid, THREADS, 0, ctx->stream>>>((const u8 *)x, n, mean, partials, ticket, d_out); break;
        case K_I16: k_sq_dev<i16><<<grid, THREADS, 0, ctx->stream>>>((const i16 *)x, n, mean, partials, ticket, d_out); break;
        case K_I32: k_sq_dev<i32><<<grid, THREADS, 0, ctx->stream>>>((const i32 *)x, n, mean, partials, ticket, d_out); break;
        case K_I64: k_sq_dev<i64><<<grid, THREADS, 0, ctx->stream>>>((const i64 *)x, n, mean, partials, ticket, d_out); break;
        default: k_sq_dev<f64><<<grid, THREADS, 0, ctx->stream>>>((const f64 *)x, n, mean, partials, ticket, d_out); break;
    }
    RFB_CHECK_LAUNCH(ctx);
    f64 ss = 0.0;
    RFB_CUDA(cudaMemcpyAsync(&ss, d_out, 8, cudaMemcpyDeviceToHost, ctx->stream));
    RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = sqrt(ss / (f64)l);
    return RFB_OK;
}
