// rfb_mgpu.cu — every visible B200 from ONE host process, behind the C ABI (include/rfb200.h, rfb_mgpu_*).
//
// The reference splits a column by row range over its worker threads and merges the partial aggregates (pool_split_by /
// pool_chunk_aligned, core/pool.c:450-507; the merge loops of core/math.c:2222-2228).  Here the workers are GPUs: device g takes
// rows [g*n/N, (g+1)*n/N) of the HOST column over its own PCIe link (one host thread per device drives that device's context:
// chunked cudaMemcpyAsync overlapped with the fused kernel, rfb_host.cu), and the N partial results are merged on the host —
// ungrouped folds: N rfb_fold_t records (wrapping integer sums, error-free f64 (hi, lo) pairs, min / max); group-by: the N
// (key, sum, count) lists, merged by key in device order, which is row order, so the groups keep their first-occurrence
// numbering.  No collective is needed inside one process: the partials are a few bytes (folds) or groups x 24 bytes.
#include <thread>
#include <unordered_map>
#include <vector>

#include "rfb_common.cuh"

struct rfb_mgpu {
    int n;
    rfb_ctx_t *ctx[RFB_MGPU_MAX];
};

namespace {

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

// the fold merge of rfb_host.cu's chunk pipeline, restated for device partials that were already rounded once: (sum, err) is
// still an unevaluated pair, so the merge stays error-free and is rounded once more at the end
void merge_fold(rfb_fold_t *acc, const rfb_fold_t *part, int vkind, bool first) {
    if (first) { *acc = *part; return; }
    const bool acc_empty = acc->nonnull == 0, part_empty = part->nonnull == 0;
    acc->rows += part->rows;
    acc->nonnull += part->nonnull;
    if (vkind == K_F64) {
        const double a = acc->sum_f64, b = part->sum_f64, t = a + b, bp = t - a;
        acc->sum_f64_err = acc->sum_f64_err + part->sum_f64_err + ((a - (t - bp)) + (b - bp));
        acc->sum_f64 = t;
        if (!part_empty) {
            acc->min_f64 = acc_empty ? part->min_f64 : (part->min_f64 < acc->min_f64 ? part->min_f64 : acc->min_f64);
            acc->max_f64 = acc_empty ? part->max_f64 : (part->max_f64 > acc->max_f64 ? part->max_f64 : acc->max_f64);
        }
    } else {
        i64 s = (i64)((u64)acc->sum_i64 + (u64)part->sum_i64);
        if (vkind == K_I32) s = (i64)(i32)(u32)(u64)s;   // I32 / TIME sums live in 32 bits (core/math.c:1865)
        acc->sum_i64 = s;
        if (!part_empty) {
            acc->min_i64 = acc_empty ? part->min_i64 : (part->min_i64 < acc->min_i64 ? part->min_i64 : acc->min_i64);
            acc->max_i64 = acc_empty ? part->max_i64 : (part->max_i64 > acc->max_i64 ? part->max_i64 : acc->max_i64);
        }
    }
}

// row range of device g: multiples of 16 rows so that every shard's base keeps the 16-byte alignment of the column
inline i64 shard_begin(i64 n, int parts, int g) {
    if (g >= parts) return n;
    return ((n / parts) * g) & ~15ll;
}

template <typename F> int run_on_all(rfb_mgpu *m, F f) {
    std::vector<std::thread> th;
    std::vector<int> rc(m->n, RFB_OK);
    for (int g = 0; g < m->n; g++)
        th.emplace_back([&, g] {
            if (cudaSetDevice(m->ctx[g]->device) != cudaSuccess) { rc[g] = RFB_ERR_CUDA; return; }
            rc[g] = f(g);
        });
    for (auto &t : th) t.join();
    for (int g = 0; g < m->n; g++)
        if (rc[g]) return rc[g];
    return RFB_OK;
}

}  // namespace

extern "C" int rfb_mgpu_create(int ndev, rfb_mgpu_t **out) {
    RFB_ARG(out, "rfb_mgpu_create: out");
    *out = nullptr;
    const int have = rfb_device_count();
    if (have <= 0) { rfb_set_error("no CUDA device available: librfb200 has no CPU fallback"); return RFB_ERR_CUDA; }
    if (ndev <= 0 || ndev > have) ndev = have;
    if (ndev > RFB_MGPU_MAX) ndev = RFB_MGPU_MAX;
    rfb_mgpu *m = (rfb_mgpu *)calloc(1, sizeof(rfb_mgpu));
    if (!m) return RFB_ERR_NOMEM;
    for (int g = 0; g < ndev; g++) {
        int rc = rfb_ctx_create(g, &m->ctx[g]);
        if (rc) { for (int j = 0; j < g; j++) rfb_ctx_destroy(m->ctx[j]); free(m); return rc; }
        m->n = g + 1;
    }
    *out = m;
    return RFB_OK;
}

extern "C" void rfb_mgpu_destroy(rfb_mgpu_t *m) {
    if (!m) return;
    for (int g = 0; g < m->n; g++) rfb_ctx_destroy(m->ctx[g]);
    free(m);
}

extern "C" int rfb_mgpu_devices(const rfb_mgpu_t *m) { return m ? m->n : 0; }
extern "C" rfb_ctx_t *rfb_mgpu_ctx(rfb_mgpu_t *m, int g) { return (m && g >= 0 && g < m->n) ? m->ctx[g] : nullptr; }

extern "C" int rfb_mgpu_filter_fold_host(rfb_mgpu_t *m, int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k, int folds,
                                         int val_type, const void *val, int64_t n, int64_t chunk_rows, rfb_fold_t *out, int64_t *h2d_bytes) {
    RFB_ARG(m && m->n > 0 && out && n >= 0 && (val || n == 0) && (!pred || k), "rfb_mgpu_filter_fold_host");
    const int vsz = rfb_type_size(val_type), psz = pred ? rfb_type_size(pred_type) : 0;
    if (!vsz || (pred && !psz)) { rfb_set_error("mgpu fold: unsupported element type"); return RFB_ERR_TYPE; }
    const int parts = (n < (i64)m->n * 65536) ? 1 : m->n;     // small columns: one device, the rest would only add latency
    std::vector<rfb_fold_t> part(parts);
    std::vector<i64> bytes(parts, 0);
    rfb_mgpu sub = *m;
    sub.n = parts;
    int rc = run_on_all(&sub, [&](int g) {
        const i64 r0 = shard_begin(n, parts, g), r1 = shard_begin(n, parts, g + 1);
        const char *v = (const char *)val + r0 * vsz;
        if (pred) return rfb_filter_fold_host(m->ctx[g], cmp_op, pred_type, (const char *)pred + r0 * psz, k, folds, val_type, v, r1 - r0, chunk_rows, &part[g], &bytes[g]);
        return rfb_fold_host(m->ctx[g], folds, val_type, v, r1 - r0, chunk_rows, &part[g], &bytes[g]);
    });
    if (rc) return rc;
    const int vk = rfb_kind_of(val_type);
    i64 total = 0;
    for (int g = 0; g < parts; g++) { merge_fold(out, &part[g], vk, g == 0); total += bytes[g]; }
    if (vk == K_F64) {
        const double hi = out->sum_f64, lo = out->sum_f64_err, s = hi + lo;
        out->sum_f64_err = (hi - s) + lo;
        out->sum_f64 = s;
    }
    if (h2d_bytes) *h2d_bytes = total;
    return RFB_OK;
}

// ---- host-layer entries with several columns (one device): the columns are shipped whole (pinned or staged copies on the copy
// stream), then the fused device entry point runs.  HBM holds them easily (1e9 rows x 24 B = 24 of 180 GB).

// select {s: (sum v) c: (count v) from t by k [where (cmp p c)]}: HOST columns in, HOST group lists out
extern "C" int rfb_group_sum_count_host(rfb_ctx_t *ctx, int key_type, const void *keys, const int64_t *val, int64_t n, int cmp_op, int pred_type,
                                        const void *pred, const rfb_scalar_t *k, int64_t max_groups, int64_t *out_keys, int64_t *out_sums,
                                        int64_t *out_counts, int64_t *groups, int64_t *h2d_bytes) {
    RFB_ARG(ctx && groups && n >= 0 && max_groups >= 0 && ((keys && val) || n == 0) && (!pred || k), "rfb_group_sum_count_host");
    RFB_ARG((out_keys && out_sums && out_counts) || max_groups == 0, "rfb_group_sum_count_host: outputs");
    const int ksz = rfb_type_size(key_type), psz = pred ? rfb_type_size(pred_type) : 0;
    if (!(ksz == 4 || ksz == 8) || (pred && !psz)) { rfb_set_error("host group-by: unsupported key / predicate type"); return RFB_ERR_TYPE; }
    *groups = 0;
    if (h2d_bytes) *h2d_bytes = 0;
    if (n == 0) return RFB_OK;
    const bool pred_is_val = pred && (const void *)pred == (const void *)val;
    const size_t kb = align256((size_t)n * ksz), vb = align256((size_t)n * 8), pb = (pred && !pred_is_val) ? align256((size_t)n * psz) : 0;
    const i64 cap = max_groups < n ? max_groups : n;
    const size_t ob = align256((size_t)(cap > 0 ? cap : 1) * 8);
    void *buf;
    int rc = rfb_ensure_aux2(ctx, kb + vb + pb + 3 * ob, &buf);
    if (rc) return rc;
    char *dk = (char *)buf, *dv = dk + kb, *dp = dv + vb, *dok = dp + pb, *dos = dok + ob, *doc = dos + ob;
    rc = rfb_copy_h2d(ctx, dk, keys, (size_t)n * ksz, ctx->copy_stream);
    if (!rc) rc = rfb_copy_h2d(ctx, dv, val, (size_t)n * 8, ctx->copy_stream);
    if (!rc && pb) rc = rfb_copy_h2d(ctx, dp, pred, (size_t)n * psz, ctx->copy_stream);
    if (rc) return rc;
    RFB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    if (h2d_bytes) *h2d_bytes = n * (ksz + 8) + (pb ? n * psz : 0);
    i64 g = 0;
    rc = rfb_group_sum_count_dev(ctx, key_type, dk, (const i64 *)dv, n, cmp_op, pred_type, pred ? (pred_is_val ? (const void *)dv : (const void *)dp) : nullptr, k,
                                 cap, (i64 *)dok, (i64 *)dos, (i64 *)doc, &g);
    if (rc) return rc;
    if (g) {
        RFB_CUDA(cudaMemcpyAsync(out_keys, dok, (size_t)g * 8, cudaMemcpyDeviceToHost, ctx->stream));
        RFB_CUDA(cudaMemcpyAsync(out_sums, dos, (size_t)g * 8, cudaMemcpyDeviceToHost, ctx->stream));
        RFB_CUDA(cudaMemcpyAsync(out_counts, doc, (size_t)g * 8, cudaMemcpyDeviceToHost, ctx->stream));
        RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    *groups = g;
    return RFB_OK;
}

// (fold (+ (* a b) c)) over three HOST F64 columns
extern "C" int rfb_fma_fold_host(rfb_ctx_t *ctx, int folds, const double *a, const double *b, const double *c, int64_t n, rfb_fold_t *out,
                                 int64_t *h2d_bytes) {
    RFB_ARG(ctx && out && n >= 0 && ((a && b && c) || n == 0), "rfb_fma_fold_host");
    const size_t cb = align256((size_t)(n > 0 ? n : 1) * 8);
    void *buf;
    int rc = rfb_ensure_aux2(ctx, 3 * cb, &buf);
    if (rc) return rc;
    char *d = (char *)buf;
    if (n > 0) {
        rc = rfb_copy_h2d(ctx, d, a, (size_t)n * 8, ctx->copy_stream);
        if (!rc) rc = rfb_copy_h2d(ctx, d + cb, b, (size_t)n * 8, ctx->copy_stream);
        if (!rc) rc = rfb_copy_h2d(ctx, d + 2 * cb, c, (size_t)n * 8, ctx->copy_stream);
        if (rc) return rc;
        RFB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    }
    if (h2d_bytes) *h2d_bytes = 3 * n * 8;
    return rfb_fma_fold_dev(ctx, folds, (const double *)d, (const double *)(d + cb), (const double *)(d + 2 * cb), n, out);
}

// select {s: (sum v) c: (count v) from t by k [where (cmp p c)]} over HOST columns: every device ships its row range, groups it
// (rfb_group_sum_count_dev), and the per-device group lists are merged by key on the host in device (= row) order
extern "C" int rfb_mgpu_group_sum_count_host(rfb_mgpu_t *m, int key_type, const void *keys, const int64_t *val, int64_t n, int cmp_op,
                                             int pred_type, const void *pred, const rfb_scalar_t *k, int64_t max_groups, int64_t *out_keys,
                                             int64_t *out_sums, int64_t *out_counts, int64_t *groups, int64_t *h2d_bytes) {
    RFB_ARG(m && m->n > 0 && groups && n >= 0 && max_groups >= 0 && ((keys && val) || n == 0) && (!pred || k), "rfb_mgpu_group_sum_count_host");
    RFB_ARG((out_keys && out_sums && out_counts) || max_groups == 0, "rfb_mgpu_group_sum_count_host: outputs");
    const int ksz = rfb_type_size(key_type), psz = pred ? rfb_type_size(pred_type) : 0;
    if (!(ksz == 4 || ksz == 8) || (pred && !psz)) { rfb_set_error("mgpu group-by: unsupported key / predicate type"); return RFB_ERR_TYPE; }
    *groups = 0;
    if (n == 0) return RFB_OK;
    const int parts = (n < (i64)m->n * 65536) ? 1 : m->n;
    struct Part { std::vector<i64> k, s, c; i64 groups = 0, bytes = 0; };
    std::vector<Part> part(parts);
    rfb_mgpu sub = *m;
    sub.n = parts;
    const bool pred_is_val = pred && (const void *)pred == (const void *)val;
    int rc = run_on_all(&sub, [&](int g) -> int {
        const i64 r0 = shard_begin(n, parts, g), rows = shard_begin(n, parts, g + 1) - r0;
        Part &p = part[g];
        if (rows == 0) return RFB_OK;
        const i64 cap = max_groups < rows ? max_groups : rows;
        p.k.resize((size_t)cap); p.s.resize((size_t)cap); p.c.resize((size_t)cap);
        const void *pp = pred ? (pred_is_val ? (const void *)(val + r0) : (const void *)((const char *)pred + r0 * psz)) : nullptr;
        return rfb_group_sum_count_host(m->ctx[g], key_type, (const char *)keys + r0 * ksz, val + r0, rows, cmp_op, pred_type, pp, k, cap,
                                        p.k.data(), p.s.data(), p.c.data(), &p.groups, &p.bytes);
    });
    if (rc) return rc;
    // merge in device order = row order: a key keeps the number of its first occurrence; sums are sticky-null and wrap
    std::unordered_map<i64, i64> slot;
    i64 g_out = 0, total = 0;
    for (int g = 0; g < parts; g++) {
        const Part &p = part[g];
        total += p.bytes;
        for (i64 i = 0; i < p.groups; i++) {
            auto it = slot.find(p.k[(size_t)i]);
            if (it == slot.end()) {
                if (g_out >= max_groups) { rfb_set_error("mgpu group-by: more than %lld groups", (long long)max_groups); return RFB_ERR_ARG; }
                slot.emplace(p.k[(size_t)i], g_out);
                out_keys[g_out] = p.k[(size_t)i]; out_sums[g_out] = p.s[(size_t)i]; out_counts[g_out] = p.c[(size_t)i];
                g_out++;
            } else {
                const i64 o = it->second, a = out_sums[o], b = p.s[(size_t)i];
                out_sums[o] = (a == NULL_I64 || b == NULL_I64) ? NULL_I64 : (i64)((u64)a + (u64)b);
                out_counts[o] += p.c[(size_t)i];
            }
        }
    }
    *groups = g_out;
    if (h2d_bytes) *h2d_bytes = total;
    return RFB_OK;
}
