// k_join.cu — equi-join row matching (sm_100a): SURVEY §8(f) rank 4.
//
//   rfb_find_rows_dev    ray_find -> index_find_i64 (reference core/index.c:1507-1574) for one key column and
//                        index_left_join_obj (core/index.c:2886-2928) for a tuple of key columns: for every probe row the
//                        FIRST build row with an equal key (tuple), else NULL_I64
//   rfb_inner_join_dev   index_inner_join_obj (core/index.c:2930-3000): the matching (probe row, build row) pairs in
//                        ascending probe-row order
//   rfb_asof_join_dev    index_asof_join_obj (core/index.c:3194-3268): per probe row the LAST build row of its key whose
//                        time is <= the probe row's time (binary search over the key's build rows in row order)
//
// The reference inserts the build rows sequentially into an open-addressing table that keeps the first row of every key
// (core/index.c:2905-2909, 1558-1564) and probes it row by row.  The device builds one table of REPRESENTATIVE build rows
// (a slot is claimed with a 64-bit CAS on the row id; a probe compares the whole key tuple against the slot's
// representative), keeps min(row) per slot with an atomic minimum — the first row, without the sequential order — and probes
// it with one thread per probe row.  Keys compare by bit pattern, a null is a key like any other (__index_list_cmp_row,
// core/index.c:59-105).
#include "rfb_common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int MAX_KEY_COLS = 8;

__host__ __device__ __forceinline__ u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

struct KeyCols {
    const i64 *col[MAX_KEY_COLS];
    int ncols;
    __device__ __forceinline__ u64 hash(i64 row) const {
        u64 h = 0x9E3779B97F4A7C15ULL;
        for (int c = 0; c < ncols; c++) h = mix64(h ^ (u64)__ldg(col[c] + row)) + 0x9E3779B97F4A7C15ULL;
        return h;
    }
};
__device__ __forceinline__ bool same_tuple(const KeyCols &a, i64 ra, const KeyCols &b, i64 rb) {
    for (int c = 0; c < a.ncols; c++)
        if (__ldg(a.col[c] + ra) != __ldg(b.col[c] + rb)) return false;
    return true;
}

__device__ __forceinline__ u64 ld_relaxed(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct JoinTable {
    i64 *rep;       // [cap] representative build row of the slot, NULL_I64 = empty (row ids are >= 0)
    u64 *first;     // [cap] smallest build row with the slot's key
    u64 mask;
};

__global__ void __launch_bounds__(THREADS, 4) k_join_build(KeyCols build, i64 n, JoinTable t) {
    for (i64 row = (i64)blockIdx.x * THREADS + threadIdx.x; row < n; row += (i64)gridDim.x * THREADS) {
        u64 s = build.hash(row) & t.mask;
        while (true) {
            i64 cur = (i64)ld_relaxed((const u64 *)&t.rep[s]);
            if (cur == NULL_I64) {
                cur = (i64)atomicCAS((unsigned long long *)&t.rep[s], (unsigned long long)NULL_I64, (unsigned long long)row);
                if (cur == NULL_I64) break;
            }
            if (cur == row || same_tuple(build, cur, build, row)) break;
            s = (s + 1) & t.mask;
        }
        if (__ldcg(&t.first[s]) > (u64)row) atomicMin((unsigned long long *)&t.first[s], (unsigned long long)row);
    }
}

__global__ void __launch_bounds__(THREADS, 4) k_join_probe(KeyCols build, KeyCols probe, i64 n, JoinTable t, i64 *__restrict__ ids) {
    for (i64 row = (i64)blockIdx.x * THREADS + threadIdx.x; row < n; row += (i64)gridDim.x * THREADS) {
        u64 s = probe.hash(row) & t.mask;
        i64 found = NULL_I64;
        while (true) {
            const i64 cur = __ldg(&t.rep[s]);
            if (cur == NULL_I64) break;
            if (same_tuple(build, cur, probe, row)) { found = (i64)__ldg(&t.first[s]); break; }
            s = (s + 1) & t.mask;
        }
        __stcs(ids + row, found);
    }
}

__global__ void k_fill_i64(i64 *p, i64 n, i64 v) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void __launch_bounds__(THREADS) k_take_i64(const i64 *__restrict__ src, const i64 *__restrict__ idx, i64 n, i64 *__restrict__ out) {
    for (i64 i = (i64)blockIdx.x * THREADS + threadIdx.x; i < n; i += (i64)gridDim.x * THREADS) out[i] = __ldg(src + ld_stream(idx + i));
}

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

extern "C" int rfb_find_rows_dev(rfb_ctx_t *ctx, int ncols, const int64_t *const *build_cols, int64_t build_len,
                                 const int64_t *const *probe_cols, int64_t probe_len, int64_t *ids) {
    RFB_ARG(ctx && ncols >= 1 && ncols <= MAX_KEY_COLS && build_len >= 0 && probe_len >= 0 && build_cols && probe_cols && (ids || probe_len == 0),
            "rfb_find_rows_dev");
    for (int c = 0; c < ncols; c++) RFB_ARG((build_cols[c] || build_len == 0) && (probe_cols[c] || probe_len == 0), "rfb_find_rows_dev: key column");
    if (probe_len == 0) return RFB_OK;
    if (build_len == 0) {
        k_fill_i64<<<rfb_grid_for(ctx, probe_len, 256, 8), 256, 0, ctx->stream>>>(ids, probe_len, NULL_I64);
        RFB_CHECK_LAUNCH(ctx);
        return RFB_OK;
    }
    KeyCols b, p;
    b.ncols = p.ncols = ncols;
    for (int c = 0; c < ncols; c++) { b.col[c] = build_cols[c]; p.col[c] = probe_cols[c]; }
    i64 cap = 1024;
    while (cap < 2 * build_len) cap <<= 1;          // load factor <= 0.5
    void *w;
    const size_t b1 = align256((size_t)cap * 8);
    int rc = rfb_ensure_work(ctx, 2 * b1, &w);
    if (rc) return rc;
    JoinTable t{(i64 *)w, (u64 *)((char *)w + b1), (u64)(cap - 1)};
    k_fill_i64<<<rfb_grid_for(ctx, cap, 256, 8), 256, 0, ctx->stream>>>(t.rep, cap, NULL_I64);
    RFB_CHECK_LAUNCH(ctx);
    RFB_CUDA(cudaMemsetAsync(t.first, 0xFF, (size_t)cap * 8, ctx->stream));
    k_join_build<<<rfb_grid_for(ctx, build_len, THREADS * 4, 4), THREADS, 0, ctx->stream>>>(b, build_len, t);
    RFB_CHECK_LAUNCH(ctx);
    k_join_probe<<<rfb_grid_for(ctx, probe_len, THREADS * 4, 4), THREADS, 0, ctx->stream>>>(b, p, probe_len, t, ids);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

extern "C" int rfb_inner_join_dev(rfb_ctx_t *ctx, int ncols, const int64_t *const *build_cols, int64_t build_len,
                                  const int64_t *const *probe_cols, int64_t probe_len, int64_t *probe_ids, int64_t *build_ids,
                                  int64_t *count) {
    RFB_ARG(ctx && count && probe_len >= 0 && ((probe_ids && build_ids) || probe_len == 0), "rfb_inner_join_dev");
    *count = 0;
    if (probe_len == 0) return RFB_OK;
    void *aux;
    int rc = rfb_ensure_aux(ctx, (size_t)probe_len * 8, &aux);
    if (rc) return rc;
    i64 *ids = (i64 *)aux;
    rc = rfb_find_rows_dev(ctx, ncols, build_cols, build_len, probe_cols, probe_len, ids);
    if (rc) return rc;
    rfb_scalar_t none;
    memset(&none, 0, sizeof(none));
    none.type = RFB_I64;
    none.v.i64 = NULL_I64;
    rc = rfb_cmp_where_dev(ctx, RFB_NE, RFB_I64, ids, probe_len, &none, probe_ids, count);      // matched probe rows, ascending
    if (rc) return rc;
    if (*count > 0) {
        k_take_i64<<<rfb_grid_for(ctx, *count, THREADS * 4, 8), THREADS, 0, ctx->stream>>>(ids, probe_ids, *count, build_ids);
        RFB_CHECK_LAUNCH(ctx);
    }
    return RFB_OK;
}

// ------------------------------------------------------------------ asof join
//
// The reference keeps, per key tuple, the list of build rows in row order (push_raw, core/index.c:3219-3223) and binary-searches
// it for the last entry whose time is <= the probe time (index_bin_i64 / index_bin_i32, core/index.c:3103-3138) — the search
// itself assumes the key's build rows are ordered by time.  The device gets the same lists from the grouping kernels (group
// the build rows by key tuple, order them by group with the stable sort), finds a probe row's key group through the join
// table (first build row of the key -> its group id) and runs the identical search.
namespace {
template <typename T>
__global__ void __launch_bounds__(THREADS, 4)
k_asof_search(const i64 *__restrict__ first, const i64 *__restrict__ gid_of_build, const i64 *__restrict__ rows, const i64 *__restrict__ offsets,
              const T *__restrict__ build_time, const T *__restrict__ probe_time, i64 n, i64 *__restrict__ ids) {
    for (i64 i = (i64)blockIdx.x * THREADS + threadIdx.x; i < n; i += (i64)gridDim.x * THREADS) {
        const i64 f = ld_stream(first + i);
        i64 found = NULL_I64;
        if (f != NULL_I64) {
            const i64 g = __ldg(gid_of_build + f);
            const i64 base = __ldg(offsets + g), len = __ldg(offsets + g + 1) - base;
            const T t = ld_stream(probe_time + i);
            i64 left = 0, right = len - 1, idx = -1;
            while (left <= right) {
                const i64 mid = left + (right - left) / 2;
                if (__ldg(build_time + __ldg(rows + base + mid)) <= t) { idx = mid; left = mid + 1; } else right = mid - 1;
            }
            if (idx >= 0) found = __ldg(rows + base + idx);
        }
        __stcs(ids + i, found);
    }
}
}  // namespace

extern "C" int rfb_asof_join_dev(rfb_ctx_t *ctx, int ncols, const int64_t *const *build_cols, int time_type, const void *build_time,
                                 int64_t build_len, const int64_t *const *probe_cols, const void *probe_time, int64_t probe_len, int64_t *ids) {
    RFB_ARG(ctx && ncols >= 1 && ncols <= MAX_KEY_COLS && build_len >= 0 && probe_len >= 0 && build_cols && probe_cols && (ids || probe_len == 0) &&
            (build_time || build_len == 0) && (probe_time || probe_len == 0), "rfb_asof_join_dev");
    const int tk = rfb_kind_of(time_type);
    if (!(tk == K_I64 || tk == K_I32) || time_type == RFB_SYMBOL) { rfb_set_error("asof join: time column type %d (I32/DATE/TIME or I64/TIMESTAMP)", time_type); return RFB_ERR_TYPE; }
    if (probe_len == 0) return RFB_OK;
    if (build_len == 0) {
        k_fill_i64<<<rfb_grid_for(ctx, probe_len, 256, 8), 256, 0, ctx->stream>>>(ids, probe_len, NULL_I64);
        RFB_CHECK_LAUNCH(ctx);
        return RFB_OK;
    }
    void *buf;
    const size_t bb = align256((size_t)build_len * 8), bp = align256((size_t)probe_len * 8);
    int rc = rfb_ensure_aux2(ctx, 4 * bb + 256 + bp, &buf);
    if (rc) return rc;
    i64 *gids = (i64 *)buf, *firsts = (i64 *)((char *)buf + bb), *rows = (i64 *)((char *)buf + 2 * bb), *offsets = (i64 *)((char *)buf + 3 * bb);
    i64 *first = (i64 *)((char *)buf + 4 * bb + 256);
    rfb_group_info_t info;
    rc = rfb_group_keys_i64_dev(ctx, ncols, build_cols, nullptr, build_len, gids, firsts, &info);
    if (rc) return rc;
    rc = rfb_group_rows_dev(ctx, gids, nullptr, build_len, info.groups, rows, offsets);
    if (rc) return rc;
    rc = rfb_find_rows_dev(ctx, ncols, build_cols, build_len, probe_cols, probe_len, first);
    if (rc) return rc;
    const int grid = rfb_grid_for(ctx, probe_len, THREADS * 2, 4);
    if (tk == K_I64) k_asof_search<i64><<<grid, THREADS, 0, ctx->stream>>>(first, gids, rows, offsets, (const i64 *)build_time, (const i64 *)probe_time, probe_len, ids);
    else k_asof_search<i32><<<grid, THREADS, 0, ctx->stream>>>(first, gids, rows, offsets, (const i32 *)build_time, (const i32 *)probe_time, probe_len, ids);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

// ------------------------------------------------------------------ window join
//
// The reference (core/join.c:358-485) sorts the right table by (key tuple, time), so the rows of one key form one block; its index
// keeps the block's first and last row per key (core/index.c:3303-3316) and AGGR_ITER's WINDOW branch (core/aggr.c:131-160)
// binary-searches the block on the 4-byte time column for every left row's window and folds the rows inside it.  The device takes
// the same sorted right table: the block's first row comes from the join table (first build row of the key), its last row from
// a gallop + binary search for the last row whose key tuple still equals the first row's, then the reference's two searches
// (indexr_bin / indexl_bin, core/aggr.c:39-72) and the grouped partial of the aggregate over the window's rows.
namespace {
enum { W_SUM = 0, W_MIN = 1, W_MAX = 2, W_COUNT = 3, W_AVG = 4 };

__device__ __forceinline__ i64 bin_r(i32 val, const i32 *__restrict__ t, i64 offset, i64 len) {   // last row with time <= val, else the first
    i64 left = 0, right = len - 1, idx = 0;
    while (left <= right) { const i64 mid = left + (right - left) / 2; if (__ldg(t + offset + mid) <= val) { idx = mid; left = mid + 1; } else right = mid - 1; }
    return idx + offset;
}
__device__ __forceinline__ i64 bin_l(i32 val, const i32 *__restrict__ t, i64 offset, i64 len) {   // first row with time >= val, else the first
    i64 left = 0, right = len - 1, idx = 0;
    while (left <= right) { const i64 mid = left + (right - left) / 2; if (__ldg(t + offset + mid) < val) left = mid + 1; else { idx = mid; right = mid - 1; } }
    return idx + offset;
}

template <int OP, typename V>
__global__ void __launch_bounds__(THREADS, 4)
k_window_fold(KeyCols right, i64 rl, const i32 *__restrict__ rtime, const V *__restrict__ val, const i64 *__restrict__ first, const i64 *__restrict__ last,
              const i32 *__restrict__ wlo, const i32 *__restrict__ whi, i64 ll, int jtype, void *out) {
    for (i64 i = (i64)blockIdx.x * THREADS + threadIdx.x; i < ll; i += (i64)gridDim.x * THREADS) {
        const i64 f = ld_stream(first + i);
        bool none = f == NULL_I64;
        i64 li = 0, ri = -1;
        if (!none) {
            i64 lo = f;
            if (last) lo = ld_stream(last + i);                // the reference's index carries the block's last row
            else {
                // last row of the key's block: gallop, then binary search, on "same key tuple as row f"
                i64 step = 1;
                while (lo + step < rl && same_tuple(right, f, right, lo + step)) { lo += step; step <<= 1; }
                i64 hi = lo + step < rl ? lo + step : rl;      // row lo is in the block, row hi (if < rl) is not
                while (lo + 1 < hi) { const i64 mid = lo + (hi - lo) / 2; if (same_tuple(right, f, right, mid)) lo = mid; else hi = mid; }
            }
            const i64 n = lo - f + 1;
            const i32 a = ld_stream(wlo + i), b = ld_stream(whi + i);
            li = jtype == 0 ? bin_r(a, rtime, f, n) : bin_l(a, rtime, f, n);
            ri = bin_r(b, rtime, f, n);
            if (__ldg(rtime + li) > b || (jtype == 1 && __ldg(rtime + ri) < a)) none = true;
        }
        if constexpr (OP == W_COUNT) { ((i64 *)out)[i] = none ? 0 : (ri - li + 1 > 0 ? ri - li + 1 : 0); }
        else if constexpr (OP == W_AVG) {      // f64 sum of the window's non-null values in row order / their count (core/aggr.c:1545-1575, :2060)
            f64 so = 0.0;
            i64 co = 0;
            if (!none)
                for (i64 x = li; x <= ri; x++) {
                    const V v = __ldg(val + x);
                    if (!Elem<V>::is_null(v)) { so = __dadd_rn(so, (f64)v); co++; }
                }
            ((f64 *)out)[i] = co == 0 ? null_f64() : __ddiv_rn(so, (f64)co);
        } else if constexpr (Elem<V>::kind == K_F64) {
            f64 acc = null_f64();
            if (!none) {
                acc = OP == W_SUM ? 0.0 : (OP == W_MIN ? bits_f64(0x7FF0000000000000ULL) : null_f64());
                for (i64 x = li; x <= ri; x++) {
                    const f64 v = __ldg(val + x);
                    if (OP == W_SUM) acc = (isnan64(acc) || isnan64(v)) ? null_f64() : acc + v;
                    else if (OP == W_MIN) acc = OpMinF()(acc, v);
                    else acc = OpMaxF()(acc, v);
                }
            }
            ((f64 *)out)[i] = acc;
        } else {
            i64 acc = NULL_I64;
            if (!none) {
                acc = OP == W_SUM ? 0 : (OP == W_MIN ? RFB_INF_I64 : NULL_I64);
                for (i64 x = li; x <= ri; x++) {
                    const i64 v = __ldg(val + x);
                    if (OP == W_SUM) acc = (acc == NULL_I64 || v == NULL_I64) ? NULL_I64 : (i64)((u64)acc + (u64)v);
                    else if (OP == W_MIN) acc = OpMinI()(acc, v);
                    else acc = OpMaxI()(acc, v);
                }
            }
            ((i64 *)out)[i] = acc;
        }
    }
}

template <typename V>
int window_launch(rfb_ctx_t *ctx, int op, KeyCols r, i64 rl, const i32 *rtime, const void *val, const i64 *first, const i64 *last, const i32 *wlo,
                  const i32 *whi, i64 ll, int jtype, void *out) {
    const int grid = rfb_grid_for(ctx, ll, THREADS, 4);
    switch (op) {
        case RFB_A_SUM: k_window_fold<W_SUM, V><<<grid, THREADS, 0, ctx->stream>>>(r, rl, rtime, (const V *)val, first, last, wlo, whi, ll, jtype, out); break;
        case RFB_A_MIN: k_window_fold<W_MIN, V><<<grid, THREADS, 0, ctx->stream>>>(r, rl, rtime, (const V *)val, first, last, wlo, whi, ll, jtype, out); break;
        case RFB_A_MAX: k_window_fold<W_MAX, V><<<grid, THREADS, 0, ctx->stream>>>(r, rl, rtime, (const V *)val, first, last, wlo, whi, ll, jtype, out); break;
        case RFB_A_AVG: k_window_fold<W_AVG, V><<<grid, THREADS, 0, ctx->stream>>>(r, rl, rtime, (const V *)val, first, last, wlo, whi, ll, jtype, out); break;
        default: k_window_fold<W_COUNT, V><<<grid, THREADS, 0, ctx->stream>>>(r, rl, rtime, (const V *)val, first, last, wlo, whi, ll, jtype, out); break;
    }
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}
}  // namespace

extern "C" int rfb_window_join_dev(rfb_ctx_t *ctx, int ncols, const int64_t *const *right_cols, const int32_t *right_time, int64_t right_len,
                                   const int64_t *const *left_cols, int64_t left_len, const int32_t *win_lo, const int32_t *win_hi, int jtype,
                                   int op, int val_type, const void *val, void *out) {
    RFB_ARG(ctx && ncols >= 1 && ncols <= MAX_KEY_COLS && right_len >= 0 && left_len >= 0 && right_cols && left_cols && (jtype == 0 || jtype == 1) &&
            ((win_lo && win_hi && out) || left_len == 0) && ((right_time && val) || right_len == 0), "rfb_window_join_dev");
    const int vk = rfb_kind_of(val_type);
    if (!(vk == K_I64 || vk == K_F64) || val_type == RFB_SYMBOL || !(op == RFB_A_SUM || op == RFB_A_MIN || op == RFB_A_MAX || op == RFB_A_COUNT || op == RFB_A_AVG)) {
        rfb_set_error("window join: aggregate %d over value type %d (sum / min / max / count / avg of I64-kind or F64 values)", op, val_type);
        return RFB_ERR_TYPE;
    }
    if (left_len == 0) return RFB_OK;
    void *buf;
    int rc = rfb_ensure_aux2(ctx, (size_t)left_len * 8, &buf);
    if (rc) return rc;
    i64 *first = (i64 *)buf;
    rc = rfb_find_rows_dev(ctx, ncols, right_cols, right_len, left_cols, left_len, first);
    if (rc) return rc;
    KeyCols r;
    r.ncols = ncols;
    for (int c = 0; c < ncols; c++) r.col[c] = right_cols[c];
    if (vk == K_F64) return window_launch<f64>(ctx, op, r, right_len, right_time, val, first, nullptr, win_lo, win_hi, left_len, jtype, out);
    return window_launch<i64>(ctx, op, r, right_len, right_time, val, first, nullptr, win_lo, win_hi, left_len, jtype, out);
}

// the aggregate alone, over an index that already holds every left row's block [first, last] (NULL_I64 = no block): what aggr_*
// receive from the reference's index_window_join_obj (core/index.c:3287-3346, AGGR_ITER's WINDOW branch core/aggr.c:131-160)
extern "C" int rfb_window_aggr_dev(rfb_ctx_t *ctx, const int32_t *right_time, const int64_t *first, const int64_t *last, int64_t left_len,
                                   const int32_t *win_lo, const int32_t *win_hi, int jtype, int op, int val_type, const void *val, void *out) {
    RFB_ARG(ctx && left_len >= 0 && (jtype == 0 || jtype == 1) && ((first && last && win_lo && win_hi && out && right_time && val) || left_len == 0), "rfb_window_aggr_dev");
    const int vk = rfb_kind_of(val_type);
    if (!(vk == K_I64 || vk == K_F64) || val_type == RFB_SYMBOL || !(op == RFB_A_SUM || op == RFB_A_MIN || op == RFB_A_MAX || op == RFB_A_COUNT || op == RFB_A_AVG)) {
        rfb_set_error("window aggregate %d over value type %d (sum / min / max / count / avg of I64-kind or F64 values)", op, val_type);
        return RFB_ERR_TYPE;
    }
    if (left_len == 0) return RFB_OK;
    KeyCols r;
    r.ncols = 0;
    if (vk == K_F64) return window_launch<f64>(ctx, op, r, 0, right_time, val, first, last, win_lo, win_hi, left_len, jtype, out);
    return window_launch<i64>(ctx, op, r, 0, right_time, val, first, last, win_lo, win_hi, left_len, jtype, out);
}
