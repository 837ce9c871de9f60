/*
 * rfb200.h — C ABI of librfb200.so: the B200 (sm_100a) implementation of RayforceDB's vectorised columnar
 * execution hot path.  Plain pointers and sizes only; no C++/torch types.  Host side stays pure C.
 *
 * Two layers live in this header:
 *   1. device layer  (rfb_*_dev)  — operands are DEVICE pointers; kernels are enqueued on the context's stream.
 *                                   This is what a host that keeps columns resident in HBM calls, and what
 *                                   bench.py times as `value`.
 *   2. host layer    (rfb_*_host) — operands are HOST pointers (column payloads, i.e. (char*)obj + 16 of a
 *                                   reference obj_t); the call pins, ships the column to HBM with cudaMemcpyAsync
 *                                   in a chunked pipeline overlapped with the kernels, and returns host results.
 *                                   bench.py times this as `e2e`.
 * The obj_t-level operator surface (ray_lt, ray_where, ray_sum, index_group, aggr_sum, ray_sort_asc, ...) that the
 * reference's evaluator binds is in rfb200_ops.h and is implemented in pure C on top of this header.
 *
 * Element type codes are the reference's (core/rayforce.h:50-62); nulls are its in-band sentinels
 * (core/rayforce.h:97-100).  Each entry point cites the reference function whose loop it replaces
 * (paths inside the reference tree, commit 2151d51d).
 *
 * Error convention: every function returns RFB_OK (0) or a negative rfb_status; rfb_last_error() gives the text.
 * RFB_ERR_TYPE / RFB_ERR_LENGTH correspond to the reference's err_type / err_length (core/error.h:86-97) so the
 * operator layer can raise the same Rayfall errors.  There is NO CPU fallback: without a CUDA device every compute
 * entry point fails with RFB_ERR_CUDA.
 */
#ifndef RFB200_H
#define RFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFB_ABI_VERSION 1

typedef enum {
    RFB_OK = 0,
    RFB_ERR_TYPE = -1,   /* unsupported element type combination (reference: err_type) */
    RFB_ERR_LENGTH = -2, /* vector lengths differ (reference: err_length) */
    RFB_ERR_CUDA = -3,   /* CUDA runtime failure, or no device */
    RFB_ERR_ARG = -4,    /* bad argument (NULL pointer, negative size, unknown op) */
    RFB_ERR_NOMEM = -5   /* device or pinned allocation failed (reference: err_limit) */
} rfb_status;

/* element types == reference TYPE_* */
enum { RFB_B8 = 1, RFB_U8 = 2, RFB_I16 = 3, RFB_I32 = 4, RFB_I64 = 5, RFB_SYMBOL = 6, RFB_DATE = 7, RFB_TIME = 8,
       RFB_TIMESTAMP = 9, RFB_F64 = 10 };

/* comparison operators: ray_eq/ne/lt/gt/le/ge (core/cmp.c:692-697) */
enum { RFB_EQ = 0, RFB_NE = 1, RFB_LT = 2, RFB_GT = 3, RFB_LE = 4, RFB_GE = 5 };

/* which folds a reduction kernel computes (bit set).  ray_sum/min/max/cnt (core/math.c:1785-2045) */
enum { RFB_F_SUM = 1, RFB_F_CNT = 2, RFB_F_MIN = 4, RFB_F_MAX = 8, RFB_F_ROWS = 16, RFB_F_ALL = 31 };

/* element-wise arithmetic: ray_add/sub/mul/div/fdiv/mod/xbar (core/math.c:2436-2442) */
enum { RFB_ADD = 0, RFB_SUB = 1, RFB_MUL = 2, RFB_DIV = 3, RFB_FDIV = 4, RFB_MOD = 5, RFB_XBAR = 6 };
enum { RFB_ROUND = 0, RFB_FLOOR = 1, RFB_CEIL = 2 };

/* grouped aggregates: aggr_sum/min/max/count/avg (core/aggr.c:1078-1453, 2013-2133) */
enum { RFB_A_SUM = 0, RFB_A_MIN = 1, RFB_A_MAX = 2, RFB_A_COUNT = 3, RFB_A_AVG = 4, RFB_A_MED = 5, RFB_A_DEV = 6,
       RFB_A_FIRST = 7, RFB_A_LAST = 8 /* aggr_first / aggr_last (core/aggr.c:441-577, 851-1075) */ };

/* group index kinds (core/index.h:31-36) */
enum { RFB_INDEX_IDS = 0, RFB_INDEX_SHIFT = 1 };

#define RFB_NULL_I16 ((int16_t)0x8000)
#define RFB_NULL_I32 ((int32_t)0x80000000)
#define RFB_NULL_I64 ((int64_t)0x8000000000000000LL)
#define RFB_INF_I64 ((int64_t)0x7FFFFFFFFFFFFFFFLL)
#define RFB_INDEX_SCOPE_LIMIT (4096 * 128) /* core/index.h:29 */

/* A scalar operand ("atom").  `type` is an element type code; the value sits in the matching member. */
typedef struct {
    int32_t type;
    int32_t _pad;
    union {
        int64_t i64;
        double f64;
        int32_t i32;
        int16_t i16;
        uint8_t u8;
    } v;
} rfb_scalar_t;

/* Result of a fold over one column (all requested folds at once).
 * sum: the null-skipping sum in the reference's accumulator type for the column (core/math.c:1850-1871):
 *      U8/I16/I64 -> sum_i64 (wraps mod 2^64); I32/TIME -> sum_i64 holds the value wrapped to 32 bits, sign-extended;
 *      F64 -> sum_f64 (NaN skipped).  nonnull = non-null elements folded (after the filter) — what ray_cnt / ray_avg use.
 *      rows = elements selected by the filter, nulls included; exact whenever RFB_F_ROWS is requested (and always
 *      without a filter); -1 when it was not requested and the kernel took the path that never looks at nulls.
 * min/max: null-skipping; typed null when nothing was folded (MINI64(NULL,y)=y, core/ops.h:185).
 *      integers in min_i64/max_i64 (sign-extended), F64 in min_f64/max_f64. */
typedef struct {
    int64_t rows;
    int64_t nonnull;
    int64_t sum_i64;
    double sum_f64;
    int64_t min_i64, max_i64;
    double min_f64, max_f64;
    double sum_f64_err; /* F64 sums are accumulated error-free (TwoSum) as hi + lo; sum_f64 is hi + lo rounded once and this
                           is the residual it dropped, so partial results can be merged without losing the 1-ULP property */
} rfb_fold_t;

typedef struct rfb_ctx rfb_ctx_t; /* opaque: device, stream, scratch, pinned staging */

/* ------------------------------------------------------------------ context / memory / transfers */
int rfb_abi_version(void);
const char *rfb_last_error(void);               /* thread-local, never NULL */
int rfb_device_count(void);                     /* 0 without a usable CUDA device; never fails */
int rfb_ctx_create(int device, rfb_ctx_t **out);
void rfb_ctx_destroy(rfb_ctx_t *ctx);
int rfb_ctx_set_stream(rfb_ctx_t *ctx, void *cuda_stream); /* adopt a caller-owned cudaStream_t (NULL = own) */
void *rfb_ctx_stream(rfb_ctx_t *ctx);
/* Redirect the rfb_fold_t written by the *_fold_dev kernels to caller-provided device-visible memory (NULL = back to
 * the context's own mapped pinned slot).  Used when the next consumer is on the device (e.g. an NCCL all-reduce of
 * per-GPU partial aggregates); calls must then pass out == NULL. */
int rfb_ctx_set_result_ptr(rfb_ctx_t *ctx, void *device_visible);
int rfb_ctx_sm_count(rfb_ctx_t *ctx);
int rfb_sync(rfb_ctx_t *ctx);
int64_t rfb_launch_count(rfb_ctx_t *ctx);       /* kernels launched through this context so far */
/* The tuning knobs (RFB_GROUP_STRATEGY, RFB_PART_MIN_ROWS, RFB_ACCUM_TMA) are read from the environment once, on first
 * use — never on the per-call path.  A host that changes them afterwards calls this to have them read again. */
void rfb_options_reload(void);

int rfb_dev_alloc(rfb_ctx_t *ctx, size_t bytes, void **dptr);
int rfb_dev_free(rfb_ctx_t *ctx, void *dptr);
int rfb_dev_memset(rfb_ctx_t *ctx, void *dptr, int byte, size_t bytes);
int rfb_dev_mem_info(rfb_ctx_t *ctx, size_t *free_bytes, size_t *total_bytes);   /* cudaMemGetInfo of the context's device */
int rfb_host_pin(void *p, size_t bytes);        /* cudaHostRegister: column payload stays where the host put it */
int rfb_host_unpin(void *p);
int rfb_host_alloc_pinned(size_t bytes, void **p);
int rfb_host_free_pinned(void *p);
/* column shipping: async on the context stream; `pinned` tells whether src/dst is page-locked */
int rfb_h2d(rfb_ctx_t *ctx, void *dst_dev, const void *src_host, size_t bytes);
int rfb_d2h(rfb_ctx_t *ctx, void *dst_host, const void *src_dev, size_t bytes);
/* synchronous device -> host copy through the driver only (no helper threads): returns with the bytes in place */
int rfb_d2h_sync_plain(rfb_ctx_t *ctx, void *dst_host, const void *src_dev, size_t bytes);

/* deterministic synthetic columns generated in HBM (bench / tests): x[i] = splitmix64(seed, i) % modulus (+ offset),
 * every `null_every`-th element (if > 0) replaced by the type's null.  type: I32, I64 or F64 (F64: value / scale). */
int rfb_fill_splitmix_dev(rfb_ctx_t *ctx, int type, void *x, int64_t n, uint64_t seed, uint64_t modulus,
                          int64_t offset, int64_t null_every, double f64_scale);

/* ------------------------------------------------------------------ device layer: scan / fold */

/* ray_sum / ray_min / ray_max / ray_cnt over one column (core/math.c:1785-2045; driver unop_fold :2176-2231).
 * folds: bit set of RFB_F_*.  Result is written to *out (host memory) after the stream is synchronised.
 * Every *_fold_dev call accepts out == NULL: the kernel is only enqueued and the result is collected later with
 * rfb_fold_result() (one outstanding fold per context). */
int rfb_fold_dev(rfb_ctx_t *ctx, int folds, int type, const void *x, int64_t n, rfb_fold_t *out);
int rfb_fold_result(rfb_ctx_t *ctx, rfb_fold_t *out);

/* Fused `select {(fold v) from t where (cmp p k)}`: predicate scan + selection + gather + fold in ONE pass
 * (replaces ray_lt -> ray_where -> filter_collect -> ray_sum: core/cmp.c:335, core/ops.c:255,
 * core/rayforce.c:1100, core/math.c:1874-1890).  pred/val may be the same column (8 B/row) or differ.
 * The mask, the id vector and the gathered column are never materialised. */
int rfb_filter_fold_dev(rfb_ctx_t *ctx, int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k, int folds,
                        int val_type, const void *val, int64_t n, rfb_fold_t *out);

/* Fused `select {(fold v) from t where (and|or (cmp p1 k1) (cmp p2 k2) ...)}`: up to 4 range predicates over 8-byte columns
 * (I64-kind or F64), combined with `and` (conjunction != 0) or `or`, tested in the same pass that folds `val`
 * (I64-kind or F64).  Replaces per-conjunct masks + and_op/or_op (core/logic.c:34-110) + ray_where + filter_collect +
 * the fold: (npred + 1) x 8 B per row instead.  rows = rows selected (always counted here). */
typedef struct {
    int32_t op;          /* RFB_EQ .. RFB_GE */
    int32_t type;        /* element type of col */
    const void *col;     /* device pointer */
    rfb_scalar_t k;
} rfb_pred_t;
int rfb_multi_filter_fold_dev(rfb_ctx_t *ctx, int npred, const rfb_pred_t *preds, int conjunction, int folds, int val_type,
                              const void *val, int64_t n, rfb_fold_t *out);

/* Fused `(fold (+ (* a b) c))` over three F64 columns with the reference's NaN propagation (MULF64/ADDF64,
 * core/ops.h:155,164) and NaN-skipping fold: replaces ray_mul -> ray_add -> ray_sum/ray_cnt. */
int rfb_fma_fold_dev(rfb_ctx_t *ctx, int folds, const double *a, const double *b, const double *c, int64_t n,
                     rfb_fold_t *out);

/* ------------------------------------------------------------------ device layer: operator-exact building blocks */

/* ray_eq..ray_ge -> cmp_map (core/cmp.c:335-683): mask[i] = OP(x[i], y[i]) as 0/1 bytes.
 * xn / yn: element count, or -1 when that side is the atom *xs / *ys.  Returns RFB_ERR_LENGTH if both are vectors
 * of different length. */
int rfb_cmp_dev(rfb_ctx_t *ctx, int op, int xt, const void *x, int64_t xn, const rfb_scalar_t *xs, int yt,
                const void *y, int64_t yn, const rfb_scalar_t *ys, uint8_t *mask);

/* and / or / not on B8 masks: and_op_partial / or_op_partial (core/logic.c:34-86; the right side may be one broadcast byte:
 * bn == -1, value bs) and ray_not (core/order.c:422-443; b ignored).  out may alias a (the reference updates in place). */
enum { RFB_M_AND = 0, RFB_M_OR = 1, RFB_M_NOT = 2 };
int rfb_mask_logic_dev(rfb_ctx_t *ctx, int op, const uint8_t *a, int64_t n, const uint8_t *b, int64_t bn, uint8_t bs,
                       uint8_t *out);

/* ray_where -> ops_where (core/ops.c:255-273): ascending row ids of the set bytes.  ids must have room for n.
 * *count (host) receives the number written. */
int rfb_where_dev(rfb_ctx_t *ctx, const uint8_t *mask, int64_t n, int64_t *ids, int64_t *count);

/* ray_where(ray_<cmp>(x, k)) without the mask: ids of rows where OP(x[i], k) */
int rfb_cmp_where_dev(rfb_ctx_t *ctx, int op, int type, const void *x, int64_t n, const rfb_scalar_t *k, int64_t *ids,
                      int64_t *count);

/* filter_collect -> at_ids (core/rayforce.c:1036-1159): out[i] = col[ids[i]] */
int rfb_gather_dev(rfb_ctx_t *ctx, int type, const void *col, const int64_t *ids, int64_t m, void *out);

/* ray_sum(MAPFILTER[col, ids]) without materialising the gathered column (core/math.c:1874-1890) */
int rfb_gather_fold_dev(rfb_ctx_t *ctx, int folds, int type, const void *col, const int64_t *ids, int64_t m,
                        rfb_fold_t *out);

/* ray_add..ray_xbar -> binop_map (core/math.c:2280-2345), vector or atom operands, over the reference's full type matrix
 * (ray_add_partial .. ray_xbar_partial, core/math.c:251-1782): B8 / U8 / I16 / I32 / I64 / DATE / TIME / TIMESTAMP / F64 operands
 * incl. the unit conversions (DATE + TIME -> TIMESTAMP ...).  rfb_binop_type answers the result element type for I32/I64/F64
 * operands (infer_math_type & co, core/math.c:92-249) or RFB_ERR_TYPE; rfb_binop_type_form answers it for any operand types and
 * one operand form (0 vector-vector, 1 vector-atom, 2 atom-vector) — RFB_ERR_TYPE exactly where the reference has no case. */
int rfb_binop_type(int op, int xt, int yt);
int rfb_binop_type_form(int op, int form, int xt, int yt);
int rfb_binop_dev(rfb_ctx_t *ctx, int op, int xt, const void *x, int64_t xn, const rfb_scalar_t *xs, int yt,
                  const void *y, int64_t yn, const rfb_scalar_t *ys, void *out);
/* ray_round / ray_floor / ray_ceil -> unop_map (core/math.c:2047-2117, 2233-2278) */
int rfb_unop_f64_dev(rfb_ctx_t *ctx, int op, const double *x, int64_t n, double *out);

/* ------------------------------------------------------------------ device layer: group-by */

/* index_group -> index_group_i64 (core/index.c:2094-2106): scope (min/max) + first-occurrence group numbering.
 * keys: I64 column; filter: row ids or NULL; len: number of (filtered) rows.
 * Outputs (device): group_ids[len] row-position -> gid; first_ids[len] (first `groups` valid) position of each
 * group's first row.  Dense path when range <= len (perfect hash, core/index.c:2013-2055), else open-addressing
 * hash (core/index.c:1777-1911).  Both number groups by first occurrence (reference order at -c 1 / dense). */
typedef struct {
    int32_t index_type; /* RFB_INDEX_SHIFT if dense and range <= RFB_INDEX_SCOPE_LIMIT else RFB_INDEX_IDS */
    int32_t dense;
    int64_t groups;
    int64_t min, max, range;
} rfb_group_info_t;
int rfb_group_i64_dev(rfb_ctx_t *ctx, const int64_t *keys, const int64_t *filter, int64_t len, int64_t *group_ids,
                      int64_t *first_ids, rfb_group_info_t *info);

/* index_group_list, perfect-hash key-fusion path (core/index.c:2308-2424): groups rows by the TUPLE of `ncols` (<= 8) I64-kind
 * key columns, numbered by first occurrence like the single-key index.  Every column's scope is taken, the tuple is fused
 * into one key sum_c (col_c - min_c) * stride_c and grouped by rfb_group_i64_dev.  When the product of the key
 * ranges does not fit 62 bits the tuples are grouped by row hash instead (the reference hashes rows and radix-partitions,
 * core/index.c:2556-2729; the device probes one open-addressing table of representative rows): info->dense = 0, min/max unset.
 * Outputs as rfb_group_i64_dev (info->min/max/range describe the fused key). */
int rfb_group_keys_i64_dev(rfb_ctx_t *ctx, int ncols, const int64_t *const *cols, const int64_t *filter, int64_t len,
                           int64_t *group_ids, int64_t *first_ids, rfb_group_info_t *info);

/* aggr_sum/min/max/count/avg (core/aggr.c AGGR_ITER :73-161): out[gid] (+)= val[row]; sticky-null sum, +INF-init
 * min, NULL-init max, row count, f64 avg.  out: `groups` elements of rfb_aggr_type(op, val_type). */
int rfb_aggr_type(int op, int val_type);
/* op RFB_A_MED = aggr_med (core/aggr.c:2136-2246): median of each group's values in ray_asc order (nulls / NaN sort first and
 * count as values; I64/TIMESTAMP/F64 values, any other type gives all-null like the reference's default branch).
 * op RFB_A_DEV = aggr_dev (core/aggr.c:2250-2906): population standard deviation of the non-null values, f64 sums of x and
 * x*x per group, sqrt(max(sumsq/n - mean^2, 0)); 0 rows -> null, 1 row -> 0. */
/* op RFB_A_FIRST = aggr_first: the value at each group's first row (null or not: the reference's first_ids fast path);
 * op RFB_A_LAST = aggr_last: each group's last NON-NULL value, null if none (the reference's single-chunk result; above its
 * parallel threshold its own answer depends on the thread count, DESIGN.md Q18: rfb_aggr_last_dev takes the chunk count). */
int rfb_aggr_dev(rfb_ctx_t *ctx, int op, int val_type, const void *val, const int64_t *filter,
                 const int64_t *group_ids, int64_t len, int64_t groups, void *out);
/* aggr_last exactly as the reference computes it on `nchunks` worker chunks (core/aggr.c:262-295 aggr_map, AGGR_COLLECT with
 * `if (out == null) out = in`): per group the last non-null value inside the FIRST chunk of len / nchunks rows that has one.
 * nchunks = the reference's pool_split_by_mem(len, groups, width) (core/pool.c:450-478); 1 = the group's last non-null value. */
int rfb_aggr_last_dev(rfb_ctx_t *ctx, int val_type, const void *val, const int64_t *filter, const int64_t *group_ids,
                      int64_t len, int64_t groups, int64_t nchunks, void *out);

/* aggr_row / aggr_collect (core/aggr.c:3021-3136): the rows of every group.  out_rows[len] = row ids (filter[i], or i without
 * a filter) ordered by group id and, inside a group, by position (the order AGGR_ITER pushes them); offsets[groups+1] = where
 * each group starts.  aggr_row's list g is out_rows[offsets[g] .. offsets[g+1]); aggr_collect's is rfb_gather_dev of it. */
int rfb_group_rows_dev(rfb_ctx_t *ctx, const int64_t *group_ids, const int64_t *filter, int64_t len, int64_t groups,
                       int64_t *out_rows, int64_t *offsets);

/* ray_med (core/math.c:2529-2626): median of a U8 / I16 / I64 vector -> *out (host).  Like the reference it sorts the whole
 * column (nulls first) but takes the middle of the NON-NULL count, and adds the two middle elements as integers.
 * ray_dev (core/math.c:2628-2700): sqrt(sum((x - mean)^2) / n) over the non-null values, two passes. */
int rfb_med_dev(rfb_ctx_t *ctx, int type, const void *x, int64_t n, double *out);
int rfb_stddev_dev(rfb_ctx_t *ctx, int type, const void *x, int64_t n, double *out);

/* Fused `select {s: (sum v) c: (count v) from t by k [where (cmp p kk)]}` on a dense key domain: one scope pass +
 * one accumulate pass; never materialises group_ids.  keys: I64 or I32 (I32 is a superset of the reference, Q1).
 * Outputs (device arrays sized `max_groups`): group keys in first-occurrence order, sums (I64, sticky null), counts.
 * *groups (host) = number of groups.  pred may be NULL (no filter). */
int rfb_group_sum_count_dev(rfb_ctx_t *ctx, int key_type, const void *keys, const int64_t *val, int64_t n,
                            int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k, int64_t max_groups,
                            int64_t *out_keys, int64_t *out_sums, int64_t *out_counts, int64_t *groups);

/* ------------------------------------------------------------------ device layer: equi-join row matching */

/* ray_find -> index_find_i64 (core/index.c:1507-1574) for ncols == 1, index_left_join_obj (core/index.c:2886-2928) for a
 * tuple of ncols (<= 8) I64-kind key columns: ids[i] = the FIRST build row whose key (tuple) equals probe row i's, else
 * NULL_I64.  Keys compare by bit pattern (a null is a key like any other, core/index.c:59-105).  This is the left join's
 * row index: column c of the joined table is rfb_gather_dev(right column c, ids). */
int rfb_find_rows_dev(rfb_ctx_t *ctx, int ncols, const int64_t *const *build_cols, int64_t build_len,
                      const int64_t *const *probe_cols, int64_t probe_len, int64_t *ids);

/* index_inner_join_obj (core/index.c:2930-3000): the matching pairs, probe rows ascending: probe_ids[j], build_ids[j]
 * (the first build row of probe row probe_ids[j]'s key); both sized probe_len; *count (host) = number of pairs. */
int rfb_inner_join_dev(rfb_ctx_t *ctx, int ncols, const int64_t *const *build_cols, int64_t build_len,
                       const int64_t *const *probe_cols, int64_t probe_len, int64_t *probe_ids, int64_t *build_ids,
                       int64_t *count);

/* index_asof_join_obj (core/index.c:3194-3268): ids[i] = the LAST build row with probe row i's key tuple whose time is <= probe
 * row i's time, else NULL_I64.  The search is the reference's binary search over the key's build rows in row order, so like
 * there the build rows of one key must be ordered by time.  time_type: I32/DATE/TIME or I64/TIMESTAMP, same on both sides. */
int rfb_asof_join_dev(rfb_ctx_t *ctx, int ncols, const int64_t *const *build_cols, int time_type, const void *build_time,
                      int64_t build_len, const int64_t *const *probe_cols, const void *probe_time, int64_t probe_len,
                      int64_t *ids);

/* ray_distinct -> index_distinct_i64 (core/index.c:551-607), its direct-addressing branch (key range <= len or <= 2^20): the
 * distinct keys in ASCENDING order; out must hold min(n, range) entries, *count (host) = how many.  RFB_ERR_ARG when the
 * range is not dense: the reference's hash branch emits the keys in the slot order of its own table, which a parallel
 * build cannot reproduce — callers keep that case on the CPU body. */
int rfb_distinct_i64_dev(rfb_ctx_t *ctx, const int64_t *keys, int64_t n, int64_t *out, int64_t *count);

/* Window join aggregate (core/join.c:358-485, index_window_join_obj core/index.c:3287-3346, AGGR_ITER's WINDOW branch
 * core/aggr.c:131-160).  The right table must be ordered by (key tuple, time) — ray_window_join sorts it itself — so that a key's
 * rows form one block.  For left row i: the rows of its key's block inside the window [win_lo[i], win_hi[i]] on the 4-byte time
 * column (jtype 0 = window-join: from the last row at or before win_lo; jtype 1 = window-join1: from the first row at or after
 * it) are folded with the GROUPED aggregate: op RFB_A_SUM (sticky null) / MIN / MAX over I64-kind or F64 values -> out[i] of
 * the value type, RFB_A_COUNT -> I64, RFB_A_AVG -> F64 (sum of the non-null values in row order / their count); no block or
 * no row in the window -> null (count: 0). */
int rfb_window_join_dev(rfb_ctx_t *ctx, int ncols, const int64_t *const *right_cols, const int32_t *right_time, int64_t right_len,
                        const int64_t *const *left_cols, int64_t left_len, const int32_t *win_lo, const int32_t *win_hi, int jtype,
                        int op, int val_type, const void *val, void *out);

/* The aggregate alone, for an index that already holds every left row's block [first[i], last[i]] of the sorted right table
 * (first[i] == NULL_I64: the key has no block): what aggr_* receive from the reference's own index_window_join_obj. */
int rfb_window_aggr_dev(rfb_ctx_t *ctx, const int32_t *right_time, const int64_t *first, const int64_t *last, int64_t left_len,
                        const int32_t *win_lo, const int32_t *win_hi, int jtype, int op, int val_type, const void *val, void *out);

/* ------------------------------------------------------------------ device layer: sort */

/* ray_sort_asc / ray_sort_desc (core/sort.c:430-479, 691-740): stable permutation (I64 row ids).
 * Types: U8/B8, I16, I32/DATE/TIME, I64/TIMESTAMP, F64 (NaN first ascending, -0.0 < +0.0). */
int rfb_sort_dev(rfb_ctx_t *ctx, int type, const void *x, int64_t n, int descending, int64_t *perm);

/* ------------------------------------------------------------------ host layer (e2e): HOST pointers in, host results out */

/* Same contract as rfb_filter_fold_dev but pred/val are HOST column payloads.  The column is shipped in chunks of
 * `chunk_rows` (0 = default) through pinned staging (or directly if the caller pinned it with rfb_host_pin), each
 * chunk's kernel overlapping the next chunk's copy.  h2d_bytes (optional) receives the bytes copied. */
int rfb_filter_fold_host(rfb_ctx_t *ctx, int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k, int folds,
                         int val_type, const void *val, int64_t n, int64_t chunk_rows, rfb_fold_t *out,
                         int64_t *h2d_bytes);
int rfb_fold_host(rfb_ctx_t *ctx, int folds, int type, const void *x, int64_t n, int64_t chunk_rows, rfb_fold_t *out,
                  int64_t *h2d_bytes);
/* ------------------------------------------------------------------ one process per GPU: the final merge over NVLink peer memory
 *
 * the only exchange of the path is the final merge of the per-GPU partial aggregates.  For ungrouped folds that is a
 * 72-byte record per GPU; an NCCL all-reduce + device-to-host copy + host synchronisation costs more than the merge is worth
 * (~80 us next to a 1.1 ms scan).  Instead every rank owns a small mailbox in its HBM, exported to the peers with CUDA IPC:
 *     rfb_peer_mailbox_create(ctx, handle)         allocate it, return its 64-byte IPC handle (exchange the handles of all ranks
 *                                                  by any means: torch.distributed all_gather, MPI, a file)
 *     rfb_peer_mailbox_bind(ctx, rank, world, h)   h = world x 64 bytes, the handles in rank order: maps every peer's mailbox
 *     rfb_fold_allreduce_peers(ctx, type, out)     after an asynchronous fold launch (out == NULL form of rfb_fold_dev /
 *                                                  rfb_filter_fold_dev / ...) on this context: ONE 32-thread kernel pushes this
 *                                                  rank's partial into every peer's mailbox (P2P stores over NVLink + a sequence
 *                                                  flag), waits for the peers' partials, folds them in rank order and reports
 *                                                  the merged rfb_fold_t where a single-GPU fold reports.  Every rank must call
 *                                                  it the same number of times.  Bit-identical on all ranks.
 *                                                  out == NULL: launch only — queries can be queued back to back (fold, exchange,
 *                                                  fold, exchange ... on the context's stream) without a host synchronisation;
 *     rfb_fold_peers_result(ctx, out)              drains the stream and reports the merged result of the last exchange. */
int rfb_peer_mailbox_create(rfb_ctx_t *ctx, void *ipc_handle_64);
/* The exchange step of a group-by sharded by row range (`select {(sum v) (count v)} by k`), over NVLink peer memory:
 *     rfb_peer_groups_create(ctx, capacity, h)     this rank's exchange buffer (two halves of capacity (key, sum, count) rows) and
 *                                                  its CUDA IPC handle; rfb_peer_groups_bind maps every peer's (handles in rank order)
 *     rfb_group_merge_peers(ctx, keys, sums, counts, n_local, out_keys, out_sums, out_counts, max_groups, &groups)
 *                                                  keys / sums / counts: this rank's result of rfb_group_sum_count_dev (device, local
 *                                                  first-occurrence order).  Publishes them, meets the peers (sequence flags), folds
 *                                                  all ranks' lists read IN PLACE over NVLink into direct-address tables (wrapping
 *                                                  sums with the sticky null of ADDI64, counts, first position) and emits the groups
 *                                                  in GLOBAL first-occurrence order (rank r holds rows before rank r + 1).  Every rank
 *                                                  must call it the same number of times and ends with the same lists.  Dense key
 *                                                  domains (kmax - kmin < 2^24); RFB_ERR_TYPE for a wider one — the caller then
 *                                                  gathers the lists and re-groups them (rfb_group_i64_dev + rfb_aggr_dev). */
int rfb_peer_groups_create(rfb_ctx_t *ctx, int64_t capacity, void *ipc_handle_64);
int rfb_peer_groups_bind(rfb_ctx_t *ctx, int rank, int world, const void *handles);
int rfb_group_merge_peers(rfb_ctx_t *ctx, const int64_t *keys, const int64_t *sums, const int64_t *counts, int64_t n_local,
                          int64_t *out_keys, int64_t *out_sums, int64_t *out_counts, int64_t max_groups, int64_t *groups);
int rfb_peer_mailbox_bind(rfb_ctx_t *ctx, int rank, int world, const void *handles);
int rfb_fold_allreduce_peers(rfb_ctx_t *ctx, int val_type, rfb_fold_t *out);
int rfb_fold_peers_result(rfb_ctx_t *ctx, rfb_fold_t *out);

/* The fused multi-column queries over HOST columns (configs 3-5 end to end): the columns are shipped whole, then
 * rfb_group_sum_count_dev / rfb_fma_fold_dev run; group lists come back into HOST arrays of max_groups entries. */
int rfb_group_sum_count_host(rfb_ctx_t *ctx, int key_type, const void *keys, const int64_t *val, int64_t n, int cmp_op, int pred_type,
                             const void *pred, const rfb_scalar_t *k, int64_t max_groups, int64_t *out_keys, int64_t *out_sums,
                             int64_t *out_counts, int64_t *groups, int64_t *h2d_bytes);
int rfb_fma_fold_host(rfb_ctx_t *ctx, int folds, const double *a, const double *b, const double *c, int64_t n, rfb_fold_t *out,
                      int64_t *h2d_bytes);

/* ------------------------------------------------------------------ every visible GPU from one host process
 *
 * The reference's pool_split_by / pool_chunk_aligned (core/pool.c:450-507) hand row ranges to worker threads and merge their partials
 * (core/math.c:2222-2228).  rfb_mgpu_* does the same with GPUs as the workers, from pure C: device g takes rows [g*n/N, (g+1)*n/N)
 * of a HOST column over its own PCIe link (one host thread per device drives the chunked copy + fused kernel pipeline of the
 * host layer) and the N partials are merged on the host — folds: N rfb_fold_t records; group-by: the N (key, sum, count) lists
 * merged by key in row order, so the groups keep their first-occurrence numbering.  Results equal the single-GPU ones bit for bit
 * (fp64 sums: within 1 ULP of the exact sum, as everywhere). */
#define RFB_MGPU_MAX 16
typedef struct rfb_mgpu rfb_mgpu_t;
int rfb_mgpu_create(int ndev /* <= 0: every visible device */, rfb_mgpu_t **out);
void rfb_mgpu_destroy(rfb_mgpu_t *m);
int rfb_mgpu_devices(const rfb_mgpu_t *m);
rfb_ctx_t *rfb_mgpu_ctx(rfb_mgpu_t *m, int g);   /* device g's context (borrowed) */
/* rfb_filter_fold_host over all devices; pred == NULL: plain fold (rfb_fold_host) */
int rfb_mgpu_filter_fold_host(rfb_mgpu_t *m, int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k, int folds,
                              int val_type, const void *val, int64_t n, int64_t chunk_rows, rfb_fold_t *out, int64_t *h2d_bytes);
/* rfb_group_sum_count_dev over HOST columns on all devices; outputs are HOST arrays of max_groups entries */
int rfb_mgpu_group_sum_count_host(rfb_mgpu_t *m, int key_type, const void *keys, const int64_t *val, int64_t n, int cmp_op,
                                  int pred_type, const void *pred, const rfb_scalar_t *k, int64_t max_groups, int64_t *out_keys,
                                  int64_t *out_sums, int64_t *out_counts, int64_t *groups, int64_t *h2d_bytes);

/* ------------------------------------------------------------------ column files: the reference's on-disk column format
 * 16-byte object header (mmod 0xfd, type, attrs, len) + raw payload (written by `set`, core/binary.c:264-307; mapped back by
 * `get`, core/unary.c:60-133).  A splayed table is a directory of these, a parted table a directory per partition.
 * rfb_column_file_open maps one read-only; `payload` is then a host column for every *_host entry point. */
typedef struct {
    int32_t type;
    int32_t attrs;
    int64_t len;
    const void *payload;
    void *map_base;   /* private: the mapping */
    size_t map_bytes;
} rfb_column_file_t;
int rfb_column_file_open(const char *path, rfb_column_file_t *out);
int rfb_column_file_close(rfb_column_file_t *f);
int rfb_column_file_write(const char *path, int type, int attrs, const void *payload, int64_t len);

#ifdef __cplusplus
}
#endif
#endif /* RFB200_H */
