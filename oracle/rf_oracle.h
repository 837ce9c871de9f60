/*
 * rf_oracle.h — CPU restatement of the RayforceDB vectorised execution hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (rayforce_b200/, include/) may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Every function is a scalar, single-threaded, strictly left-to-right restatement of the algorithm at the cited
 * reference location (paths relative to the reference tree, commit 2151d51d).  It is pinned two ways
 * (tests/test_oracle_*.py): against golden vectors transcribed from the reference's own tests (tests/golden/) and
 * against the reference itself compiled from source (oracle/_ref/librayforce_ref.so, see oracle/Makefile).
 *
 * Conventions: columns are plain contiguous little-endian arrays; nulls are in-band sentinels
 * (core/rayforce.h:97-100); element types use the reference's type codes (core/rayforce.h:50-62).
 * A length of RFO_ATOM (-1) marks a scalar operand ("atom") that is broadcast.
 */
#ifndef RF_ORACLE_H
#define RF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element type codes == reference TYPE_* (core/rayforce.h:50-62) */
enum { RFO_B8 = 1, RFO_U8 = 2, RFO_I16 = 3, RFO_I32 = 4, RFO_I64 = 5, RFO_SYMBOL = 6, RFO_DATE = 7, RFO_TIME = 8,
       RFO_TIMESTAMP = 9, RFO_F64 = 10 };

enum { RFO_EQ = 0, RFO_NE = 1, RFO_LT = 2, RFO_GT = 3, RFO_LE = 4, RFO_GE = 5 };           /* core/cmp.c:692-697 */
enum { RFO_SUM = 0, RFO_MIN = 1, RFO_MAX = 2, RFO_CNT = 3, RFO_AVG = 4, RFO_COUNT = 5, RFO_MED = 6, RFO_DEV = 7, RFO_FIRST = 8, RFO_LAST = 9 };   /* core/math.c:2388-2445, :2529-2700 */
enum { RFO_ADD = 0, RFO_SUB = 1, RFO_MUL = 2, RFO_DIV = 3, RFO_FDIV = 4, RFO_MOD = 5, RFO_XBAR = 6 };    /* core/math.c:2436-2441 */
enum { RFO_ROUND = 0, RFO_FLOOR = 1, RFO_CEIL = 2 };                                       /* core/math.c:2430-2432 */
enum { RFO_INDEX_IDS = 0, RFO_INDEX_SHIFT = 1 };                                            /* core/index.h:31-36 */

#define RFO_ATOM (-1)
#define RFO_OK 0
#define RFO_ERR_TYPE (-1)   /* reference: err_type(...)   */
#define RFO_ERR_LENGTH (-2) /* reference: err_length(...) */

#define RFO_NULL_I16 ((int16_t)0x8000)
#define RFO_NULL_I32 ((int32_t)0x80000000)
#define RFO_NULL_I64 ((int64_t)0x8000000000000000LL)
#define RFO_INF_I32 ((int32_t)0x7FFFFFFF)
#define RFO_INF_I64 ((int64_t)0x7FFFFFFFFFFFFFFFLL)
#define RFO_INDEX_SCOPE_LIMIT (4096 * 128) /* core/index.h:29 */

int rfo_type_size(int type);

/* ---- predicate scan: core/cmp.c:35-68 loops, :77-332 type matrix, :335-683 driver ----
 * out[i] = OP(x[i], y[i]) as 0/1 bytes; either side may be an atom (len RFO_ATOM).  Returns result length or <0. */
int64_t rfo_cmp(int op, int xt, const void *x, int64_t xn, int yt, const void *y, int64_t yn, uint8_t *out);

/* ---- selection vector: core/ops.c:255-273 ---- ids must hold n entries; returns the count written */
int64_t rfo_where(const uint8_t *mask, int64_t n, int64_t *ids);

/* ---- gather: core/rayforce.c:1036-1098 ---- out[i] = col[ids[i]] (element size from type) */
int rfo_at_ids(int type, const void *col, const int64_t *ids, int64_t m, void *out);

/* ---- ungrouped folds: core/math.c:1785-2045 (+ ray_avg :2445-2526, ops_count core/ops.c:169) ----
 * Writes the result atom into out (8 bytes, zero-padded) and its element type into *out_type. */
int rfo_fold(int op, int type, const void *x, int64_t n, void *out, int *out_type);

/* exact (128-bit float accumulated) sum of the non-NaN entries, rounded once; used for the fp64 tolerance rule */
double rfo_sum_f64_exact(const double *x, int64_t n);

/* ---- element-wise arithmetic: core/math.c:55-90 loops, :92-249 result types, :251-1782 matrix ----
 * Supported element types: I32, I64, F64 on either side (vector or atom).  *out_type receives the result type;
 * out must hold max(xn,yn,1) elements of it.  rfo_binop_type alone answers "what type would it be". */
int rfo_binop_type(int op, int xt, int yt);
/* the full type matrix of core/math.c:251-1782 (U8 / I16 / B8 / DATE / TIME / TIMESTAMP operands too): result vector type for
 * form 0 vector-vector, 1 vector-atom, 2 atom-vector, or RFO_ERR_TYPE when the reference has no such case. */
int rfo_binop_form(int op, int form, int xt, int yt);
int64_t rfo_binop(int op, int xt, const void *x, int64_t xn, int yt, const void *y, int64_t yn, void *out,
                  int *out_type);
int rfo_unop_f64(int op, const double *x, int64_t n, double *out); /* round/floor/ceil core/math.c:2047-2117 */

/* ---- group index: core/index.c:402-435 (scope), :2002-2092 (perfect hash), :1777-1911 + core/hash.c (hash) ----
 * keys: I64 column; filter: row ids or NULL; len: number of (filtered) rows.
 * Results (caller-allocated): first_ids[len] (only [0,groups) written), group_ids[len] row->gid (always written,
 * also on the SHIFT path, for the test's convenience).  hk (optional, may be NULL) must hold `range` entries when
 * the path is dense and receives slot->gid (NULL_I64 = absent).  Group numbering = first occurrence in (filtered)
 * row order (both paths, reference at -c 1). */
typedef struct {
    int index_type; /* RFO_INDEX_SHIFT when dense and range <= RFO_INDEX_SCOPE_LIMIT, else RFO_INDEX_IDS */
    int dense;      /* 1: perfect-hash path (range <= len); 0: open-addressing hash path */
    int64_t groups;
    int64_t min, max, range;
} rfo_group_info_t;
int rfo_group_i64(const int64_t *keys, const int64_t *filter, int64_t len, int64_t *group_ids, int64_t *first_ids,
                  int64_t *hk, rfo_group_info_t *info);

/* multi-key grouping (core/index.c:2731-2793): first-occurrence numbering of key tuples */
int rfo_group_multi(int ncols, const int64_t *const *cols, const int64_t *filter, int64_t len, int64_t *group_ids,
                    int64_t *first_ids, int64_t *groups_out);

/* ---- grouped aggregates: core/aggr.c:73-161 (AGGR_ITER), :1078-1453, :1455-2133 ----
 * val: column of val_type indexed by row; filter: row ids or NULL; group_ids[i] for i in [0,len).
 * out: `groups` entries of *out_type (sum/min/max: val type; count: I64; avg: F64). */
int rfo_aggr(int op, int val_type, const void *val, const int64_t *filter, const int64_t *group_ids, int64_t len,
             int64_t groups, void *out, int *out_type);

/* rfo_aggr also takes RFO_MED (aggr_med core/aggr.c:2136-2246) and RFO_DEV (aggr_dev :2250-2906): F64 per group;
 * RFO_FIRST (aggr_first :441-577, first_ids fast path: the value at the group's first row) and RFO_LAST (= rfo_aggr_last, 1 chunk) */
/* aggr_last (core/aggr.c:851-1075) as the reference computes it on `nchunks` worker chunks: the last non-null value of the group
 * inside the first chunk that has one */
int rfo_aggr_last(int val_type, const void *val, const int64_t *filter, const int64_t *group_ids, int64_t len, int64_t groups,
                  int64_t nchunks, void *out, int *out_type);

/* parted aggregates without a filter: PARTED_MAP (core/aggr.c:183-260), aggr_avg (:2065-2127).  combine: one result over all
 * partitions (groups == 1) vs one per partition. */
int rfo_parted_aggr(int op, int val_type, int nparts, const void *const *parts, const int64_t *lens, int combine, void *out, int *out_type);

/* aggr_row / aggr_collect (core/aggr.c:3021-3136): rows[len] = row ids grouped by gid in push order, offsets[groups+1] */
int rfo_group_rows(const int64_t *gid, const int64_t *filter, int64_t len, int64_t groups, int64_t *rows, int64_t *offsets);

/* ungrouped ray_med (core/math.c:2529-2626) and ray_dev (:2628-2700) */
int rfo_med(int type, const void *x, int64_t n, double *out);
int rfo_dev(int type, const void *x, int64_t n, double *out);

/* ---- equi-join row matching: ray_find / index_find_i64 (core/index.c:1507-1574), index_left_join_obj (:2886-2928),
 * index_inner_join_obj (:2930-3000).  ids[i] = first build row with probe row i's key tuple, else NULL_I64. */
int rfo_find_rows(int ncols, const int64_t *const *build, int64_t build_len, const int64_t *const *probe, int64_t probe_len, int64_t *ids);
int64_t rfo_inner_join(int ncols, const int64_t *const *build, int64_t build_len, const int64_t *const *probe, int64_t probe_len,
                       int64_t *probe_ids, int64_t *build_ids);

/* index_asof_join_obj (core/index.c:3194-3268): last build row of the probe row's key with time <= the probe time */
int rfo_asof_join(int ncols, const int64_t *const *build, int time_type, const void *build_time, int64_t build_len,
                  const int64_t *const *probe, const void *probe_time, int64_t probe_len, int64_t *ids);

/* ray_distinct -> index_distinct_i64, dense branch (core/index.c:551-577): ascending distinct keys; -1 = not dense */
int64_t rfo_distinct_i64(const int64_t *keys, int64_t n, int64_t *out);

/* window join (core/join.c:358-485, core/index.c:3287-3346, core/aggr.c:131-160): right table ordered by (key, time); per left row
 * the [first, last] block of its key, then the aggregate of the rows inside the window [wlo, whi] (jtype 0 = window-join,
 * 1 = window-join1).  Oracle only so far: the device path is a next step. */
int rfo_window_bounds(int ncols, const int64_t *const *right, int64_t rl, const int64_t *const *left, int64_t ll, int64_t *first, int64_t *last);
int rfo_window_aggr(int op, int val_type, const void *val, const int32_t *rtime, int64_t ll, const int64_t *first, const int64_t *last,
                    const int32_t *wlo, const int32_t *whi, int jtype, void *out, int *out_type);

/* ---- key sort: core/sort.c:183-428 asc, :481-689 desc ---- stable permutation, nulls/NaN first when ascending */
int rfo_sort(int type, const void *x, int64_t n, int descending, int64_t *perm);

/* ---- deterministic synthetic columns shared by tests / bench (not from the reference) ---- */
uint64_t rfo_splitmix64(uint64_t seed, uint64_t i);

#ifdef __cplusplus
}
#endif
#endif
