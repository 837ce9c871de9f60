/*
 * rfb200_ops.h — the reference-facing OPERATOR LAYER of rayforce-b200 (pure C, librfb200_ops.so).
 *
 * Every function here has the signature, argument meaning, ownership and error behaviour of the reference operator
 * of the same name (reference core/ops.h:202-204: unary_f = obj_p(obj_p), binary_f = obj_p(obj_p, obj_p)), prefixed
 * `rfb_` so that both can live in one process.  INTEGRATION.md shows how the reference binds them: its own objects are
 * linked with `-Wl,--wrap=<name>` and `__wrap_<name>` forwards to `rfb_<name>`, falling back to `__real_<name>` (the
 * reference's CPU body) whenever the GPU layer DECLINES an operand (returns NULL).
 *
 * Objects are the reference's `obj_t` (core/rayforce.h:112-133) byte for byte: 16-byte header {mmod, order, type, attrs,
 * rc:u32, union{atom value | len:i64}}, payload at +16, type codes core/rayforce.h:50-95, in-band nulls :97-100.
 * This layer never allocates or frees objects itself: it goes through the host's allocator (rfb_host_api_t), because
 * in the reference objects must come from the calling thread's heap.  A tiny malloc-based host
 * (rfb_ops_builtin_host) exists for standalone use and tests.
 *
 * Return convention of every operator below:
 *     new owned object            success (the caller drops it)
 *     host->err_type()/err_length()/...   the reference's ERR_OBJ for the same misuse ("type", "length", ...)
 *     NULL                        DECLINED: operand kind outside the GPU path (atoms only, LIST/TABLE/parted/MAPCOMMON/
 *                                 GUID/ENUM operands, fewer rows than RFB200_MIN_ROWS) — call the CPU body instead
 * Arguments are borrowed (never consumed), like the reference (eval.c:741-742).
 */
#ifndef RFB200_OPS_H
#define RFB200_OPS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the reference's object header (core/rayforce.h:112-133) */
typedef struct rfb_obj {
    uint8_t mmod;
    uint8_t order;
    int8_t type;   /* > 0 vector of that element type, < 0 atom, 0 LIST */
    uint8_t attrs;
    uint32_t rc;
    union {
        uint8_t u8;
        int16_t i16;
        int32_t i32;
        int64_t i64;
        double f64;
        struct rfb_obj *obj;
        int64_t len;  /* vectors / lists: element count; payload follows the header */
    };
} rfb_obj_t;
typedef rfb_obj_t *rfb_obj_p;

#define RFB_OBJ_PAYLOAD(o) ((void *)((char *)(o) + 16))
#define RFB_OBJ_LIST(o) ((rfb_obj_p *)RFB_OBJ_PAYLOAD(o))

/* type codes used by this layer (core/rayforce.h:50-95) */
enum { RFB_T_LIST = 0, RFB_T_B8 = 1, RFB_T_U8 = 2, RFB_T_I16 = 3, RFB_T_I32 = 4, RFB_T_I64 = 5, RFB_T_SYMBOL = 6,
       RFB_T_DATE = 7, RFB_T_TIME = 8, RFB_T_TIMESTAMP = 9, RFB_T_F64 = 10, RFB_T_MAPFILTER = 71, RFB_T_MAPGROUP = 72,
       RFB_T_PARTED = 77 /* + element type: a list of per-partition vectors (core/rayforce.h:70-82) */,
       RFB_T_NULL = 126, RFB_T_ERR = 127 };

/* What the host (the reference runtime, or the builtin one) provides.  Names follow the reference functions they are
 * bound to in INTEGRATION.md. */
typedef struct rfb_host_api {
    rfb_obj_p (*vector)(int8_t type, int64_t len); /* core/rayforce.c:237 vector(); type 0 = LIST */
    rfb_obj_p (*atom)(int8_t type);                /* core/rayforce.c:69 atom(): returns type = -type, value unset */
    rfb_obj_p (*clone_obj)(rfb_obj_p);             /* core/rayforce.c clone_obj */
    void (*drop_obj)(rfb_obj_p);                   /* core/rayforce.c drop_obj */
    rfb_obj_p (*err_type)(void);                   /* core/error.h:86 err_type(0,0,0,0) */
    rfb_obj_p (*err_length)(void);                 /* core/error.h:88 err_length(0,0,0,0,0,0) */
    rfb_obj_p (*err_limit)(void);                  /* core/error.h:92 err_limit(0): allocation / device failure */
    rfb_obj_p null_obj;                            /* &__NULL_OBJ (core/ops.c:33-36) */
    int64_t (*executors)(void);                    /* pool_get_executors_count(pool_get()) (core/pool.c:486); NULL = 1.  aggr_last's
                                                      answer depends on how many worker chunks the reference would use (Q18) */
} rfb_host_api_t;

/* Bind the layer to a host and a GPU.  Fails (non-zero) when no CUDA device is usable: there is no CPU fallback inside
 * this library — the caller keeps using its own CPU bodies. */
int rfb_ops_init(const rfb_host_api_t *host, int device);   /* device < 0: every visible GPU (rfb_mgpu_*): the one-shot fused entry points
                                                                 shard their host columns by row range over all of them */
void rfb_ops_shutdown(void);
const rfb_host_api_t *rfb_ops_builtin_host(void); /* malloc-based host for standalone use */
const char *rfb_ops_last_error(void);
void rfb_ops_set_min_rows(int64_t n); /* vectors shorter than this are DECLINED (default: env RFB200_MIN_ROWS or 0) */
int64_t rfb_ops_launches(void);       /* kernels launched so far (evidence that the GPU path ran) */
/* Cost gate (default on; env RFB200_GATE=0): OUTSIDE a query scope the single-touch element-wise operators (comparisons, arithmetic,
 * round/floor/ceil, not, and/or) are declined when no vector operand is in HBM and none is a column worth keeping there — shipping the
 * operands in and the result out over PCIe costs more than the reference's CPU loop over the same bytes. */
void rfb_ops_set_gate(int on);

/* Query scope: between begin and end a host column is shipped to HBM once (cudaMemcpyAsync) and reused by every
 * operator that sees the same (payload pointer, length, type); results produced on the device stay resident too.
 * Outside a scope every call ships its operands.  Bind begin/end around ray_select (core/query.c:607). */
void rfb_ops_scope_begin(void);
void rfb_ops_scope_end(void);
/* Residency across queries.  With it on, the HBM image of a host vector that outlives the query (a table column, a global:
 * refcount >= 2, allocated in the host's own heap) stays on the device after the scope ends and is found again by its
 * (payload pointer, length, type) the next time an operator sees that vector — a column is shipped once, not once per query.
 * This is only sound when the host reports every free and every in-place modification of its vectors:
 *     rfb_ops_note_free(obj)    the object `obj` (header address) is being freed / its block reused / reallocated
 *     rfb_ops_note_write(obj)   the object is about to be modified in place (copy-on-write hit with refcount 1, `and`/`or`
 *                               folding into their first operand, a CPU body reusing an operand as its result)
 * integration/rayforce_shim.c binds them to the reference's heap_free, heap_realloc, cow_obj, ray_and / ray_or and the CPU
 * fallbacks of the wrapped math operators.  Both may be called from any host thread.  Without the hooks leave residency off
 * (the default): every scope then starts cold.  budget_bytes <= 0 keeps the current budget (default: half the free HBM, or
 * env RFB200_RESIDENT_MB).  out = {image hits, columns shipped, images forgotten through the hooks, resident MiB}. */
void rfb_ops_set_residency(int on, int64_t budget_bytes);
void rfb_ops_note_free(const void *obj);
void rfb_ops_note_write(const void *obj);
void rfb_ops_residency_stats(long out[4]);

/* EXPERIMENTAL, env RFB200_LAZY=1: inside a scope, results of at least RFB200_LAZY_MIN bytes (default 32 MiB) stay on the
 * device; their host payload pages are protected and filled on the first CPU access (SIGSEGV handler) or at scope end.
 * out = {results left lazy, faulted in by a CPU access, dropped because the host had already freed them, filled at scope end}. */
void rfb_ops_lazy_stats(long out[4]);
void rfb_ops_set_lazy(int on, int64_t min_bytes); /* same switch at run time (min_bytes <= 0 keeps the threshold) */

/* ---- predicate scan: ray_eq/ne/lt/gt/le/ge (core/cmp.c:692-697) -> B8 vector */
rfb_obj_p rfb_ray_eq(rfb_obj_p x, rfb_obj_p y);
rfb_obj_p rfb_ray_ne(rfb_obj_p x, rfb_obj_p y);
rfb_obj_p rfb_ray_lt(rfb_obj_p x, rfb_obj_p y);
rfb_obj_p rfb_ray_gt(rfb_obj_p x, rfb_obj_p y);
rfb_obj_p rfb_ray_le(rfb_obj_p x, rfb_obj_p y);
rfb_obj_p rfb_ray_ge(rfb_obj_p x, rfb_obj_p y);

/* ---- selection vector and materialise: ray_where (core/items.c:1366), filter_map / filter_collect (core/filter.c:29-165) */
rfb_obj_p rfb_ray_where(rfb_obj_p mask);
rfb_obj_p rfb_filter_map(rfb_obj_p val, rfb_obj_p index);
rfb_obj_p rfb_filter_collect(rfb_obj_p val, rfb_obj_p index);

/* ---- ungrouped reductions: ray_sum/min/max/avg/cnt (core/math.c:2388-2526).  Accept a plain vector, a MAPFILTER pair
 *      (gather + fold in one kernel, core/math.c:1874-1890) or a MAPGROUP pair (-> rfb_aggr_*). */
rfb_obj_p rfb_ray_sum(rfb_obj_p x);
rfb_obj_p rfb_ray_min(rfb_obj_p x);
rfb_obj_p rfb_ray_max(rfb_obj_p x);
rfb_obj_p rfb_ray_avg(rfb_obj_p x);
rfb_obj_p rfb_ray_cnt(rfb_obj_p x);
/* ray_med / ray_dev (core/math.c:2529-2700): vector (med: U8/I16/I64 as in the reference), MAPFILTER or MAPGROUP operand */
rfb_obj_p rfb_ray_med(rfb_obj_p x);
rfb_obj_p rfb_ray_dev(rfb_obj_p x);

/* ---- element-wise: ray_add/sub/mul/div/fdiv/mod/xbar (core/math.c:2436-2442), ray_round/floor/ceil (:2430-2432) */
rfb_obj_p rfb_ray_add(rfb_obj_p x, rfb_obj_p y);
rfb_obj_p rfb_ray_sub(rfb_obj_p x, rfb_obj_p y);
rfb_obj_p rfb_ray_mul(rfb_obj_p x, rfb_obj_p y);
rfb_obj_p rfb_ray_div(rfb_obj_p x, rfb_obj_p y);
rfb_obj_p rfb_ray_fdiv(rfb_obj_p x, rfb_obj_p y);
rfb_obj_p rfb_ray_mod(rfb_obj_p x, rfb_obj_p y);
rfb_obj_p rfb_ray_xbar(rfb_obj_p x, rfb_obj_p y);
rfb_obj_p rfb_ray_round(rfb_obj_p x);
rfb_obj_p rfb_ray_floor(rfb_obj_p x);
rfb_obj_p rfb_ray_ceil(rfb_obj_p x);

/* ---- group-by: index_group (core/index.c:2173) builds the reference's 7-element index list
 *      [type, group_count, group_ids, shift, source, filter, first_ids] (core/index.c:1696-1699; always INDEX_TYPE_IDS
 *      here, which every aggr_* accepts); group_map (core/group.c:26); aggr_* (core/aggr.c) */
rfb_obj_p rfb_index_group(rfb_obj_p keys, rfb_obj_p filter);
rfb_obj_p rfb_group_map(rfb_obj_p val, rfb_obj_p index);
rfb_obj_p rfb_aggr_sum(rfb_obj_p val, rfb_obj_p index);
rfb_obj_p rfb_aggr_min(rfb_obj_p val, rfb_obj_p index);
rfb_obj_p rfb_aggr_max(rfb_obj_p val, rfb_obj_p index);
rfb_obj_p rfb_aggr_count(rfb_obj_p val, rfb_obj_p index);
rfb_obj_p rfb_aggr_avg(rfb_obj_p val, rfb_obj_p index);
/* (aggr_sum/min/max/avg also take a PARTED column — the per-partition vectors of a parted table, normally mmapped column
 *  files — with an INDEX_TYPE_PARTEDCOMMON index (PARTED_MAP, core/aggr.c:183-260): every partition is folded on the device
 *  and the per-partition results are combined, or returned one per partition, exactly as the reference does.) */
/* aggr_med / aggr_dev (core/aggr.c:2136-2906) -> F64 vector; aggr_row / aggr_collect (core/aggr.c:3021-3136) -> LIST of
 * per-group row-id / value vectors */
rfb_obj_p rfb_aggr_med(rfb_obj_p val, rfb_obj_p index);
rfb_obj_p rfb_aggr_stddev(rfb_obj_p val, rfb_obj_p index);   /* the reference's aggr_dev (rfb_aggr_dev names the C ABI's device entry) */
rfb_obj_p rfb_aggr_row(rfb_obj_p val, rfb_obj_p index);
rfb_obj_p rfb_aggr_collect(rfb_obj_p val, rfb_obj_p index);

/* aggr_first (core/aggr.c:441-577, indices with first_ids) / aggr_last (core/aggr.c:897-1075) -> vector of val's type */
rfb_obj_p rfb_aggr_first(rfb_obj_p val, rfb_obj_p index);
rfb_obj_p rfb_aggr_last(rfb_obj_p val, rfb_obj_p index);
/* index_group_list (core/index.c:2731-2793): keys = LIST of 2..8 I64-kind key columns -> the same 7-element index as index_group */
rfb_obj_p rfb_index_group_list(rfb_obj_p keys, rfb_obj_p filter);

/* ---- masks: one fold step of the special forms `and` / `or` (core/logic.c:89-264: res = res OP next, in place in res's payload;
 *      1 = done on the device, 0 = declined, < 0 = device failure) and ray_not (core/order.c:422-443) */
int rfb_mask_logic_inplace(int is_or, rfb_obj_p res, rfb_obj_p next);
rfb_obj_p rfb_ray_not(rfb_obj_p x);

/* ---- materialise / order: at_ids (core/rayforce.c:1100-1201; ids = bare host array of row ids), ray_asc / ray_desc
 *      (core/order.c:74-244: sorted values), ray_xasc / ray_xdesc (core/order.c:246-420: a table ordered by one or several columns) */
rfb_obj_p rfb_at_ids(rfb_obj_p obj, const int64_t *ids, int64_t len);
rfb_obj_p rfb_ray_asc(rfb_obj_p x);
rfb_obj_p rfb_ray_desc(rfb_obj_p x);
rfb_obj_p rfb_ray_xasc(rfb_obj_p table, rfb_obj_p by);
rfb_obj_p rfb_ray_xdesc(rfb_obj_p table, rfb_obj_p by);

/* ---- equi-join row matching: index_left_join_obj / index_inner_join_obj (core/index.c:2886-3000;
 *      lcols / rcols = the key column itself when len == 1, else a LIST of len key columns) and ray_find on two I64-kind
 *      vectors (core/items.c:320-323).  The joined table itself is then built by the reference's join.c from these row ids. */
rfb_obj_p rfb_index_left_join_obj(rfb_obj_p lcols, rfb_obj_p rcols, int64_t len);
rfb_obj_p rfb_index_inner_join_obj(rfb_obj_p lcols, rfb_obj_p rcols, int64_t len);
rfb_obj_p rfb_index_asof_join_obj(rfb_obj_p lcols, rfb_obj_p lxcol, rfb_obj_p rcols, rfb_obj_p rxcol);   /* core/index.c:3194-3268 */
rfb_obj_p rfb_ray_find(rfb_obj_p x, rfb_obj_p y);
rfb_obj_p rfb_ray_distinct(rfb_obj_p x);            /* ray_distinct, dense I64-kind key ranges (core/index.c:551-577): ascending distinct keys */
rfb_obj_p rfb_ray_in(rfb_obj_p x, rfb_obj_p y);      /* ray_in -> index_in_i64_i64 (core/index.c:1291-1370): B8 mask of the x values present in y */

/* ---- key sort: ray_sort_asc/desc (core/sort.c:430,691) = ray_iasc/idesc (core/order.c:32): stable i64 permutation */
rfb_obj_p rfb_ray_sort_asc(rfb_obj_p x);
rfb_obj_p rfb_ray_sort_desc(rfb_obj_p x);

/* ---- fused query entry points (what the evaluator's operator-at-a-time protocol cannot express).  Plugin-style
 *      functions a Rayfall script can bind with (loadfn "librfb200_ops.so" "<name>" <arity>) (core/dynlib.c:191):
 *      rfb_where_lt_sum(col, k)   = (sum col) over rows where (< col k), one 8 B/row pass
 *      rfb_where_fold(args, n)    = args: [cmp-op i64 atom 0..5 (eq ne lt gt le ge), fold i64 atom 0..3 (sum min max avg),
 *                                   pred column, constant atom, value column] */
rfb_obj_p rfb_where_lt_sum(rfb_obj_p col, rfb_obj_p k);
rfb_obj_p rfb_where_fold(rfb_obj_p *args, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* RFB200_OPS_H */
