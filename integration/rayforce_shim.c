/*
 * rayforce_shim.c — the reference-side binding of rayforce-b200.
 *
 * This file is what a RayforceDB maintainer adds to the reference build.  It is compiled against the reference's own
 * headers and linked with the reference's UNMODIFIED objects using GNU ld's `--wrap`: every reference to one of the
 * hot-path operators (from the builtin table in core/env.c:119-271, from core/query.c, core/eval.c, ...) is redirected
 * to `__wrap_<name>` below, which offers the call to the GPU operator layer (include/rfb200_ops.h) and otherwise
 * continues with `__real_<name>`, the reference's CPU body.
 *
 *   gcc -include $REF/core/def.h -I$REF -Iinclude -c integration/rayforce_shim.c
 *   gcc -o rayforce_dropin $REF_OBJECTS rayforce_shim.o -Lrayforce_b200 -lrfb200_ops -lrfb200 \
 *       -Wl,--wrap=ray_lt,--wrap=ray_sum,... (the list is WRAPPED_SYMBOLS in oracle/Makefile)
 *
 * The GPU layer DECLINES (returns NULL) whatever is outside its path — atoms, lists, tables, parted/MAPCOMMON columns,
 * symbols/GUIDs, vectors shorter than RFB200_MIN_ROWS — so the reference's behaviour for those is untouched.
 * Environment: RFB200_DISABLE=1 keeps every call on the CPU bodies; RFB200_MIN_ROWS=n sets the size gate; RFB200_DEVICES=all
 * binds every visible GPU (the fused one-shot entry points shard host columns over them); RFB200_RESIDENT=0 / RFB200_GATE=0 /
 * RFB200_LAZY=1 switch cross-query residency, the cost gate and lazily materialised results;
 * RFB200_SHIM_STATS=1 prints per-operator GPU/CPU call counts and the kernel-launch count at exit.
 */
#include <pthread.h>
#include <time.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "core/rayforce.h"
#include "core/error.h"
#include "core/ops.h"
#include "core/eval.h"
#include "core/pool.h"

#include "rfb200_ops.h"

/* ---- the host API vtable: the reference's own allocator and error constructors */
static rfb_obj_p h_vector(int8_t t, int64_t n) { return (rfb_obj_p)vector(t, n); }
static rfb_obj_p h_atom(int8_t t) { return (rfb_obj_p)atom(t); }
static rfb_obj_p h_clone(rfb_obj_p o) { return (rfb_obj_p)clone_obj((obj_p)o); }
static void h_drop(rfb_obj_p o) { drop_obj((obj_p)o); }
static rfb_obj_p h_err_type(void) { return (rfb_obj_p)err_type(0, 0, 0, 0); }
static rfb_obj_p h_err_length(void) { return (rfb_obj_p)err_length(0, 0, 0, 0, 0, 0); }
static rfb_obj_p h_err_limit(void) { return (rfb_obj_p)err_limit(0); }
static int64_t h_executors(void) { return (int64_t)pool_get_executors_count(pool_get()); }
static rfb_host_api_t host_api;

static int state = 0; /* 0 = not tried, 1 = GPU layer bound, -1 = unavailable (CPU bodies only) */
static pthread_t owner;
static int want_stats = 0;

enum { S_EQ, S_NE, S_LT, S_GT, S_LE, S_GE, S_WHERE, S_COLLECT, S_SUM, S_MIN, S_MAX, S_AVG, S_ADD, S_SUB, S_MUL, S_DIV, S_FDIV,
       S_MOD, S_XBAR, S_ROUND, S_FLOOR, S_CEIL, S_INDEX_GROUP, S_AGGR_SUM, S_AGGR_MIN, S_AGGR_MAX, S_AGGR_COUNT, S_AGGR_AVG, S_SORT_ASC,
       S_SORT_DESC, S_SELECT, S_MED, S_DEV, S_AGGR_MED, S_AGGR_DEV, S_AGGR_ROW, S_AGGR_COLLECT, S_FIND, S_LEFT_JOIN, S_INNER_JOIN, S_IN, S_ASOF_JOIN, S_DISTINCT,
       S_AND, S_OR, S_NOT, S_AT_IDS, S_ASC, S_DESC, S_XASC, S_XDESC, S_AGGR_FIRST, S_AGGR_LAST, S_GROUP_LIST, S_N };
static const char *S_NAME[S_N] = {"ray_eq", "ray_ne", "ray_lt", "ray_gt", "ray_le", "ray_ge", "ray_where", "filter_collect", "ray_sum",
                                  "ray_min", "ray_max", "ray_avg", "ray_add", "ray_sub", "ray_mul", "ray_div", "ray_fdiv", "ray_mod", "ray_xbar",
                                  "ray_round", "ray_floor", "ray_ceil", "index_group", "aggr_sum", "aggr_min", "aggr_max", "aggr_count",
                                  "aggr_avg", "ray_sort_asc", "ray_sort_desc", "ray_select", "ray_med", "ray_dev", "aggr_med", "aggr_dev",
                                  "aggr_row", "aggr_collect", "ray_find", "index_left_join_obj", "index_inner_join_obj", "ray_in", "index_asof_join_obj", "ray_distinct",
                                  "ray_and", "ray_or", "ray_not", "at_ids", "ray_asc", "ray_desc", "ray_xasc", "ray_xdesc", "aggr_first", "aggr_last",
                                  "index_group_list"};
static long n_gpu[S_N], n_cpu[S_N];
static double t_gpu[S_N], t_cpu[S_N];   /* wall milliseconds per family (only measured with RFB200_SHIM_STATS=1) */
static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* RFB200_SHIM_TRACE=<ms>: report every GPU-served operator call that took longer (first-touch shipments, fault-ins) */
static void trace_slow(int slot, double dt) {
    static double limit = -2.0;
    if (limit < -1.0) { const char *e = getenv("RFB200_SHIM_TRACE"); limit = e ? atof(e) : -1.0; }
    if (limit >= 0.0 && dt >= limit) fprintf(stderr, "[rfb200 shim] slow call: %-16s %10.2f ms\n", S_NAME[slot], dt);
}

static void print_stats(void) {
    long g = 0, c = 0;
    for (int i = 0; i < S_N; i++) { g += n_gpu[i]; c += n_cpu[i]; }
    fprintf(stderr, "[rfb200 shim] operator calls handled on the GPU: %ld, on the reference CPU bodies: %ld, kernels launched: %lld\n", g, c,
            (long long)rfb_ops_launches());
    for (int i = 0; i < S_N; i++)
        if (n_gpu[i] || n_cpu[i]) fprintf(stderr, "[rfb200 shim]   %-16s gpu %8ld   cpu %8ld   ms on gpu %10.2f   ms on cpu %10.2f\n", S_NAME[i], n_gpu[i], n_cpu[i], t_gpu[i], t_cpu[i]);
    long lz[4];
    rfb_ops_lazy_stats(lz);
    long rs[4];
    rfb_ops_residency_stats(rs);
    fprintf(stderr, "[rfb200 shim] HBM residency: %ld operand images found in HBM, %ld columns shipped, %ld images dropped by the free / write hooks, %ld MiB resident\n",
            rs[0], rs[1], rs[2], rs[3]);
    if (lz[0]) fprintf(stderr, "[rfb200 shim] lazy results: %ld left on the device, %ld faulted in by a CPU access, %ld dropped unread, %ld filled at scope end\n", lz[0], lz[1], lz[2], lz[3]);
}

static int stats_on(void) {
    static int known = -1;
    if (known < 0) known = getenv("RFB200_SHIM_STATS") != NULL;
    return known;
}

static int gpu_ok(void) {
    if (state == 0) {
        const char *d = getenv("RFB200_DISABLE");
        want_stats = getenv("RFB200_SHIM_STATS") != NULL;
        host_api.vector = h_vector; host_api.atom = h_atom; host_api.clone_obj = h_clone; host_api.drop_obj = h_drop;
        host_api.err_type = h_err_type; host_api.err_length = h_err_length; host_api.err_limit = h_err_limit;
        host_api.null_obj = (rfb_obj_p)NULL_OBJ;
        host_api.executors = h_executors;
        if (d && d[0] == '1') state = -1;
        else if (rfb_ops_init(&host_api, (getenv("RFB200_DEVICES") && !strcmp(getenv("RFB200_DEVICES"), "all")) ? -1 : 0) == 0) {
            state = 1;
            owner = pthread_self();
            /* this binding reports every free / in-place write (heap_free, heap_realloc, cow_obj, the CPU fallbacks below), so
             * column images may stay in HBM across queries; RFB200_RESIDENT=0 keeps every query cold */
            const char *r = getenv("RFB200_RESIDENT");
            rfb_ops_set_residency(!(r && r[0] == '0'), 0);
        }
        else {
            state = -1;
            fprintf(stderr, "[rfb200 shim] GPU layer unavailable (%s): using the reference CPU bodies\n", rfb_ops_last_error());
        }
        if (want_stats) atexit(print_stats);
    }
    /* objects must be allocated on the calling thread's heap and the layer keeps per-query state: one owner thread */
    return state == 1 && pthread_equal(owner, pthread_self());
}

#define WRAP1(sym, slot)                                                   \
    obj_p __real_##sym(obj_p x);                                           \
    obj_p __wrap_##sym(obj_p x) {                                          \
        const double t0 = stats_on() ? now_ms() : 0.0;                     \
        if (gpu_ok()) {                                                    \
            obj_p r = (obj_p)rfb_##sym((rfb_obj_p)x);                      \
            if (r) { n_gpu[slot]++; if (want_stats) { const double dt = now_ms() - t0; t_gpu[slot] += dt; trace_slow(slot, dt); } return r; } \
        }                                                                  \
        n_cpu[slot]++;                                                     \
        obj_p c = __real_##sym(x);                                         \
        if (want_stats) t_cpu[slot] += now_ms() - t0;                      \
        return c;                                                          \
    }
#define WRAP2(sym, slot)                                                   \
    obj_p __real_##sym(obj_p x, obj_p y);                                  \
    obj_p __wrap_##sym(obj_p x, obj_p y) {                                 \
        const double t0 = stats_on() ? now_ms() : 0.0;                     \
        if (gpu_ok()) {                                                    \
            obj_p r = (obj_p)rfb_##sym((rfb_obj_p)x, (rfb_obj_p)y);        \
            if (r) { n_gpu[slot]++; if (want_stats) { const double dt = now_ms() - t0; t_gpu[slot] += dt; trace_slow(slot, dt); } return r; } \
        }                                                                  \
        n_cpu[slot]++;                                                     \
        obj_p c = __real_##sym(x, y);                                      \
        if (want_stats) t_cpu[slot] += now_ms() - t0;                      \
        return c;                                                          \
    }

WRAP2(ray_eq, S_EQ) WRAP2(ray_ne, S_NE) WRAP2(ray_lt, S_LT) WRAP2(ray_gt, S_GT) WRAP2(ray_le, S_LE) WRAP2(ray_ge, S_GE)
WRAP1(ray_where, S_WHERE) WRAP2(filter_collect, S_COLLECT)
WRAP1(ray_sum, S_SUM) WRAP1(ray_min, S_MIN) WRAP1(ray_max, S_MAX) WRAP1(ray_avg, S_AVG)
WRAP2(ray_add, S_ADD) WRAP2(ray_sub, S_SUB) WRAP2(ray_mul, S_MUL) WRAP2(ray_div, S_DIV) WRAP2(ray_fdiv, S_FDIV) WRAP2(ray_mod, S_MOD) WRAP2(ray_xbar, S_XBAR)
WRAP1(ray_round, S_ROUND) WRAP1(ray_floor, S_FLOOR) WRAP1(ray_ceil, S_CEIL)
WRAP2(index_group, S_INDEX_GROUP)
WRAP2(aggr_sum, S_AGGR_SUM) WRAP2(aggr_min, S_AGGR_MIN) WRAP2(aggr_max, S_AGGR_MAX) WRAP2(aggr_count, S_AGGR_COUNT) WRAP2(aggr_avg, S_AGGR_AVG)
WRAP1(ray_sort_asc, S_SORT_ASC) WRAP1(ray_sort_desc, S_SORT_DESC)
WRAP1(ray_med, S_MED) WRAP1(ray_dev, S_DEV)
WRAP2(ray_find, S_FIND) WRAP2(ray_in, S_IN) WRAP1(ray_distinct, S_DISTINCT)
#define WRAPJ(sym, slot)                                                   \
    obj_p __real_##sym(obj_p l, obj_p r, i64_t n);                         \
    obj_p __wrap_##sym(obj_p l, obj_p r, i64_t n) {                        \
        if (gpu_ok()) {                                                    \
            obj_p res = (obj_p)rfb_##sym((rfb_obj_p)l, (rfb_obj_p)r, n);   \
            if (res) { n_gpu[slot]++; return res; }                        \
        }                                                                  \
        n_cpu[slot]++;                                                     \
        return __real_##sym(l, r, n);                                      \
    }
WRAPJ(index_left_join_obj, S_LEFT_JOIN) WRAPJ(index_inner_join_obj, S_INNER_JOIN)
obj_p __real_index_asof_join_obj(obj_p lc, obj_p lx, obj_p rc, obj_p rx);
obj_p __wrap_index_asof_join_obj(obj_p lc, obj_p lx, obj_p rc, obj_p rx) {
    if (gpu_ok()) {
        obj_p res = (obj_p)rfb_index_asof_join_obj((rfb_obj_p)lc, (rfb_obj_p)lx, (rfb_obj_p)rc, (rfb_obj_p)rx);
        if (res) { n_gpu[S_ASOF_JOIN]++; return res; }
    }
    n_cpu[S_ASOF_JOIN]++;
    return __real_index_asof_join_obj(lc, lx, rc, rx);
}
#define rfb_aggr_dev rfb_aggr_stddev   /* the operator layer's name for the reference's aggr_dev */
WRAP2(aggr_med, S_AGGR_MED) WRAP2(aggr_dev, S_AGGR_DEV) WRAP2(aggr_row, S_AGGR_ROW) WRAP2(aggr_collect, S_AGGR_COLLECT)

/* a query is the residency scope: columns are shipped to HBM once per select (core/query.c:607) */
obj_p __real_ray_select(obj_p obj);
obj_p __wrap_ray_select(obj_p obj) {
    const int on = gpu_ok();
    const double t0 = stats_on() ? now_ms() : 0.0;
    if (on) rfb_ops_scope_begin();
    n_cpu[S_SELECT]++;
    obj_p r = __real_ray_select(obj);
    if (on) rfb_ops_scope_end();
    if (want_stats) t_cpu[S_SELECT] += now_ms() - t0;   /* the whole query, operators included */
    return r;
}

/* ---- residency hooks: the layer keeps HBM images of host vectors keyed by their payload address; the reference tells it when
 * an address stops meaning what it meant.  heap_free / heap_realloc see every internal block going away (drop_obj ends there),
 * cow_obj returning its argument means "about to be modified in place". */
nil_t __real_heap_free(raw_p ptr);
nil_t __wrap_heap_free(raw_p ptr) {
    if (state == 1 && ptr) rfb_ops_note_free(ptr);
    __real_heap_free(ptr);
}
raw_p __real_heap_realloc(raw_p ptr, i64_t size);
raw_p __wrap_heap_realloc(raw_p ptr, i64_t size) {
    if (state == 1 && ptr) rfb_ops_note_free(ptr);
    return __real_heap_realloc(ptr, size);
}
obj_p __real_cow_obj(obj_p obj);
obj_p __wrap_cow_obj(obj_p obj) {
    obj_p r = __real_cow_obj(obj);
    if (state == 1 && r == obj) rfb_ops_note_write(obj);
    return r;
}

WRAP1(ray_not, S_NOT) WRAP1(ray_asc, S_ASC) WRAP1(ray_desc, S_DESC) WRAP2(ray_xasc, S_XASC) WRAP2(ray_xdesc, S_XDESC)
WRAP2(aggr_first, S_AGGR_FIRST) WRAP2(aggr_last, S_AGGR_LAST) WRAP2(index_group_list, S_GROUP_LIST)

obj_p __real_at_ids(obj_p obj, i64_t ids[], i64_t len);
obj_p __wrap_at_ids(obj_p obj, i64_t ids[], i64_t len) {
    if (gpu_ok()) {
        obj_p r = (obj_p)rfb_at_ids((rfb_obj_p)obj, (const int64_t *)ids, len);
        if (r) { n_gpu[S_AT_IDS]++; return r; }
    }
    n_cpu[S_AT_IDS]++;
    return __real_at_ids(obj, ids, len);
}

/* `and` / `or` are special forms (core/logic.c:89-264): they evaluate their operands one by one and fold each into the first
 * result IN PLACE.  The wrapper keeps that protocol — same evaluation order, same in-place result object — and offers every
 * (B8 vector, B8 vector | b8 atom) step to the device; any other step goes through the reference's own body on the two
 * already-evaluated operands (values other than LISTs and symbol atoms evaluate to themselves, core/eval.c:884-893). */
obj_p __real_ray_and(obj_p *x, i64_t n);
obj_p __real_ray_or(obj_p *x, i64_t n);
static obj_p logic_fold(int is_or, obj_p *x, i64_t n) {
    const int slot = is_or ? S_OR : S_AND;
    if (!gpu_ok() || n < 2) { n_cpu[slot]++; return is_or ? __real_ray_or(x, n) : __real_ray_and(x, n); }
    obj_p res = eval(x[0]);
    if (IS_ERR(res)) return res;
    int on_gpu = 0;
    for (i64_t i = 1; i < n; i++) {
        obj_p next = eval(x[i]);
        if (IS_ERR(next)) { drop_obj(res); return next; }
        if (res->type == -TYPE_B8 && next->type == TYPE_B8) { obj_p t = res; res = next; next = t; }   /* core/logic.c:178-182 */
        if (rfb_mask_logic_inplace(is_or, (rfb_obj_p)res, (rfb_obj_p)next) == 1) { on_gpu = 1; drop_obj(next); continue; }
        if (res->type == TYPE_LIST || res->type == -TYPE_SYMBOL || next->type == TYPE_LIST || next->type == -TYPE_SYMBOL) {
            drop_obj(res);                             /* would be evaluated a second time: the reference's answer for these is a type error */
            drop_obj(next);
            return err_type(0, 0, 0, 0);
        }
        rfb_ops_note_write(res);                       /* the CPU body folds into res's payload */
        obj_p pair[2] = {res, next};
        obj_p r = is_or ? __real_ray_or(pair, 2) : __real_ray_and(pair, 2);   /* takes its own references, returns one to res (or an error) */
        drop_obj(res);
        drop_obj(next);
        if (IS_ERR(r)) return r;
        res = r;
    }
    if (on_gpu) n_gpu[slot]++; else n_cpu[slot]++;
    return res;
}
obj_p __wrap_ray_and(obj_p *x, i64_t n) { return logic_fold(0, x, n); }
obj_p __wrap_ray_or(obj_p *x, i64_t n) { return logic_fold(1, x, n); }
