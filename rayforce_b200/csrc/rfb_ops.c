/*
 * rfb_ops.c — reference-facing operator layer (pure C) on top of the C ABI of include/rfb200.h.
 * See include/rfb200_ops.h for the contract.  No arithmetic on column data happens in this file: it classifies the
 * operands exactly like the reference's dispatchers do, ships column payloads to HBM (once per query scope), calls the
 * device-layer entry points and wraps the results into host-allocated objects.
 */
#define _GNU_SOURCE
#include "rfb200_ops.h"

#include <math.h>
#include <signal.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#include "rfb200.h"

typedef rfb_obj_p obj_p;

/* synchronous, driver-staged copy that involves none of this library's helper threads (safe inside the fault handler) */
static int rfb_d2h_plain(rfb_ctx_t *ctx, void *dst_host, const void *src_dev, size_t bytes) { return rfb_d2h_sync_plain(ctx, dst_host, src_dev, bytes); }

/* ------------------------------------------------------------------ state */

/* A device image of a host vector's payload, or a scratch / result buffer (host == NULL).
 * Identity of a host vector = (payload pointer, length, type).  That identity is only trustworthy while the host has neither
 * freed nor modified the object.  The layer learns about both from hooks the binding installs (rfb_ops_note_free /
 * rfb_ops_note_write: the reference's heap_free, heap_realloc, cow_obj, the in-place CPU bodies of the wrapped math operators and
 * `and` / `or`, see integration/rayforce_shim.c); the sampled fingerprint stays as a second line of defence.  Entries are found
 * newest-first, registering a payload forgets every older entry with the same pointer, and nothing is recycled while the
 * operator call that handed it out is still running (entries are only released at the end of the outermost scope). */
typedef struct {
    const void *host; /* payload pointer identity; NULL = scratch */
    int64_t len;
    int type;
    void *dev;
    size_t bytes;
    uint64_t print;   /* fingerprint of sampled payload words */
    uint64_t used;    /* LRU clock of the last operator that touched it */
    int keep;         /* survives the end of its scope: an image of a live host vector, kept resident across queries */
} col_entry_t;

typedef struct {
    void *dev;
    size_t bytes;
} pool_entry_t;

#define POOL_MAX 64
static struct {
    int ready;
    const rfb_host_api_t *host;
    rfb_ctx_t *ctx;
    rfb_mgpu_t *mgpu;            /* rfb_ops_init(host, -1): every visible device; ctx = device 0's context */
    int64_t min_rows;
    int scope_depth;
    col_entry_t *cols;
    int ncols, cap;
    pool_entry_t pool[POOL_MAX]; /* free device buffers kept for reuse */
    int npool;
    uint64_t clock;
    int residency;               /* images of live host vectors stay in HBM after the scope ends (needs the hooks) */
    size_t resident_bytes, resident_budget;
    uint64_t bloom[4];           /* payload pointers with an entry: lets the free hook leave after one test */
    volatile int hook_lock;
    long stat_hits, stat_ships, stat_forgotten;
    char err[256];
} G;

static void set_err(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(G.err, sizeof(G.err), fmt, ap);
    va_end(ap);
}
const char *rfb_ops_last_error(void) { return G.err; }

/* lazily materialised results (defined further down) */
static int lazy_on = -1;
static long lazy_stats[4]; /* registered, faulted in, dropped, filled at scope end */
static void lazy_resolve_all(void);
static int lazy_backed_by(const void *dev);
static void lazy_resolve_dev(const void *dev);

/* ------------------------------------------------------------------ builtin malloc host (standalone / tests) */

static int type_size(int t) {
    switch (t) {
        case RFB_T_B8: case RFB_T_U8: return 1;
        case RFB_T_I16: return 2;
        case RFB_T_I32: case RFB_T_DATE: case RFB_T_TIME: return 4;
        case RFB_T_I64: case RFB_T_SYMBOL: case RFB_T_TIMESTAMP: case RFB_T_F64: return 8;
        default: return 0;
    }
}
static int is_listlike(int t) { return t == RFB_T_LIST || t == RFB_T_MAPFILTER || t == RFB_T_MAPGROUP; }

static rfb_obj_t bh_null = {0, 0, RFB_T_NULL, 0, 1, {0}};
static rfb_obj_t bh_err = {0, 0, RFB_T_ERR, 0, 1, {0}};
static __thread const char *bh_last_err = "";

static obj_p bh_vector(int8_t type, int64_t len) {
    const int t = type < 0 ? -type : type;
    const int64_t w = (t == RFB_T_LIST) ? 8 : type_size(t);
    obj_p o = (obj_p)calloc(1, 16 + (size_t)(len > 0 ? len : 0) * (size_t)(w ? w : 8) + 16);
    if (!o) return NULL;
    o->type = (int8_t)t;
    o->rc = 1;
    o->len = len;
    return o;
}
static obj_p bh_atom(int8_t type) {
    obj_p o = (obj_p)calloc(1, 16);
    if (!o) return NULL;
    o->type = (int8_t)-type;
    o->rc = 1;
    return o;
}
static obj_p bh_clone(obj_p o) {
    if (o && o != &bh_null && o != &bh_err) __atomic_add_fetch(&o->rc, 1, __ATOMIC_RELAXED);
    return o;
}
static void bh_drop(obj_p o) {
    if (!o || o == &bh_null || o == &bh_err) return;
    if (__atomic_sub_fetch(&o->rc, 1, __ATOMIC_ACQ_REL) != 0) return;
    if (is_listlike(o->type))
        for (int64_t i = 0; i < o->len; i++) bh_drop(RFB_OBJ_LIST(o)[i]);
    free(o);
}
static obj_p bh_err_type(void) { bh_last_err = "type"; return &bh_err; }
static obj_p bh_err_length(void) { bh_last_err = "length"; return &bh_err; }
static obj_p bh_err_limit(void) { bh_last_err = "limit"; return &bh_err; }
const char *rfb_ops_builtin_last_error(void) { return bh_last_err; }

static const rfb_host_api_t builtin_host = {bh_vector, bh_atom, bh_clone, bh_drop, bh_err_type, bh_err_length, bh_err_limit, &bh_null, NULL};
const rfb_host_api_t *rfb_ops_builtin_host(void) { return &builtin_host; }

/* ------------------------------------------------------------------ init / scope / device columns */

int rfb_ops_init(const rfb_host_api_t *host, int device) {
    if (G.ready) return 0;
    if (!host) { set_err("rfb_ops_init: host api is NULL"); return RFB_ERR_ARG; }
    int rc;
    if (device < 0) {            /* every visible GPU: the operator-at-a-time path runs on device 0, the one-shot fused entry points
                                    (host column in, result out) shard the column by row range over all of them */
        rc = rfb_mgpu_create(0, &G.mgpu);
        if (!rc) G.ctx = rfb_mgpu_ctx(G.mgpu, 0);
    } else rc = rfb_ctx_create(device, &G.ctx);
    if (rc) { set_err("%s", rfb_last_error()); return rc; }
    G.host = host;
    const char *e = getenv("RFB200_MIN_ROWS");
    G.min_rows = e ? atoll(e) : 0;
    size_t free_b = 0, total_b = 0;
    e = getenv("RFB200_RESIDENT_MB");
    if (e) G.resident_budget = (size_t)atoll(e) << 20;
    else if (rfb_dev_mem_info(G.ctx, &free_b, &total_b) == RFB_OK) G.resident_budget = free_b / 2;   /* the other half: results, workspaces */
    else G.resident_budget = (size_t)32 << 30;
    G.ready = 1;
    return 0;
}

static unsigned bloom_bit(const void *p) { return (unsigned)((((uintptr_t)p >> 4) * 0x9E3779B97F4A7C15ULL) >> 56); }
static void bloom_add(const void *p) { const unsigned b = bloom_bit(p); G.bloom[b >> 6] |= 1ULL << (b & 63); }
static int bloom_has(const void *p) { const unsigned b = bloom_bit(p); return (int)((G.bloom[b >> 6] >> (b & 63)) & 1); }

static void free_buffer(void *dev, size_t bytes) {
    if (G.npool < POOL_MAX) { G.pool[G.npool].dev = dev; G.pool[G.npool].bytes = bytes; G.npool++; }
    else rfb_dev_free(G.ctx, dev);
}

/* end of the outermost scope: scratch buffers go back to the pool; images of live host vectors stay when residency is on,
 * oldest ones first out when they exceed the budget */
static void release_columns(int everything) {
    int w = 0;
    G.resident_bytes = 0;
    memset(G.bloom, 0, sizeof(G.bloom));
    for (int i = 0; i < G.ncols; i++) {
        col_entry_t *e = &G.cols[i];
        /* with the hooks bound, a result the host has not read yet (lazy) keeps its device buffer: the host frees it (no copy
         * at all), reads it (filled by the fault) or the budget pushes it out (filled first) */
        const int pending = !everything && G.residency && lazy_on == 1 && lazy_backed_by(e->dev);
        if (!everything && G.residency && ((e->host && e->keep) || pending)) {
            G.resident_bytes += e->bytes;
            if (e->host) bloom_add(e->host);
            G.cols[w++] = *e;
        } else {
            if (lazy_on == 1) lazy_resolve_dev(e->dev);
            free_buffer(e->dev, e->bytes);
        }
    }
    G.ncols = w;
    while (G.resident_bytes > G.resident_budget && G.ncols > 0) {
        int lru = 0;
        for (int i = 1; i < G.ncols; i++)
            if (G.cols[i].used < G.cols[lru].used) lru = i;
        G.resident_bytes -= G.cols[lru].bytes;
        if (lazy_on == 1) lazy_resolve_dev(G.cols[lru].dev);
        free_buffer(G.cols[lru].dev, G.cols[lru].bytes);
        G.cols[lru] = G.cols[--G.ncols];
    }
}

void rfb_ops_shutdown(void) {
    if (!G.ready) return;
    rfb_sync(G.ctx);
    if (lazy_on == 1) lazy_resolve_all();
    release_columns(1);
    for (int i = 0; i < G.npool; i++) rfb_dev_free(G.ctx, G.pool[i].dev);
    G.npool = 0;
    free(G.cols);
    if (G.mgpu) rfb_mgpu_destroy(G.mgpu); else rfb_ctx_destroy(G.ctx);
    memset(&G, 0, sizeof(G));
}

void rfb_ops_set_residency(int on, int64_t budget_bytes) {
    if (!G.ready) return;
    if (!on && G.residency && G.scope_depth == 0) { rfb_sync(G.ctx); if (lazy_on == 1) lazy_resolve_all(); G.residency = 0; release_columns(1); }
    G.residency = on ? 1 : 0;
    if (budget_bytes > 0) G.resident_budget = (size_t)budget_bytes;
}
void rfb_ops_residency_stats(long out[4]) { out[0] = G.stat_hits; out[1] = G.stat_ships; out[2] = G.stat_forgotten; out[3] = (long)(G.resident_bytes >> 20); }

/* ---- hooks: the host tells the layer that an object is gone or is about to change under its pointer.  Callable from any
 * host thread (the reference frees temporaries on its pool workers); the operator entry points themselves run on one thread
 * and never concurrently with a hook (workers only run while that thread waits in the host's pool). */
static void lazy_on_free(const void *obj);
static void forget_payload(const void *payload, const col_entry_t *except) {
    for (int i = G.ncols - 1; i >= 0; i--)
        if (G.cols[i].host == payload && &G.cols[i] != except) {
            G.cols[i].host = NULL;           /* now a scratch buffer: released with the scope (never recycled mid-call) */
            G.cols[i].keep = 0;
            G.stat_forgotten++;
            if (G.scope_depth == 0) {        /* between queries: give the memory back right away */
                G.resident_bytes -= G.cols[i].bytes < G.resident_bytes ? G.cols[i].bytes : G.resident_bytes;
                free_buffer(G.cols[i].dev, G.cols[i].bytes);
                G.cols[i] = G.cols[--G.ncols];
            }
        }
}
void rfb_ops_note_free(const void *obj) {
    if (!G.ready || !obj) return;
    const void *payload = (const char *)obj + 16;
    if (lazy_on == 1) lazy_on_free(obj);
    if (G.ncols == 0 || !bloom_has(payload)) return;
    while (__sync_lock_test_and_set(&G.hook_lock, 1)) { }
    forget_payload(payload, NULL);
    __sync_lock_release(&G.hook_lock);
}
void rfb_ops_note_write(const void *obj) { rfb_ops_note_free(obj); }

void rfb_ops_set_min_rows(int64_t n) { G.min_rows = n; }
void rfb_ops_lazy_stats(long out[4]) { out[0] = lazy_stats[0]; out[1] = lazy_stats[1]; out[2] = lazy_stats[2]; out[3] = lazy_stats[3]; }
int64_t rfb_ops_launches(void) { return G.ready ? rfb_launch_count(G.ctx) : 0; }
void rfb_ops_scope_begin(void) { G.scope_depth++; }
void rfb_ops_scope_end(void) {
    if (G.scope_depth > 0 && --G.scope_depth == 0 && G.ready) {
        rfb_sync(G.ctx);
        if (lazy_on == 1 && !G.residency) lazy_resolve_all();   /* without the hooks nothing may stay pending past the query */
        release_columns(0);
    }
}

/* 64 strided 8-byte samples + the tail, mixed.  The host may free a vector and get the same block back for a different
 * vector of the same shape inside one scope; a stale HBM image must not be served for it. */
static uint64_t fingerprint(const void *payload, size_t bytes) {
    uint64_t h = 0x9E3779B97F4A7C15ULL ^ bytes, w;
    if (bytes < 8) { w = 0; memcpy(&w, payload, bytes); return h ^ w; }
    const size_t step = (bytes / 64) & ~(size_t)7;
    for (int i = 0; i < 64; i++) {
        memcpy(&w, (const char *)payload + (step ? step * i : 0), 8);
        h = (h ^ w) * 0xBF58476D1CE4E5B9ULL;
        h ^= h >> 29;
        if (!step) break;
    }
    memcpy(&w, (const char *)payload + bytes - 8, 8);
    return (h ^ w) * 0x94D049BB133111EBULL;
}

/* RFB200_OPS_TRACE=<ms>: report the layer's own slow steps (column shipments, device allocations, lazy fault-ins, kernels) */
static double ops_now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static double ops_trace_limit(void) {
    static double limit = -2.0;
    if (limit < -1.0) { const char *e = getenv("RFB200_OPS_TRACE"); limit = e ? atof(e) : -1.0; }
    return limit;
}
static void ops_trace(const char *what, double t0, size_t bytes) {
    const double limit = ops_trace_limit();
    if (limit < 0.0) return;
    const double dt = ops_now_ms() - t0;
    if (dt >= limit) fprintf(stderr, "[rfb200 ops] %-22s %10.2f ms  %12zu bytes\n", what, dt, bytes);
}

/* ------------------------------------------------------------------ lazily materialised results (opt-in: RFB200_LAZY=1)
 *
 * Inside a query scope most operator results are consumed by the next GPU operator (mask -> where -> gather/fold), yet
 * the evaluator's protocol wants every result as a host object.  With RFB200_LAZY=1 a large result's payload is not
 * copied back: the pages that belong to it alone are made inaccessible (mprotect PROT_NONE) and the device buffer stays
 * attached.  The first CPU access faults; the SIGSEGV handler copies the bytes back, re-opens the pages and returns, so
 * any CPU consumer still sees correct data ("lazily materialised payload").  At the end of the scope every
 * still-pending result whose pages are still mapped and still protected is materialised; results whose block the host
 * already unmapped (a dropped >= 32 MB vector in the reference: core/heap.c:291-308) are simply forgotten.
 * EXPERIMENTAL and off by default: it installs a SIGSEGV handler in the host process. */

#define MAX_LAZY 64
typedef struct {
    char *lo, *hi;        /* protected page range (inside the payload) */
    const char *payload;  /* start of the payload */
    void *dev;            /* device image of the payload */
    volatile int state;   /* 0 free, 1 pending, 2 materialised */
} lazy_t;
static lazy_t LZ[MAX_LAZY];
static size_t lazy_min = 32u << 20;
static volatile int lazy_lock;
static struct sigaction lazy_old_segv;

static void lazy_invalidate_image(void *dev) {
    for (int i = 0; i < G.ncols; i++)
        if (G.cols[i].dev == dev) G.cols[i].host = NULL; /* the host has touched (maybe changed) the bytes */
}

static void lazy_fill(lazy_t *z, int in_handler) { /* pages -> read/write, bytes <- device */
    const double t0 = ops_trace_limit() >= 0.0 ? ops_now_ms() : 0.0;
    mprotect(z->lo, (size_t)(z->hi - z->lo), PROT_READ | PROT_WRITE);
    rfb_sync(G.ctx);
    /* inside the signal handler: plain synchronous copy, free of our own threads and locks; otherwise the fast ring */
    if (in_handler) rfb_d2h_plain(G.ctx, z->lo, (const char *)z->dev + (z->lo - z->payload), (size_t)(z->hi - z->lo));
    else { rfb_d2h(G.ctx, z->lo, (const char *)z->dev + (z->lo - z->payload), (size_t)(z->hi - z->lo)); rfb_sync(G.ctx); }
    if (!in_handler) ops_trace("lazy fill (scope end)", t0, (size_t)(z->hi - z->lo));
    else if (ops_trace_limit() >= 0.0) { char b[96]; const int k = snprintf(b, sizeof b, "[rfb200 ops] lazy fault-in %.2f ms %zu bytes\n", ops_now_ms() - t0, (size_t)(z->hi - z->lo)); if (k > 0) (void)!write(2, b, (size_t)k); }
}

static void lazy_segv(int sig, siginfo_t *si, void *uc) {
    char *a = (char *)si->si_addr;
    for (int i = 0; i < MAX_LAZY; i++) {
        lazy_t *z = &LZ[i];
        if (z->state == 1 && a >= z->lo && a < z->hi) {
            while (__sync_lock_test_and_set(&lazy_lock, 1)) { }
            if (z->state == 1) {
                lazy_fill(z, 1);
                lazy_invalidate_image(z->dev);
                z->state = 2;
                lazy_stats[1]++;
            }
            __sync_lock_release(&lazy_lock);
            return; /* retry the faulting access */
        }
    }
    /* not ours: hand over to whoever was there before (default action: die with the usual core) */
    if (lazy_old_segv.sa_flags & SA_SIGINFO) { if (lazy_old_segv.sa_sigaction) { lazy_old_segv.sa_sigaction(sig, si, uc); return; } }
    else if (lazy_old_segv.sa_handler != SIG_DFL && lazy_old_segv.sa_handler != SIG_IGN) { lazy_old_segv.sa_handler(sig); return; }
    signal(SIGSEGV, SIG_DFL);
    raise(SIGSEGV);
}

static int lazy_handler_installed;
static int lazy_install(void) {
    if (lazy_handler_installed) return 1;
    struct sigaction sa;
    memset(&sa, 0, sizeof(sa));
    sa.sa_sigaction = lazy_segv;
    sa.sa_flags = SA_SIGINFO | SA_NODEFER;
    sigemptyset(&sa.sa_mask);
    if (sigaction(SIGSEGV, &sa, &lazy_old_segv) != 0) return 0;
    lazy_handler_installed = 1;
    return 1;
}

static int lazy_enabled(void) {
    if (lazy_on < 0) {
        const char *e = getenv("RFB200_LAZY"), *m = getenv("RFB200_LAZY_MIN");
        lazy_on = (e && e[0] == '1') ? 1 : 0;
        if (m) lazy_min = (size_t)atoll(m);
        if (lazy_on && !lazy_install()) lazy_on = 0;
    }
    return lazy_on;
}

void rfb_ops_set_lazy(int on, int64_t min_bytes) {
    if (G.scope_depth == 0 && lazy_on == 1) lazy_resolve_all();
    if (min_bytes > 0) lazy_min = (size_t)min_bytes;
    lazy_on = (on && lazy_install()) ? 1 : 0;
}

/* the host frees an object whose payload is still pending: nobody ever read it — open the pages again (the allocator is
 * about to write its own words there) and forget it, without the copy */
static void lazy_on_free(const void *obj) {
    const char *payload = (const char *)obj + 16;
    for (int i = 0; i < MAX_LAZY; i++) {
        lazy_t *z = &LZ[i];
        if (z->state == 1 && z->payload == payload) {
            while (__sync_lock_test_and_set(&lazy_lock, 1)) { }
            if (z->state == 1) {
                mprotect(z->lo, (size_t)(z->hi - z->lo), PROT_READ | PROT_WRITE);
                z->state = 0;
                lazy_stats[2]++;
            }
            __sync_lock_release(&lazy_lock);
        }
    }
}

/* is [lo, hi) still one of OUR protected ranges?  /proc/self/maps (sorted by address): every byte of the range must lie in a
 * mapping without read and write permission.  The range may span SEVERAL such mappings — mprotect splits mappings, and two
 * neighbours with different histories (an allocator arena that grew by a second mmap) do not merge back into one line — so the
 * lines are walked in order; a hole or a readable / writable piece means the host has unmapped or reused the block. */
static int lazy_still_ours(const char *lo, const char *hi) {
    FILE *f = fopen("/proc/self/maps", "r");
    if (!f) return 0;
    char line[512];
    unsigned long cur = (unsigned long)lo;
    const unsigned long end = (unsigned long)hi;
    int ours = 0;
    while (fgets(line, sizeof(line), f)) {
        unsigned long a, b;
        char perms[8];
        if (sscanf(line, "%lx-%lx %7s", &a, &b, perms) != 3) continue;
        if (b <= cur) continue;                       /* below the part still to be covered */
        if (a > cur) break;                           /* a hole: unmapped */
        if (perms[0] != '-' || perms[1] != '-') break;   /* somebody else's memory now */
        cur = b;
        if (cur >= end) { ours = 1; break; }
    }
    fclose(f);
    return ours;
}

/* forget pending entries that overlap a range the host has handed out again */
static void lazy_forget_overlaps(const char *lo, const char *hi, lazy_t *except) {
    for (int i = 0; i < MAX_LAZY; i++) {
        lazy_t *z = &LZ[i];
        if (z != except && z->state == 1 && lo < z->hi && hi > z->lo) { z->state = 0; lazy_stats[2]++; }
    }
}

/* the device image of a still-pending lazy payload, or NULL */
static void *lazy_device_image(const void *payload) {
    for (int i = 0; i < MAX_LAZY; i++)
        if (LZ[i].state == 1 && LZ[i].payload == (const char *)payload) return LZ[i].dev;
    return NULL;
}

/* is a still-pending lazy payload backed by this device buffer? */
static int lazy_backed_by(const void *dev) {
    for (int i = 0; i < MAX_LAZY; i++)
        if (LZ[i].state == 1 && LZ[i].dev == dev) return 1;
    return 0;
}
/* the device buffer is about to be recycled: materialise what still hangs on it */
static void lazy_resolve_dev(const void *dev) {
    for (int i = 0; i < MAX_LAZY; i++) {
        lazy_t *z = &LZ[i];
        if (z->state != 1 || z->dev != dev) continue;
        while (__sync_lock_test_and_set(&lazy_lock, 1)) { }
        if (z->state == 1) {
            if (lazy_still_ours(z->lo, z->hi)) { lazy_fill(z, 0); lazy_stats[3]++; } else lazy_stats[2]++;
            z->state = 0;
        }
        __sync_lock_release(&lazy_lock);
    }
}

/* resolve everything that is still pending (end of the outermost scope, before the device buffers are recycled) */
static void lazy_resolve_all(void) {
    for (int i = 0; i < MAX_LAZY; i++) {
        lazy_t *z = &LZ[i];
        if (z->state != 1) { z->state = 0; continue; }
        while (__sync_lock_test_and_set(&lazy_lock, 1)) { }
        if (z->state == 1) {
            if (lazy_still_ours(z->lo, z->hi)) { lazy_fill(z, 0); lazy_stats[3]++; } else lazy_stats[2]++;
        }
        z->state = 0;
        __sync_lock_release(&lazy_lock);
    }
}

/* try to leave `r`'s payload on the device; returns 1 when the result is lazy (edges copied, interior protected) */
static int lazy_register(obj_p r, size_t bytes, void *dev) {
    if (!lazy_enabled() || bytes < lazy_min || (G.scope_depth < 2 && !G.residency)) return 0; /* depth 1 without the hooks = a bare call: would resolve at once */
    const long pg = sysconf(_SC_PAGESIZE);
    char *p = (char *)RFB_OBJ_PAYLOAD(r);
    char *lo = (char *)(((uintptr_t)p + (uintptr_t)pg - 1) & ~(uintptr_t)(pg - 1));
    char *hi = (char *)(((uintptr_t)p + bytes) & ~(uintptr_t)(pg - 1));
    if (hi <= lo + 16 * pg) return 0;
    int slot = -1;
    for (int i = 0; i < MAX_LAZY; i++)
        if (LZ[i].state == 0 || LZ[i].state == 2) { slot = i; break; }
    if (slot < 0) return 0;
    lazy_forget_overlaps((const char *)r, p + bytes, NULL);
    /* the partial pages at both ends are shared with neighbours: copy them now */
    if (lo > p && rfb_d2h_plain(G.ctx, p, dev, (size_t)(lo - p)) != RFB_OK) return 0;
    if (p + bytes > hi && rfb_d2h_plain(G.ctx, hi, (const char *)dev + (hi - p), (size_t)(p + bytes - hi)) != RFB_OK) return 0;
    if (mprotect(lo, (size_t)(hi - lo), PROT_NONE) != 0) return 0;
    lazy_t *z = &LZ[slot];
    z->lo = lo; z->hi = hi; z->payload = p; z->dev = dev;
    __sync_synchronize();
    z->state = 1;
    lazy_stats[0]++;
    return 1;
}

/* a device buffer of at least `bytes` (from the reuse pool when one fits within 2x) */
static void *dev_buffer(size_t bytes, size_t *got) {
    if (bytes == 0) bytes = 16;
    int best = -1;
    for (int i = 0; i < G.npool; i++)
        if (G.pool[i].bytes >= bytes && G.pool[i].bytes <= 2 * bytes + 4096 && (best < 0 || G.pool[i].bytes < G.pool[best].bytes)) best = i;
    if (best >= 0) {
        void *d = G.pool[best].dev;
        *got = G.pool[best].bytes;
        G.pool[best] = G.pool[--G.npool];
        return d;
    }
    void *d = NULL;
    const double t0 = ops_trace_limit() >= 0.0 ? ops_now_ms() : 0.0;
    if (rfb_dev_alloc(G.ctx, bytes, &d) != RFB_OK) {
        /* out of device memory: drop the pool and retry once */
        for (int i = 0; i < G.npool; i++) rfb_dev_free(G.ctx, G.pool[i].dev);
        G.npool = 0;
        if (rfb_dev_alloc(G.ctx, bytes, &d) != RFB_OK) { set_err("%s", rfb_last_error()); return NULL; }
    }
    *got = bytes;
    ops_trace("device allocation", t0, bytes);
    return d;
}

/* Track a device buffer until the end of the scope (or this call).  host != NULL registers it as the HBM image of that
 * host payload so later operators find it; every older entry with the same pointer is forgotten first. */
static col_entry_t *track(void *dev, size_t bytes, const void *host, int64_t len, int type, int keep) {
    if (G.ncols == G.cap) {
        const int ncap = G.cap ? 2 * G.cap : 64;
        col_entry_t *n = (col_entry_t *)realloc(G.cols, (size_t)ncap * sizeof(col_entry_t));
        if (!n) return NULL;
        G.cols = n;
        G.cap = ncap;
    }
    if (host) { forget_payload(host, NULL); bloom_add(host); }
    col_entry_t *e = &G.cols[G.ncols++];
    e->host = host; e->len = len; e->type = type; e->dev = dev; e->bytes = bytes; e->used = ++G.clock; e->keep = keep;
    e->print = host ? fingerprint(host, (size_t)len * type_size(type)) : 0;
    return e;
}

/* a host vector whose image may outlive the query: it lives in the reference's own heap (mmod 0xff: not a mapped file, which
 * can change or be unmapped behind the hooks) and somebody besides the evaluator's stack holds it (a table column, a global) */
static int worth_keeping(obj_p v) { return G.residency && v->mmod == 0xff && v->rc >= 2; }

/* device image of a host payload (the payload of vector `v`, or a bare array the reference hands over — at_ids' row ids):
 * a hit on a live entry, else one cudaMemcpyAsync */
static void *dev_payload(const void *payload, int64_t len, int type, int keep) {
    const int w = type_size(type);
    if (lazy_on == 1) {
        void *img = lazy_device_image(payload);   /* never read a pending payload: that would fault it in */
        if (img) return img;
        lazy_forget_overlaps((const char *)payload - 16, (const char *)payload + (size_t)len * w, NULL);
    }
    for (int i = G.ncols - 1; i >= 0; i--)
        if (G.cols[i].host == payload) {
            if (G.cols[i].len == len && G.cols[i].type == type && G.cols[i].print == fingerprint(payload, (size_t)len * w)) {
                G.cols[i].used = ++G.clock;
                if (!G.cols[i].keep && keep) G.cols[i].keep = 1;
                G.stat_hits++;
                return G.cols[i].dev;
            }
            forget_payload(payload, NULL);   /* same pointer, other shape or other bytes: the block was reused */
            break;
        }
    size_t got = 0;
    void *d = dev_buffer((size_t)len * w, &got);
    if (!d) return NULL;
    const double t0 = ops_trace_limit() >= 0.0 ? ops_now_ms() : 0.0;
    if (len > 0 && rfb_h2d(G.ctx, d, payload, (size_t)len * w) != RFB_OK) { set_err("%s", rfb_last_error()); rfb_dev_free(G.ctx, d); return NULL; }
    ops_trace("column shipment", t0, (size_t)len * w);
    if (!track(d, got, payload, len, type, keep)) { rfb_sync(G.ctx); rfb_dev_free(G.ctx, d); return NULL; }
    G.stat_ships++;
    return d;
}
static void *dev_column(obj_p v) { return dev_payload(RFB_OBJ_PAYLOAD(v), v->len, v->type, worth_keeping(v)); }

/* is the payload of this vector in HBM already? */
static int is_resident(obj_p v) {
    if (!v || v->type <= 0) return 0;
    const void *payload = RFB_OBJ_PAYLOAD(v);
    if (lazy_on == 1 && lazy_device_image(payload)) return 1;
    for (int i = G.ncols - 1; i >= 0; i--)
        if (G.cols[i].host == payload) return G.cols[i].len == v->len && G.cols[i].type == v->type;
    return 0;
}

/* scratch / result buffer on the device, alive until the end of the scope (or call) */
static void *dev_temp(size_t bytes) {
    size_t got = 0;
    void *d = dev_buffer(bytes, &got);
    if (d && !track(d, got, NULL, 0, 0, 0)) { rfb_dev_free(G.ctx, d); return NULL; }
    return d;
}

static col_entry_t *entry_of_dev(void *dev) {
    for (int i = G.ncols - 1; i >= 0; i--)
        if (G.cols[i].dev == dev) return &G.cols[i];
    return NULL;
}

/* the payload of host vector `r` (len elements of `type`) becomes the content of device buffer `dev`: copied back now, or left
 * lazy; `dev` is registered as the payload's HBM image, whatever lived at that address before is forgotten */
static int fill_host_vector(obj_p r, int type, int64_t len, void *dev) {
    col_entry_t *e = entry_of_dev(dev);
    if (len > 0 && lazy_register(r, (size_t)len * type_size(type), dev)) {
        if (e) { forget_payload(RFB_OBJ_PAYLOAD(r), e); e->host = NULL; e->len = len; e->type = type; }   /* found through the lazy table */
        return RFB_OK;
    }
    if (len > 0) {
        if (rfb_d2h(G.ctx, RFB_OBJ_PAYLOAD(r), dev, (size_t)len * type_size(type)) != RFB_OK || rfb_sync(G.ctx) != RFB_OK) {
            set_err("%s", rfb_last_error());
            return RFB_ERR_CUDA;
        }
    }
    if (e) {
        forget_payload(RFB_OBJ_PAYLOAD(r), e);   /* an older vector that lived at this address */
        e->host = RFB_OBJ_PAYLOAD(r); e->len = len; e->type = type;
        e->keep = G.residency;   /* with the hooks bound a result's image lives until the host frees or rewrites the vector */
        e->print = fingerprint(RFB_OBJ_PAYLOAD(r), (size_t)len * type_size(type));
        bloom_add(e->host);
    }
    return RFB_OK;
}

/* host vector of `type` filled from device memory; the device copy is registered as its HBM image */
static obj_p to_host_vector(int type, int64_t len, void *dev) {
    obj_p r = G.host->vector((int8_t)type, len);
    if (!r || r->type == RFB_T_ERR) return r ? r : G.host->err_limit();
    if (fill_host_vector(r, type, len, dev) != RFB_OK) { G.host->drop_obj(r); return G.host->err_limit(); }
    return r;
}

typedef struct { int entered; } call_scope_t;
static call_scope_t enter(void) { call_scope_t c = {1}; G.scope_depth++; return c; }
static void leave(call_scope_t c) { if (c.entered) rfb_ops_scope_end(); }

static obj_p status_to_obj(int rc) {
    set_err("%s", rfb_last_error());
    switch (rc) {
        case RFB_ERR_TYPE: return G.host->err_type();
        case RFB_ERR_LENGTH: return G.host->err_length();
        default: return G.host->err_limit();
    }
}

/* ------------------------------------------------------------------ operand classification */

static int is_num_type(int t) { return type_size(t) != 0; }
static int is_vec(obj_p o) { return o && o->type > 0 && is_num_type(o->type); }
static int is_atom(obj_p o) { return o && o->type < 0 && is_num_type(-o->type); }
static int too_small(int64_t n) { return n < G.min_rows; }

/* Cost gate for the single-touch element-wise operators (comparisons, arithmetic, round/floor/ceil, not): their result goes
 * back over PCIe, so when NO vector operand is in HBM yet and none is a column worth keeping there (a table column / global
 * that later queries will touch again), shipping operands in and the result out costs more than the reference's CPU loop over
 * the same bytes — such calls are declined.  Inside a query scope everything is taken: the result feeds the next operator on
 * the device (mask -> where -> gather / fold).  RFB200_GATE=0 switches the gate off. */
static int gate_state = -1;
void rfb_ops_set_gate(int on) { gate_state = on ? 1 : 0; }
static int gated_out(obj_p x, obj_p y) {
    if (gate_state < 0) { const char *e = getenv("RFB200_GATE"); gate_state = (e && e[0] == '0') ? 0 : 1; }
    if (!gate_state || G.scope_depth > 0) return 0;
    const int xv = x && x->type > 0, yv = y && y->type > 0;
    if ((xv && (is_resident(x) || worth_keeping(x))) || (yv && (is_resident(y) || worth_keeping(y)))) return 0;
    return 1;
}

static rfb_scalar_t scalar_of(obj_p a) {
    rfb_scalar_t s;
    memset(&s, 0, sizeof(s));
    s.type = -a->type;
    switch (type_size(s.type)) {
        case 1: s.v.u8 = a->u8; break;
        case 2: s.v.i16 = a->i16; break;
        case 4: s.v.i32 = a->i32; break;
        default: s.v.i64 = a->i64; break;
    }
    return s;
}

static obj_p make_atom(int type, const void *bits) {
    obj_p a = G.host->atom((int8_t)type);
    if (!a) return G.host->err_limit();
    a->i64 = 0;
    memcpy(&a->i64, bits, (size_t)type_size(type));
    return a;
}

/* ------------------------------------------------------------------ comparisons */

static obj_p cmp_op(int op, obj_p x, obj_p y) {
    if (!G.ready) return NULL;
    const int xv = is_vec(x), yv = is_vec(y);
    if (!(xv || yv) || !((xv || is_atom(x)) && (yv || is_atom(y)))) return NULL; /* atoms only, lists, tables, enums... */
    const int64_t n = xv ? x->len : y->len;
    if (xv && yv && x->len != y->len) return G.host->err_length();
    if (too_small(n) || gated_out(x, y)) return NULL;
    call_scope_t sc = enter();
    obj_p res = NULL;
    const int xt = xv ? x->type : -x->type, yt = yv ? y->type : -y->type;
    rfb_scalar_t xs, ys;
    void *dx = NULL, *dy = NULL;
    if (xv) { dx = dev_column(x); if (!dx) { res = G.host->err_limit(); goto out; } } else xs = scalar_of(x);
    if (yv) { dy = dev_column(y); if (!dy) { res = G.host->err_limit(); goto out; } } else ys = scalar_of(y);
    void *dm = dev_temp((size_t)n);
    if (!dm) { res = G.host->err_limit(); goto out; }
    int rc = rfb_cmp_dev(G.ctx, op, xt, dx, xv ? n : -1, xv ? NULL : &xs, yt, dy, yv ? n : -1, yv ? NULL : &ys, (uint8_t *)dm);
    if (rc) { res = status_to_obj(rc); goto out; }
    res = to_host_vector(RFB_T_B8, n, dm);
out:
    leave(sc);
    return res;
}
obj_p rfb_ray_eq(obj_p x, obj_p y) { return cmp_op(RFB_EQ, x, y); }
obj_p rfb_ray_ne(obj_p x, obj_p y) { return cmp_op(RFB_NE, x, y); }
obj_p rfb_ray_lt(obj_p x, obj_p y) { return cmp_op(RFB_LT, x, y); }
obj_p rfb_ray_gt(obj_p x, obj_p y) { return cmp_op(RFB_GT, x, y); }
obj_p rfb_ray_le(obj_p x, obj_p y) { return cmp_op(RFB_LE, x, y); }
obj_p rfb_ray_ge(obj_p x, obj_p y) { return cmp_op(RFB_GE, x, y); }

/* ------------------------------------------------------------------ where / filter */

obj_p rfb_ray_where(obj_p mask) {
    if (!G.ready || !mask) return NULL;
    if (mask->type != RFB_T_B8) return (mask->type > 0 && is_num_type(mask->type)) ? G.host->err_type() : NULL;
    if (too_small(mask->len)) return NULL;
    call_scope_t sc = enter();
    obj_p res;
    const int64_t n = mask->len;
    void *dm = dev_column(mask), *di = dev_temp((size_t)(n > 0 ? n : 1) * 8);
    int64_t cnt = 0;
    if (!dm || !di) { res = G.host->err_limit(); goto out; }
    int rc = rfb_where_dev(G.ctx, (const uint8_t *)dm, n, (int64_t *)di, &cnt);
    if (rc) { res = status_to_obj(rc); goto out; }
    res = to_host_vector(RFB_T_I64, cnt, di);
out:
    leave(sc);
    return res;
}

static obj_p pair_of(obj_p a, obj_p b, int type) {
    obj_p r = G.host->vector(RFB_T_LIST, 2);
    if (!r || r->type == RFB_T_ERR) return r ? r : G.host->err_limit();
    RFB_OBJ_LIST(r)[0] = G.host->clone_obj(a);
    RFB_OBJ_LIST(r)[1] = G.host->clone_obj(b);
    r->type = (int8_t)type;
    return r;
}
obj_p rfb_filter_map(obj_p val, obj_p index) { return (G.ready && is_vec(val) && index) ? pair_of(val, index, RFB_T_MAPFILTER) : NULL; }
obj_p rfb_group_map(obj_p val, obj_p index) { return (G.ready && is_vec(val) && index) ? pair_of(val, index, RFB_T_MAPGROUP) : NULL; }

obj_p rfb_filter_collect(obj_p val, obj_p index) {
    if (!G.ready || !is_vec(val) || !index || index->type != RFB_T_I64) return NULL;
    if (too_small(index->len)) return NULL;
    call_scope_t sc = enter();
    obj_p res;
    const int64_t m = index->len;
    void *dc = dev_column(val), *di = dev_column(index), *dout = dev_temp((size_t)(m > 0 ? m : 1) * type_size(val->type));
    if (!dc || !di || !dout) { res = G.host->err_limit(); goto out; }
    int rc = rfb_gather_dev(G.ctx, val->type, dc, (const int64_t *)di, m, dout);
    if (rc) { res = status_to_obj(rc); goto out; }
    res = to_host_vector(val->type, m, dout);
out:
    leave(sc);
    return res;
}

/* ------------------------------------------------------------------ ungrouped folds */

enum { F_SUM, F_MIN, F_MAX, F_AVG, F_CNT };

/* result atom typing: ray_sum_partial core/math.c:1850-1871, ray_min/max :1922-2045, ray_avg :2445-2526 */
static obj_p fold_result(int what, int type, const rfb_fold_t *f) {
    const int flt = (type == RFB_T_F64);
    switch (what) {
        case F_CNT: return make_atom(RFB_T_I64, &f->nonnull);
        case F_SUM:
            if (flt) return make_atom(RFB_T_F64, &f->sum_f64);
            if (type == RFB_T_I32 || type == RFB_T_TIME) { int32_t s = (int32_t)f->sum_i64; return make_atom(type, &s); }
            return make_atom(RFB_T_I64, &f->sum_i64);
        case F_MIN: case F_MAX: {
            if (flt) return make_atom(RFB_T_F64, what == F_MIN ? &f->min_f64 : &f->max_f64);
            int64_t v = what == F_MIN ? f->min_i64 : f->max_i64;
            return make_atom(type, &v); /* little endian: the low bytes are the narrow value */
        }
        default: { /* avg = sum / non-null count, 0Nf when nothing was counted (FDIVI64 / FDIVF64 core/ops.h:173-174) */
            double r;
            if (f->nonnull == 0) r = NAN;
            else if (flt) r = f->sum_f64 / (double)f->nonnull;
            else r = (double)f->sum_i64 / (double)f->nonnull;
            return make_atom(RFB_T_F64, &r);
        }
    }
}

static int fold_type_ok(int what, int type) {
    switch (what) {
        case F_SUM: return type == RFB_T_U8 || type == RFB_T_I16 || type == RFB_T_I32 || type == RFB_T_I64 || type == RFB_T_F64 || type == RFB_T_TIME;
        case F_AVG: return type == RFB_T_U8 || type == RFB_T_I16 || type == RFB_T_I32 || type == RFB_T_I64 || type == RFB_T_F64;
        case F_CNT: return type != RFB_T_B8 && type != RFB_T_SYMBOL;
        default: return type != RFB_T_B8 && type != RFB_T_SYMBOL;
    }
}

static obj_p aggr_op(int op, obj_p val, obj_p index);

static obj_p fold_op(int what, obj_p x) {
    if (!G.ready || !x) return NULL;
    if (x->type == RFB_T_MAPGROUP && x->len == 2) {
        static const int A[] = {RFB_A_SUM, RFB_A_MIN, RFB_A_MAX, RFB_A_AVG, -1};
        return A[what] < 0 ? NULL : aggr_op(A[what], RFB_OBJ_LIST(x)[0], RFB_OBJ_LIST(x)[1]);
    }
    obj_p col = x, ids = NULL;
    if (x->type == RFB_T_MAPFILTER && x->len == 2) { col = RFB_OBJ_LIST(x)[0]; ids = RFB_OBJ_LIST(x)[1]; if (!ids || ids->type != RFB_T_I64) return NULL; }
    if (!is_vec(col)) return NULL;
    if (!fold_type_ok(what, col->type)) return G.host->err_type();
    const int64_t n = ids ? ids->len : col->len;
    if (too_small(n)) return NULL;
    call_scope_t sc = enter();
    obj_p res;
    rfb_fold_t f;
    const int folds = (what == F_MIN || what == F_MAX) ? (RFB_F_MIN | RFB_F_MAX) : (RFB_F_SUM | RFB_F_CNT);
    void *dc = dev_column(col), *di = ids ? dev_column(ids) : NULL;
    if (!dc || (ids && !di)) { res = G.host->err_limit(); goto out; }
    int rc = ids ? rfb_gather_fold_dev(G.ctx, folds, col->type, dc, (const int64_t *)di, n, &f)
                 : rfb_fold_dev(G.ctx, folds, col->type, dc, n, &f);
    if (rc) { res = status_to_obj(rc); goto out; }
    res = fold_result(what, col->type, &f);
out:
    leave(sc);
    return res;
}
obj_p rfb_ray_sum(obj_p x) { return fold_op(F_SUM, x); }
obj_p rfb_ray_min(obj_p x) { return fold_op(F_MIN, x); }
obj_p rfb_ray_max(obj_p x) { return fold_op(F_MAX, x); }
obj_p rfb_ray_avg(obj_p x) { return fold_op(F_AVG, x); }
obj_p rfb_ray_cnt(obj_p x) { return fold_op(F_CNT, x); }

/* ---- ray_med / ray_dev (core/math.c:2529-2700): a plain vector, a MAPFILTER pair (the reference collects it first, so does
 *      the device: gather, then the statistic) or a MAPGROUP pair (-> aggr_med / aggr_dev).  Atoms stay on the CPU body. */
static obj_p stat_op(int is_dev, obj_p x) {
    if (!G.ready || !x) return NULL;
    if (x->type == RFB_T_MAPGROUP && x->len == 2) return aggr_op(is_dev ? RFB_A_DEV : RFB_A_MED, RFB_OBJ_LIST(x)[0], RFB_OBJ_LIST(x)[1]);
    obj_p col = x, ids = NULL;
    if (x->type == RFB_T_MAPFILTER && x->len == 2) { col = RFB_OBJ_LIST(x)[0]; ids = RFB_OBJ_LIST(x)[1]; if (!ids || ids->type != RFB_T_I64) return NULL; }
    if (!is_vec(col)) return NULL;
    const int t = col->type;
    if (is_dev) {
        if (t == RFB_T_B8 || t == RFB_T_SYMBOL) return G.host->err_type();
        if (t == RFB_T_DATE || t == RFB_T_TIMESTAMP) return NULL;   /* the reference reads a field of ray_sum's error object here: its body, its result */
    } else if (!(t == RFB_T_U8 || t == RFB_T_I16 || t == RFB_T_I64)) return G.host->err_type();   /* core/math.c:2555-2590 */
    const int64_t n = ids ? ids->len : col->len;
    if (too_small(n)) return NULL;
    call_scope_t sc = enter();
    obj_p res;
    double r = 0.0;
    void *dc = dev_column(col), *di = ids ? dev_column(ids) : NULL, *dx = dc;
    if (!dc || (ids && !di)) { res = G.host->err_limit(); goto out; }
    int rc = RFB_OK;
    if (ids) {
        dx = dev_temp((size_t)(n > 0 ? n : 1) * type_size(t));
        if (!dx) { res = G.host->err_limit(); goto out; }
        rc = rfb_gather_dev(G.ctx, t, dc, (const int64_t *)di, n, dx);
    }
    if (!rc) rc = is_dev ? rfb_stddev_dev(G.ctx, t, dx, n, &r) : rfb_med_dev(G.ctx, t, dx, n, &r);
    if (rc) { res = status_to_obj(rc); goto out; }
    res = make_atom(RFB_T_F64, &r);
out:
    leave(sc);
    return res;
}
obj_p rfb_ray_med(obj_p x) { return stat_op(0, x); }
obj_p rfb_ray_dev(obj_p x) { return stat_op(1, x); }

/* ------------------------------------------------------------------ element-wise */

static obj_p bin_op(int op, obj_p x, obj_p y) {
    if (!G.ready) return NULL;
    const int xv = is_vec(x), yv = is_vec(y);
    if (!(xv || yv) || !((xv || is_atom(x)) && (yv || is_atom(y)))) return NULL;
    const int xt = xv ? x->type : -x->type, yt = yv ? y->type : -y->type;
    /* the reference's full type matrix (core/math.c:251-1782): U8 / I16 / DATE / TIME / TIMESTAMP operands included; where it has
     * no case (DATE * DATE ...) its own body raises the type error */
    const int ot = rfb_binop_type_form(op, xv ? (yv ? 0 : 1) : 2, xt, yt);
    if (ot < 0) return NULL;
    if (xv && yv && x->len != y->len) return G.host->err_length();
    const int64_t n = xv ? x->len : y->len;
    if (too_small(n) || gated_out(x, y)) return NULL;
    call_scope_t sc = enter();
    obj_p res;
    rfb_scalar_t xs, ys;
    void *dx = NULL, *dy = NULL;
    if (xv) { dx = dev_column(x); if (!dx) { res = G.host->err_limit(); goto out; } } else xs = scalar_of(x);
    if (yv) { dy = dev_column(y); if (!dy) { res = G.host->err_limit(); goto out; } } else ys = scalar_of(y);
    void *dout = dev_temp((size_t)(n > 0 ? n : 1) * type_size(ot));
    if (!dout) { res = G.host->err_limit(); goto out; }
    int rc = rfb_binop_dev(G.ctx, op, xt, dx, xv ? n : -1, xv ? NULL : &xs, yt, dy, yv ? n : -1, yv ? NULL : &ys, dout);
    if (rc) { res = status_to_obj(rc); goto out; }
    res = to_host_vector(ot, n, dout);
out:
    leave(sc);
    return res;
}
obj_p rfb_ray_add(obj_p x, obj_p y) { return bin_op(RFB_ADD, x, y); }
obj_p rfb_ray_sub(obj_p x, obj_p y) { return bin_op(RFB_SUB, x, y); }
obj_p rfb_ray_mul(obj_p x, obj_p y) { return bin_op(RFB_MUL, x, y); }
obj_p rfb_ray_div(obj_p x, obj_p y) { return bin_op(RFB_DIV, x, y); }
obj_p rfb_ray_fdiv(obj_p x, obj_p y) { return bin_op(RFB_FDIV, x, y); }
obj_p rfb_ray_mod(obj_p x, obj_p y) { return bin_op(RFB_MOD, x, y); }
obj_p rfb_ray_xbar(obj_p x, obj_p y) { return bin_op(RFB_XBAR, x, y); }

static obj_p un_op(int op, obj_p x) {
    if (!G.ready || !x || x->type != RFB_T_F64) return NULL; /* integer / temporal inputs are returned as-is by the CPU body */
    if (too_small(x->len) || gated_out(x, NULL)) return NULL;
    call_scope_t sc = enter();
    obj_p res;
    void *dx = dev_column(x), *dout = dev_temp((size_t)(x->len > 0 ? x->len : 1) * 8);
    if (!dx || !dout) { res = G.host->err_limit(); goto out; }
    int rc = rfb_unop_f64_dev(G.ctx, op, (const double *)dx, x->len, (double *)dout);
    if (rc) { res = status_to_obj(rc); goto out; }
    res = to_host_vector(RFB_T_F64, x->len, dout);
out:
    leave(sc);
    return res;
}
obj_p rfb_ray_round(obj_p x) { return un_op(RFB_ROUND, x); }
obj_p rfb_ray_floor(obj_p x) { return un_op(RFB_FLOOR, x); }
obj_p rfb_ray_ceil(obj_p x) { return un_op(RFB_CEIL, x); }

/* ------------------------------------------------------------------ group-by */

static int is_null_obj(obj_p o) { return !o || o == G.host->null_obj || o->type == RFB_T_NULL; }

static obj_p i64_atom(int64_t v) { return make_atom(RFB_T_I64, &v); }

obj_p rfb_index_group(obj_p keys, obj_p filter) {
    if (!G.ready || !keys) return NULL;
    /* index_group has no I32/DATE/TIME case (core/index.c:2177-2224): single-key grouping exists for I64-kind keys */
    if (!(keys->type == RFB_T_I64 || keys->type == RFB_T_SYMBOL || keys->type == RFB_T_TIMESTAMP)) return NULL;
    const int filtered = !is_null_obj(filter);
    if (filtered && filter->type != RFB_T_I64) return NULL; /* parted filters */
    const int64_t len = filtered ? filter->len : keys->len;
    if (too_small(len)) return NULL;
    call_scope_t sc = enter();
    obj_p res = NULL, gids = NULL, firsts = NULL;
    rfb_group_info_t info;
    void *dk = dev_column(keys), *df = filtered ? dev_column(filter) : NULL;
    void *dg = dev_temp((size_t)(len > 0 ? len : 1) * 8), *dfi = dev_temp((size_t)(len > 0 ? len : 1) * 8);
    if (!dk || (filtered && !df) || !dg || !dfi) { res = G.host->err_limit(); goto out; }
    int rc = rfb_group_i64_dev(G.ctx, (const int64_t *)dk, (const int64_t *)df, len, (int64_t *)dg, (int64_t *)dfi, &info);
    if (rc) { res = status_to_obj(rc); goto out; }
    gids = to_host_vector(RFB_T_I64, len, dg);
    firsts = to_host_vector(RFB_T_I64, info.groups, dfi);
    res = G.host->vector(RFB_T_LIST, 7);
    if (!res || res->type == RFB_T_ERR || !gids || !firsts || gids->type == RFB_T_ERR || firsts->type == RFB_T_ERR) {
        if (res && res->type != RFB_T_ERR) { res->len = 0; G.host->drop_obj(res); }
        if (gids && gids->type != RFB_T_ERR) G.host->drop_obj(gids);
        if (firsts && firsts->type != RFB_T_ERR) G.host->drop_obj(firsts);
        res = G.host->err_limit();
        goto out;
    }
    RFB_OBJ_LIST(res)[0] = i64_atom(RFB_INDEX_IDS);
    RFB_OBJ_LIST(res)[1] = i64_atom(info.groups);
    RFB_OBJ_LIST(res)[2] = gids;
    RFB_OBJ_LIST(res)[3] = i64_atom(RFB_NULL_I64);
    RFB_OBJ_LIST(res)[4] = G.host->null_obj;
    RFB_OBJ_LIST(res)[5] = filtered ? G.host->clone_obj(filter) : G.host->null_obj;
    RFB_OBJ_LIST(res)[6] = firsts;
out:
    leave(sc);
    return res;
}

/* ---- parted columns: PARTED_MAP (core/aggr.c:183-260) and aggr_avg's parted branch (:2065-2127).
 * val = list of per-partition vectors; index = [PARTEDCOMMON, groups, -, -, -, filter, -].  Partition values follow the GROUPED
 * aggregate semantics (aggr_*_partial with one group): sum is sticky-null in the value's own width, min starts at +INF,
 * max at null, avg = (f64 sum of non-nulls, their count).  groups == 1: combined with ADD (sticky) / MIN / MAX (null-skipping)
 * / summed totals; otherwise one result per included partition. */
#define RFB_INDEX_PARTEDCOMMON 2
typedef struct { int64_t i; double f; int64_t cnt; } part_val_t;

static int parted_fold(int op, int t, obj_p part, obj_p ids, part_val_t *out) {
    rfb_fold_t f;
    const int folds = (op == RFB_A_MIN || op == RFB_A_MAX) ? (RFB_F_MIN | RFB_F_MAX) : (RFB_F_SUM | RFB_F_CNT);
    const int64_t rows = ids ? ids->len : part->len;
    memset(&f, 0, sizeof(f));
    if (rows > 0) {
        void *dc = dev_column(part), *di = ids ? dev_column(ids) : NULL;
        if (!dc || (ids && !di)) return RFB_ERR_NOMEM;
        int rc = ids ? rfb_gather_fold_dev(G.ctx, folds, t, dc, (const int64_t *)di, rows, &f) : rfb_fold_dev(G.ctx, folds, t, dc, rows, &f);
        if (rc) return rc;
    }
    const int flt = (t == RFB_T_F64), w = type_size(t);
    const int64_t null_i = w == 8 ? RFB_NULL_I64 : (w == 4 ? (int64_t)INT32_MIN : (int64_t)INT16_MIN);
    const int64_t inf_i = w == 8 ? INT64_MAX : (w == 4 ? (int64_t)INT32_MAX : (int64_t)INT16_MAX);
    out->cnt = f.nonnull;
    switch (op) {
        case RFB_A_SUM:   /* sticky: one null makes the partition's sum null */
            if (flt) out->f = f.nonnull < rows ? NAN : (rows ? f.sum_f64 : 0.0);
            else if (f.nonnull < rows) out->i = null_i;
            else out->i = w == 8 ? f.sum_i64 : (w == 4 ? (int64_t)(int32_t)f.sum_i64 : (int64_t)(int16_t)f.sum_i64);
            break;
        case RFB_A_MIN:
            if (flt) out->f = f.nonnull ? f.min_f64 : INFINITY; else out->i = f.nonnull ? f.min_i64 : inf_i;
            break;
        case RFB_A_MAX:
            if (flt) out->f = f.nonnull ? f.max_f64 : NAN; else out->i = f.nonnull ? f.max_i64 : null_i;
            break;
        default:          /* avg: f64 sum of the non-null values */
            out->f = flt ? (f.nonnull ? f.sum_f64 : 0.0) : (double)f.sum_i64;
            break;
    }
    return RFB_OK;
}

static obj_p parted_aggr(int op, obj_p val, obj_p index) {
    const int t = val->type - RFB_T_PARTED;
    if (!(t == RFB_T_I16 || t == RFB_T_I32 || t == RFB_T_I64 || t == RFB_T_F64 || t == RFB_T_DATE || t == RFB_T_TIME || t == RFB_T_TIMESTAMP)) return NULL;
    if (!(op == RFB_A_SUM || op == RFB_A_MIN || op == RFB_A_MAX || op == RFB_A_AVG)) return NULL;   /* count reads no data; med/dev: CPU body */
    /* the per-partition partials have no case for these (aggr_sum_partial core/aggr.c:1082-1104, aggr_min/max_partial :1156-1260,
     * aggr_avg :2021-2027): whatever the reference does with the partial's error stays the CPU body's business */
    if ((op == RFB_A_AVG || op == RFB_A_SUM) && t == RFB_T_TIMESTAMP) return NULL;
    if ((op == RFB_A_MIN || op == RFB_A_MAX) && t == RFB_T_I32) return NULL;
    obj_p *ix = RFB_OBJ_LIST(index);
    const int64_t n = ix[1]->i64, l = val->len;
    obj_p filter = is_null_obj(ix[5]) ? NULL : ix[5];
    const int idfilter = filter && filter->type == RFB_T_PARTED + RFB_T_I64;
    if (filter && filter->len != l) return NULL;
    if (op == RFB_A_AVG && filter) return NULL;                  /* aggr_avg hands the filter to its partials: CPU body */
    if (!filter && n != 1 && n != l) return NULL;
    int64_t total_rows = 0;
    for (int64_t i = 0; i < l; i++) {
        obj_p p = RFB_OBJ_LIST(val)[i];
        if (!p || p->type != t) return NULL;
        total_rows += p->len;
    }
    if (too_small(total_rows)) return NULL;
    const int flt = (t == RFB_T_F64), w = type_size(t);
    const int ot = op == RFB_A_AVG ? RFB_T_F64 : (flt ? RFB_T_F64 : (w == 8 ? RFB_T_I64 : (w == 4 ? RFB_T_I32 : RFB_T_I16)));   /* __v_i64 / __v_i32 / ... */
    const int64_t null_i = w == 8 ? RFB_NULL_I64 : (w == 4 ? (int64_t)INT32_MIN : (int64_t)INT16_MIN);
    call_scope_t sc = enter();
    obj_p res = NULL;
    part_val_t *pv = (part_val_t *)malloc((size_t)(l > 0 ? l : 1) * sizeof(part_val_t));
    int64_t m = 0;                                               /* included partitions */
    if (!pv) { res = G.host->err_limit(); goto out; }
    for (int64_t i = 0; i < l; i++) {
        obj_p ids = NULL;
        if (filter) {
            obj_p fe = RFB_OBJ_LIST(filter)[i];
            if (is_null_obj(fe)) continue;
            if (idfilter) {
                if (fe->type > 0 && fe->len == 0) continue;
                if (!(fe->type == -RFB_T_I64 && fe->i64 == -1)) { if (fe->type != RFB_T_I64) { res = NULL; goto out; } ids = fe; }
            }
        }
        int rc = parted_fold(op, t, RFB_OBJ_LIST(val)[i], ids, &pv[m]);
        if (rc) { res = status_to_obj(rc); goto out; }
        m++;
    }
    const int combine = (n == 1) && (!filter || idfilter);
    if (!combine && filter && !idfilter && n == 1 && m > 1) { res = NULL; goto out; }      /* the reference would overrun a 1-element result here */
    const int64_t outn = combine ? 1 : m;
    res = G.host->vector((int8_t)ot, combine ? 1 : n);
    if (!res || res->type == RFB_T_ERR) { res = G.host->err_limit(); goto out; }
    if (!combine && outn > n) { G.host->drop_obj(res); res = NULL; goto out; }
    char *o = (char *)RFB_OBJ_PAYLOAD(res);
    if (combine) {
        part_val_t acc;
        memset(&acc, 0, sizeof(acc));
        int64_t cnt = 0;
        double fsum = 0.0;
        for (int64_t j = 0; j < m; j++) {
            const part_val_t *v = &pv[j];
            if (op == RFB_A_AVG) { fsum += v->f; cnt += v->cnt; continue; }
            if (j == 0) { acc = *v; continue; }
            if (flt) {
                const int an = isnan(acc.f), bn = isnan(v->f);
                if (op == RFB_A_SUM) acc.f = (an || bn) ? NAN : acc.f + v->f;
                else if (op == RFB_A_MIN) acc.f = an ? v->f : (bn ? acc.f : (acc.f < v->f ? acc.f : v->f));
                else acc.f = an ? v->f : (bn ? acc.f : (acc.f > v->f ? acc.f : v->f));
            } else {
                const int an = acc.i == null_i, bn = v->i == null_i;
                if (op == RFB_A_SUM) {
                    if (an || bn) acc.i = null_i;
                    else { const uint64_t s = (uint64_t)acc.i + (uint64_t)v->i; acc.i = w == 8 ? (int64_t)s : (w == 4 ? (int64_t)(int32_t)s : (int64_t)(int16_t)s); }
                } else if (op == RFB_A_MIN) acc.i = an ? v->i : (bn ? acc.i : (acc.i < v->i ? acc.i : v->i));
                else acc.i = an ? v->i : (bn ? acc.i : (acc.i > v->i ? acc.i : v->i));
            }
        }
        if (op == RFB_A_AVG) { const double a = cnt == 0 ? NAN : fsum / (double)cnt; memcpy(o, &a, 8); }
        else if (m == 0) memset(o, 0, (size_t)type_size(ot));   /* (no partition included: the reference leaves the slot as allocated) */
        else if (flt) memcpy(o, &acc.f, 8);
        else memcpy(o, &acc.i, (size_t)w);
    } else {
        for (int64_t j = 0; j < m; j++) {
            if (op == RFB_A_AVG) { const double a = pv[j].cnt == 0 ? NAN : pv[j].f / (double)pv[j].cnt; memcpy(o + j * 8, &a, 8); }
            else if (flt) memcpy(o + j * 8, &pv[j].f, 8);
            else memcpy(o + j * w, &pv[j].i, (size_t)w);
        }
    }
out:
    free(pv);
    leave(sc);
    return res;
}

/* aggr_* over the WINDOW index of a window join (index_window_join_obj, core/index.c:3287-3346: [WINDOW, left rows, left time,
 * right time, [window lo, window hi], LIST of one [first, last] vector per left row or null, jtype]): AGGR_ITER's WINDOW branch
 * (core/aggr.c:131-160) on the device.  The per-row block bounds are flattened into two arrays on the host (the index keeps
 * them as one small object per row), everything else is column payloads. */
#define RFB_INDEX_WINDOW 3
static obj_p window_aggr(int op, obj_p val, obj_p index) {
    if (!(op == RFB_A_SUM || op == RFB_A_MIN || op == RFB_A_MAX || op == RFB_A_COUNT || op == RFB_A_AVG)) return NULL;
    obj_p *ix = RFB_OBJ_LIST(index);
    if (!ix[1] || ix[1]->type != -RFB_T_I64 || !ix[3] || !ix[4] || !ix[5] || !ix[6] || ix[6]->type != -RFB_T_I64) return NULL;
    obj_p rtime = ix[3], win = ix[4], blocks = ix[5];
    const int64_t ll = ix[1]->i64, jtype = ix[6]->i64;
    if (rtime->type <= 0 || type_size(rtime->type) != 4 || win->type != RFB_T_LIST || win->len != 2 || blocks->type != RFB_T_LIST || blocks->len != ll) return NULL;
    obj_p wlo = RFB_OBJ_LIST(win)[0], whi = RFB_OBJ_LIST(win)[1];
    if (!wlo || !whi || wlo->type <= 0 || whi->type <= 0 || type_size(wlo->type) != 4 || type_size(whi->type) != 4 || wlo->len != ll || whi->len != ll) return NULL;
    if (!(jtype == 0 || jtype == 1) || !is_vec(val) || val->len != rtime->len) return NULL;
    const int vt = val->type;
    if (!(vt == RFB_T_I64 || vt == RFB_T_TIMESTAMP || vt == RFB_T_F64)) return NULL;   /* other value types: CPU body */
    const int ot = rfb_aggr_type(op, vt);
    if (ot < 0) return G.host->err_type();
    if (too_small(ll) || ll == 0) return NULL;
    int64_t *fl = (int64_t *)malloc((size_t)ll * 16);
    if (!fl) return G.host->err_limit();
    for (int64_t i = 0; i < ll; i++) {
        obj_p b = RFB_OBJ_LIST(blocks)[i];
        if (is_null_obj(b)) { fl[i] = RFB_NULL_I64; fl[ll + i] = RFB_NULL_I64; continue; }
        if (b->type != RFB_T_I64 || b->len != 2) { free(fl); return NULL; }
        fl[i] = ((const int64_t *)RFB_OBJ_PAYLOAD(b))[0];
        fl[ll + i] = ((const int64_t *)RFB_OBJ_PAYLOAD(b))[1];
    }
    call_scope_t sc = enter();
    obj_p res;
    void *dv = dev_column(val), *dt = dev_column(rtime), *dlo = dev_column(wlo), *dhi = dev_column(whi);
    void *dfl = dev_temp((size_t)ll * 16), *dout = dev_temp((size_t)ll * 8);
    if (!dv || !dt || !dlo || !dhi || !dfl || !dout || rfb_h2d(G.ctx, dfl, fl, (size_t)ll * 16) != RFB_OK) { res = G.host->err_limit(); goto out; }
    int rc = rfb_window_aggr_dev(G.ctx, (const int32_t *)dt, (const int64_t *)dfl, (const int64_t *)dfl + ll, ll, (const int32_t *)dlo, (const int32_t *)dhi,
                                 (int)jtype, op, vt, dv, dout);
    if (!rc) rc = rfb_sync(G.ctx);                     /* fl is read by an asynchronous copy */
    if (rc) { res = status_to_obj(rc); goto out; }
    res = to_host_vector(ot, ll, dout);
out:
    leave(sc);
    free(fl);
    return res;
}

static obj_p aggr_op(int op, obj_p val, obj_p index) {
    if (G.ready && val && index && index->type == RFB_T_LIST && index->len == 7 && val->type > RFB_T_PARTED && val->type <= RFB_T_PARTED + RFB_T_F64) {
        obj_p *px = RFB_OBJ_LIST(index);
        if (px[0] && px[0]->type == -RFB_T_I64 && px[0]->i64 == RFB_INDEX_PARTEDCOMMON && px[1] && px[1]->type == -RFB_T_I64) return parted_aggr(op, val, index);
        return NULL;
    }
    if (!G.ready || !is_vec(val) || !index || index->type != RFB_T_LIST || index->len != 7) return NULL;
    obj_p *ix = RFB_OBJ_LIST(index);
    if (ix[0] && ix[0]->type == -RFB_T_I64 && ix[0]->i64 == RFB_INDEX_WINDOW) return window_aggr(op, val, index);
    if (!ix[0] || ix[0]->type != -RFB_T_I64 || ix[0]->i64 != RFB_INDEX_IDS) return NULL; /* SHIFT / parted indices: CPU body */
    if (!ix[1] || ix[1]->type != -RFB_T_I64) return NULL;
    obj_p gids = ix[2], filter = ix[5];
    if (!gids || gids->type != RFB_T_I64) return NULL;
    const int filtered = !is_null_obj(filter);
    if (filtered && filter->type != RFB_T_I64) return NULL;
    const int64_t groups = ix[1]->i64, len = gids->len;
    const int ot = rfb_aggr_type(op, val->type);
    if (ot < 0) return G.host->err_type();
    if (too_small(len)) return NULL;
    call_scope_t sc = enter();
    obj_p res;
    void *dv = dev_column(val), *dg = dev_column(gids), *df = filtered ? dev_column(filter) : NULL;
    void *dout = dev_temp((size_t)(groups > 0 ? groups : 1) * 8);
    if (!dv || !dg || (filtered && !df) || !dout) { res = G.host->err_limit(); goto out; }
    if (ops_trace_limit() >= 0.0) { const double tp = ops_now_ms(); rfb_sync(G.ctx); ops_trace("pending before aggr", tp, 0); }
    const double t0 = ops_trace_limit() >= 0.0 ? ops_now_ms() : 0.0;
    int rc = rfb_aggr_dev(G.ctx, op, val->type, dv, (const int64_t *)df, (const int64_t *)dg, len, groups, dout);
    if (rc) { res = status_to_obj(rc); goto out; }
    if (ops_trace_limit() >= 0.0) { rfb_sync(G.ctx); ops_trace("rfb_aggr_dev", t0, (size_t)len * 8); }
    res = to_host_vector(ot, groups, dout);
out:
    leave(sc);
    return res;
}
obj_p rfb_aggr_sum(obj_p v, obj_p i) { return aggr_op(RFB_A_SUM, v, i); }
obj_p rfb_aggr_min(obj_p v, obj_p i) { return aggr_op(RFB_A_MIN, v, i); }
obj_p rfb_aggr_max(obj_p v, obj_p i) { return aggr_op(RFB_A_MAX, v, i); }
obj_p rfb_aggr_count(obj_p v, obj_p i) { return aggr_op(RFB_A_COUNT, v, i); }
obj_p rfb_aggr_avg(obj_p v, obj_p i) { return aggr_op(RFB_A_AVG, v, i); }
obj_p rfb_aggr_med(obj_p v, obj_p i) { return aggr_op(RFB_A_MED, v, i); }
obj_p rfb_aggr_stddev(obj_p v, obj_p i) { return aggr_op(RFB_A_DEV, v, i); }   /* aggr_dev; rfb_aggr_dev is the C ABI's device-layer entry */

/* aggr_row / aggr_collect (core/aggr.c:3021-3136): a LIST with one vector per group — the row ids, or the values, of the
 * group's rows in row order.  The device orders the rows by group (one stable sort of the group ids), the host slices. */
static obj_p group_lists(int collect, obj_p val, obj_p index) {
    if (!G.ready || !index || index->type != RFB_T_LIST || index->len != 7) return NULL;
    if (collect && !is_vec(val)) return NULL;                  /* ENUM / GUID / LIST / parted columns: CPU body */
    obj_p *ix = RFB_OBJ_LIST(index);
    if (!ix[0] || ix[0]->type != -RFB_T_I64 || ix[0]->i64 != RFB_INDEX_IDS) return NULL;
    obj_p gids = ix[2], filter = ix[5];
    if (!gids || gids->type != RFB_T_I64) return NULL;
    const int filtered = !is_null_obj(filter);
    if (filtered && filter->type != RFB_T_I64) return NULL;
    const int64_t groups = ix[1]->i64, len = gids->len;
    if (too_small(len)) return NULL;
    const int ot = collect ? val->type : RFB_T_I64, w = type_size(ot);
    call_scope_t sc = enter();
    obj_p res = NULL;
    int64_t *offs = NULL;
    char *flat = NULL;
    void *dg = dev_column(gids), *df = filtered ? dev_column(filter) : NULL, *dv = collect ? dev_column(val) : NULL;
    void *drows = dev_temp((size_t)(len > 0 ? len : 1) * 8), *doffs = dev_temp((size_t)(groups + 1) * 8);
    void *dout = collect ? dev_temp((size_t)(len > 0 ? len : 1) * w) : drows;
    if (!dg || (filtered && !df) || (collect && !dv) || !drows || !doffs || !dout) { res = G.host->err_limit(); goto out; }
    int rc = rfb_group_rows_dev(G.ctx, (const int64_t *)dg, (const int64_t *)df, len, groups, (int64_t *)drows, (int64_t *)doffs);
    if (!rc && collect) rc = rfb_gather_dev(G.ctx, ot, dv, (const int64_t *)drows, len, dout);
    if (rc) { res = status_to_obj(rc); goto out; }
    offs = (int64_t *)malloc((size_t)(groups + 1) * 8);
    flat = (char *)malloc((size_t)(len > 0 ? len : 1) * w);
    if (!offs || !flat || rfb_d2h(G.ctx, offs, doffs, (size_t)(groups + 1) * 8) != RFB_OK ||
        (len > 0 && rfb_d2h(G.ctx, flat, dout, (size_t)len * w) != RFB_OK) || rfb_sync(G.ctx) != RFB_OK) { res = G.host->err_limit(); goto out; }
    res = G.host->vector(RFB_T_LIST, groups);
    if (!res || res->type == RFB_T_ERR) { res = G.host->err_limit(); goto out; }
    for (int64_t g = 0; g < groups; g++) {
        const int64_t c = offs[g + 1] - offs[g];
        obj_p v = G.host->vector((int8_t)ot, c);
        if (!v || v->type == RFB_T_ERR) {   /* hand back what exists as a shorter list, like the reference's error idiom (core/aggr.c:389-392) */
            res->len = g;
            G.host->drop_obj(res);
            res = G.host->err_limit();
            goto out;
        }
        if (c > 0) memcpy(RFB_OBJ_PAYLOAD(v), flat + (size_t)offs[g] * w, (size_t)c * w);
        RFB_OBJ_LIST(res)[g] = v;
    }
out:
    free(offs);
    free(flat);
    leave(sc);
    return res;
}
obj_p rfb_aggr_row(obj_p v, obj_p i) { return group_lists(0, v, i); }
obj_p rfb_aggr_collect(obj_p v, obj_p i) { return group_lists(1, v, i); }

/* ------------------------------------------------------------------ equi-join row matching */

static int is_key_vec(obj_p o) { return o && (o->type == RFB_T_I64 || o->type == RFB_T_SYMBOL || o->type == RFB_T_TIMESTAMP); }

/* cols: the key column itself (len == 1) or a LIST of len key columns; all I64-kind vectors of one length */
static int key_columns(obj_p cols, int64_t len, obj_p *out, int64_t *rows) {
    if (len < 1 || len > 8 || !cols) return 0;
    if (len == 1 && is_key_vec(cols)) { out[0] = cols; *rows = cols->len; return 1; }
    if (cols->type != RFB_T_LIST || cols->len != len) return 0;
    for (int64_t c = 0; c < len; c++) {
        obj_p v = RFB_OBJ_LIST(cols)[c];
        if (!is_key_vec(v) || (c > 0 && v->len != RFB_OBJ_LIST(cols)[0]->len)) return 0;
        out[c] = v;
    }
    *rows = out[0]->len;
    return 1;
}

/* index_left_join_obj / index_inner_join_obj (core/index.c:2886-3000): left = probe side, right = build side */
static obj_p join_index(int inner, obj_p lcols, obj_p rcols, int64_t len) {
    if (!G.ready) return NULL;
    obj_p l[8], r[8];
    int64_t ll = 0, rl = 0;
    if (!key_columns(lcols, len, l, &ll) || !key_columns(rcols, len, r, &rl)) return NULL;
    for (int64_t c = 0; c < len; c++)
        if (l[c]->type != r[c]->type) return NULL;
    if (too_small(ll) && too_small(rl)) return NULL;
    if (len == 1 && rl == 0) return NULL;   /* index_find_i64 answers an EMPTY vector (not nulls) when the searched column is empty
                                               (core/index.c:1512-1513, golden tests/lang.c:5118): the CPU body's quirk, its result */
    call_scope_t sc = enter();
    obj_p res = NULL;
    const void *dl[8], *dr[8];
    for (int64_t c = 0; c < len; c++) {
        dl[c] = dev_column(l[c]);
        dr[c] = dev_column(r[c]);
        if (!dl[c] || !dr[c]) { res = G.host->err_limit(); goto out; }
    }
    void *dids = dev_temp((size_t)(ll > 0 ? ll : 1) * 8), *dbid = inner ? dev_temp((size_t)(ll > 0 ? ll : 1) * 8) : NULL;
    if (!dids || (inner && !dbid)) { res = G.host->err_limit(); goto out; }
    if (!inner) {
        int rc = rfb_find_rows_dev(G.ctx, (int)len, (const int64_t *const *)dr, rl, (const int64_t *const *)dl, ll, (int64_t *)dids);
        if (rc) { res = status_to_obj(rc); goto out; }
        res = to_host_vector(RFB_T_I64, ll, dids);
    } else {
        int64_t count = 0;
        int rc = rfb_inner_join_dev(G.ctx, (int)len, (const int64_t *const *)dr, rl, (const int64_t *const *)dl, ll, (int64_t *)dids, (int64_t *)dbid, &count);
        if (rc) { res = status_to_obj(rc); goto out; }
        obj_p lids = to_host_vector(RFB_T_I64, count, dids), rids = to_host_vector(RFB_T_I64, count, dbid);
        res = G.host->vector(RFB_T_LIST, 2);
        if (!res || res->type == RFB_T_ERR || !lids || !rids || lids->type == RFB_T_ERR || rids->type == RFB_T_ERR) {
            if (res && res->type != RFB_T_ERR) { res->len = 0; G.host->drop_obj(res); }
            if (lids && lids->type != RFB_T_ERR) G.host->drop_obj(lids);
            if (rids && rids->type != RFB_T_ERR) G.host->drop_obj(rids);
            res = G.host->err_limit();
            goto out;
        }
        RFB_OBJ_LIST(res)[0] = lids;
        RFB_OBJ_LIST(res)[1] = rids;
    }
out:
    leave(sc);
    return res;
}
obj_p rfb_index_left_join_obj(obj_p lcols, obj_p rcols, int64_t len) { return join_index(0, lcols, rcols, len); }
obj_p rfb_index_inner_join_obj(obj_p lcols, obj_p rcols, int64_t len) { return join_index(1, lcols, rcols, len); }
/* ray_find(x, y) on two I64-kind vectors of the same type (core/items.c:320-323 -> index_find_i64): the first index of every y in x */
obj_p rfb_ray_find(obj_p x, obj_p y) {
    if (!is_key_vec(x) || !is_key_vec(y) || x->type != y->type) return NULL;
    return join_index(0, y, x, 1);
}

/* index_asof_join_obj(lcols, lxcol, rcols, rxcol) (core/index.c:3194-3268): lists of key columns + the time column of each side */
obj_p rfb_index_asof_join_obj(obj_p lcols, obj_p lxcol, obj_p rcols, obj_p rxcol) {
    if (!G.ready || !lcols || !rcols || !lxcol || !rxcol) return NULL;
    if (lcols->type != RFB_T_LIST || rcols->type != RFB_T_LIST || lcols->len < 1 || lcols->len > 8 || rcols->len != lcols->len) return NULL;
    const int tt = lxcol->type;
    if (tt != rxcol->type || !(tt == RFB_T_I64 || tt == RFB_T_TIMESTAMP || tt == RFB_T_I32 || tt == RFB_T_DATE || tt == RFB_T_TIME)) return NULL;
    obj_p l[8], r[8];
    int64_t ll = 0, rl = 0;
    const int64_t len = lcols->len;
    for (int64_t c = 0; c < len; c++) {
        l[c] = RFB_OBJ_LIST(lcols)[c];
        r[c] = RFB_OBJ_LIST(rcols)[c];
        if (!is_key_vec(l[c]) || !is_key_vec(r[c]) || l[c]->type != r[c]->type || l[c]->len != l[0]->len || r[c]->len != r[0]->len) return NULL;
    }
    ll = l[0]->len;
    rl = r[0]->len;
    if (lxcol->len != ll || rxcol->len != rl || (too_small(ll) && too_small(rl))) return NULL;
    call_scope_t sc = enter();
    obj_p res = NULL;
    const void *dl[8], *dr[8];
    for (int64_t c = 0; c < len; c++) {
        dl[c] = dev_column(l[c]);
        dr[c] = dev_column(r[c]);
        if (!dl[c] || !dr[c]) { res = G.host->err_limit(); goto out; }
    }
    void *dlx = dev_column(lxcol), *drx = dev_column(rxcol), *dids = dev_temp((size_t)(ll > 0 ? ll : 1) * 8);
    if (!dlx || !drx || !dids) { res = G.host->err_limit(); goto out; }
    int rc = rfb_asof_join_dev(G.ctx, (int)len, (const int64_t *const *)dr, tt, drx, rl, (const int64_t *const *)dl, dlx, ll, (int64_t *)dids);
    if (rc) { res = status_to_obj(rc); goto out; }
    res = to_host_vector(RFB_T_I64, ll, dids);
out:
    leave(sc);
    return res;
}

/* ray_in(x, y) on two I64-kind vectors of one type (core/items.c:781-783 -> index_in_i64_i64, core/index.c:1291-1370): mask of the
 * x values that occur in y = "the first matching row of y exists" */
obj_p rfb_ray_in(obj_p x, obj_p y) {
    if (!G.ready || !is_key_vec(x) || !is_key_vec(y) || x->type != y->type) return NULL;
    if (x->len == 0 || (too_small(x->len) && too_small(y->len))) return NULL;
    call_scope_t sc = enter();
    obj_p res;
    const void *dx = dev_column(x), *dy = dev_column(y);
    void *dids = dev_temp((size_t)x->len * 8), *dmask = dev_temp((size_t)x->len);
    if (!dx || !dy || !dids || !dmask) { res = G.host->err_limit(); goto out; }
    int rc = rfb_find_rows_dev(G.ctx, 1, (const int64_t *const *)&dy, y->len, (const int64_t *const *)&dx, x->len, (int64_t *)dids);
    if (!rc) {
        rfb_scalar_t none;
        memset(&none, 0, sizeof(none));
        none.type = RFB_T_I64;
        none.v.i64 = RFB_NULL_I64;
        rc = rfb_cmp_dev(G.ctx, RFB_NE, RFB_T_I64, dids, x->len, NULL, RFB_T_I64, NULL, -1, &none, (uint8_t *)dmask);
    }
    if (rc) { res = status_to_obj(rc); goto out; }
    res = to_host_vector(RFB_T_B8, x->len, dmask);
out:
    leave(sc);
    return res;
}

/* ray_distinct, hash branch (core/index.c:579-603).  The reference inserts the rows in row order into an open-addressing table of
 * next_prime(ceil(len / 0.75)) slots (slot = key % size, linear probing, core/hash.c:35-55, :129-148) and returns the table in SLOT
 * order.  The device finds the distinct keys in first-occurrence order (the sparse grouping path); which slot each of them ends up
 * in only depends on that order, so the host replays the insertions of the DISTINCT keys (not of the rows) over a sparse image of
 * the table and orders the keys by slot.  Negative keys (the reference indexes before its table for them) and nulls (skipped
 * there, after a scope pass that wraps) stay on the CPU body. */
typedef struct { int64_t slot, key; } slot_key_t;
static int cmp_slot_key(const void *a, const void *b) {
    const int64_t x = ((const slot_key_t *)a)->slot, y = ((const slot_key_t *)b)->slot;
    return x < y ? -1 : (x > y ? 1 : 0);
}
static int is_prime_i64(int64_t x) {   /* ops_is_prime, core/ops.c:66-81 */
    if (x <= 1) return 0;
    if (x <= 3) return 1;
    if (x % 2 == 0 || x % 3 == 0) return 0;
    for (int64_t i = 5; i * i <= x; i += 6)
        if (x % i == 0 || x % (i + 2) == 0) return 0;
    return 1;
}
static obj_p distinct_sparse(obj_p x, void *dx) {
    const int64_t n = x->len;
    obj_p res = NULL;
    rfb_group_info_t info;
    int64_t *hk = NULL, *occ = NULL;
    slot_key_t *sk = NULL;
    void *dg = dev_temp((size_t)n * 8), *dfi = dev_temp((size_t)n * 8);
    if (!dg || !dfi) return G.host->err_limit();
    int rc = rfb_group_i64_dev(G.ctx, (const int64_t *)dx, NULL, n, (int64_t *)dg, (int64_t *)dfi, &info);
    if (rc) return status_to_obj(rc);
    const int64_t g = info.groups;
    if (g > n / 2 + 1024) return NULL;                 /* nearly all rows distinct: the replay is the reference's own work */
    void *dk = dev_temp((size_t)(g > 0 ? g : 1) * 8);
    if (!dk) return G.host->err_limit();
    rc = rfb_gather_dev(G.ctx, RFB_T_I64, dx, (const int64_t *)dfi, g, dk);
    if (rc) return status_to_obj(rc);
    hk = (int64_t *)malloc((size_t)(g > 0 ? g : 1) * 8);
    if (!hk || rfb_d2h(G.ctx, hk, dk, (size_t)g * 8) != RFB_OK || rfb_sync(G.ctx) != RFB_OK) { res = G.host->err_limit(); goto done; }
    for (int64_t i = 0; i < g; i++)
        if (hk[i] < 0) { res = NULL; goto done; }      /* negative or null keys: CPU body */
    int64_t size = (int64_t)ceil((double)n / 0.75);
    while (!is_prime_i64(size)) size++;
    /* sparse image of the table: an open-addressing set of the occupied slot numbers */
    int64_t cap = 16;
    while (cap < 4 * g) cap <<= 1;
    occ = (int64_t *)malloc((size_t)cap * 8);
    sk = (slot_key_t *)malloc((size_t)(g > 0 ? g : 1) * sizeof(slot_key_t));
    if (!occ || !sk) { res = G.host->err_limit(); goto done; }
    for (int64_t i = 0; i < cap; i++) occ[i] = -1;
    for (int64_t i = 0; i < g; i++) {
        int64_t s = hk[i] % size;
        for (;;) {                                     /* first free slot at or after s (distinct keys never match an occupant) */
            uint64_t h = ((uint64_t)s * 0x9E3779B97F4A7C15ULL) >> 20 & (uint64_t)(cap - 1);
            while (occ[h] != -1 && occ[h] != s) h = (h + 1) & (uint64_t)(cap - 1);
            if (occ[h] == -1) { occ[h] = s; break; }
            s = (s + 1) % size;
        }
        sk[i].slot = s;
        sk[i].key = hk[i];
    }
    qsort(sk, (size_t)g, sizeof(slot_key_t), cmp_slot_key);
    res = G.host->vector((int8_t)x->type, g);
    if (!res || res->type == RFB_T_ERR) { res = res ? res : G.host->err_limit(); goto done; }
    for (int64_t i = 0; i < g; i++) ((int64_t *)RFB_OBJ_PAYLOAD(res))[i] = sk[i].key;
    res->attrs |= 1;   /* ATTR_DISTINCT */
done:
    free(hk); free(occ); free(sk);
    return res;
}

/* ray_distinct on an I64-kind vector with a dense key range (core/compose.c:867-872 -> index_distinct_i64): ascending distinct
 * keys, ATTR_DISTINCT set on the result.  A sparse range takes distinct_sparse above (the reference's hash branch, slot order). */
obj_p rfb_ray_distinct(obj_p x) {
    if (!G.ready || !is_key_vec(x) || too_small(x->len) || x->len == 0) return NULL;
    call_scope_t sc = enter();
    obj_p res = NULL;
    int64_t count = 0;
    void *dx = dev_column(x), *dout = dev_temp((size_t)x->len * 8);
    if (!dx || !dout) { res = G.host->err_limit(); goto out; }
    int rc = rfb_distinct_i64_dev(G.ctx, (const int64_t *)dx, x->len, (int64_t *)dout, &count);
    if (rc == RFB_ERR_ARG) { res = distinct_sparse(x, dx); goto out; }
    if (rc) { res = status_to_obj(rc); goto out; }
    res = to_host_vector(x->type, count, dout);
    if (res && res->type != RFB_T_ERR) res->attrs |= 1;   /* ATTR_DISTINCT (core/ops.h:52) */
out:
    leave(sc);
    return res;
}

/* ------------------------------------------------------------------ sort */

static obj_p sort_op(obj_p x, int desc) {
    if (!G.ready || !is_vec(x) || x->type == RFB_T_SYMBOL) return NULL; /* symbols sort by string: CPU body */
    if (x->attrs != 0) return NULL;                                     /* ATTR_ASC/DESC shortcut (core/sort.c:437-451) */
    if (too_small(x->len)) return NULL;
    call_scope_t sc = enter();
    obj_p res;
    void *dx = dev_column(x), *dp = dev_temp((size_t)(x->len > 0 ? x->len : 1) * 8);
    if (!dx || !dp) { res = G.host->err_limit(); goto out; }
    int rc = rfb_sort_dev(G.ctx, x->type, dx, x->len, desc, (int64_t *)dp);
    if (rc) { res = status_to_obj(rc); goto out; }
    res = to_host_vector(RFB_T_I64, x->len, dp);
out:
    leave(sc);
    return res;
}
obj_p rfb_ray_sort_asc(obj_p x) { return sort_op(x, 0); }
obj_p rfb_ray_sort_desc(obj_p x) { return sort_op(x, 1); }


/* ------------------------------------------------------------------ masks: and / or / not (core/logic.c:34-264, core/order.c:422-443) */

/* One step of the reference's `and` / `or` fold: res = res OP next, IN PLACE in res's payload like and_op_partial / or_op_partial
 * (core/logic.c:34-86) — res is a B8 vector, next a B8 vector of the same length or a b8 atom.  1 = done on the device (res's
 * payload now holds the result and its HBM image is registered, so the ray_where that follows ships nothing), 0 = declined,
 * < 0 = device failure (res untouched).  The binding evaluates the operands itself (ray_and / ray_or are special forms). */
int rfb_mask_logic_inplace(int is_or, obj_p res, obj_p next) {
    if (!G.ready || !res || !next || res->type != RFB_T_B8) return 0;
    const int nv = next->type == RFB_T_B8;
    if (!nv && next->type != -RFB_T_B8) return 0;
    if (nv && next->len != res->len) return 0;       /* the reference answers a type error: its body, its error object */
    if (too_small(res->len) || res->len == 0) return 0;
    if (gated_out(res, nv ? next : NULL)) return 0;   /* cost gate: a bare CPU-side mask pair */
    call_scope_t sc = enter();
    int done = -1;
    void *da = dev_column(res), *db = nv ? dev_column(next) : NULL, *dout = dev_temp((size_t)res->len);
    if (da && (!nv || db) && dout &&
        rfb_mask_logic_dev(G.ctx, is_or ? RFB_M_OR : RFB_M_AND, (const uint8_t *)da, res->len, (const uint8_t *)db, nv ? res->len : -1,
                           nv ? 0 : next->u8, (uint8_t *)dout) == RFB_OK) {
        if (lazy_on == 1) lazy_on_free(res);          /* a still-pending lazy payload is being replaced: open its pages first */
        if (fill_host_vector(res, RFB_T_B8, res->len, dout) == RFB_OK) done = 1;
    }
    leave(sc);
    return done;
}

obj_p rfb_ray_not(obj_p x) {
    if (!G.ready || !x || x->type != RFB_T_B8) return NULL;      /* atoms and type errors: CPU body */
    if (too_small(x->len) || x->len >= (1ll << 31) || gated_out(x, NULL)) return NULL;   /* (the reference's loop counter is 32 bits wide) */
    call_scope_t sc = enter();
    obj_p res;
    void *dx = dev_column(x), *dout = dev_temp((size_t)(x->len > 0 ? x->len : 1));
    if (!dx || !dout) { res = G.host->err_limit(); goto out; }
    int rc = rfb_mask_logic_dev(G.ctx, RFB_M_NOT, (const uint8_t *)dx, x->len, NULL, -1, 0, (uint8_t *)dout);
    if (rc) { res = status_to_obj(rc); goto out; }
    res = to_host_vector(RFB_T_B8, x->len, dout);
out:
    leave(sc);
    return res;
}

/* ------------------------------------------------------------------ at_ids, ray_asc / ray_desc, ray_xasc / ray_xdesc */

#define RFB_T_TABLE 98   /* core/rayforce.h:85: a 2-list [column names (SYMBOL vector), LIST of columns] */

/* a table whose columns are all fixed-width vectors of one length */
static int is_flat_table(obj_p t, int64_t *rows) {
    if (!t || t->type != RFB_T_TABLE || t->len != 2) return 0;
    obj_p names = RFB_OBJ_LIST(t)[0], cols = RFB_OBJ_LIST(t)[1];
    if (!names || names->type != RFB_T_SYMBOL || !cols || cols->type != RFB_T_LIST || cols->len != names->len || cols->len == 0) return 0;
    for (int64_t c = 0; c < cols->len; c++) {
        obj_p v = RFB_OBJ_LIST(cols)[c];
        if (!is_vec(v) || v->len != RFB_OBJ_LIST(cols)[0]->len) return 0;
    }
    *rows = RFB_OBJ_LIST(cols)[0]->len;
    return 1;
}

/* out column = col[ids] on the device -> new host vector of col's type */
static obj_p gather_column(obj_p col, const void *dids, int64_t m) {
    void *dc = dev_column(col), *dout = dev_temp((size_t)(m > 0 ? m : 1) * type_size(col->type));
    if (!dc || !dout) return G.host->err_limit();
    int rc = rfb_gather_dev(G.ctx, col->type, dc, (const int64_t *)dids, m, dout);
    if (rc) return status_to_obj(rc);
    return to_host_vector(col->type, m, dout);
}

/* table(names, [col[ids] for every column]) (at_ids TYPE_TABLE branch, core/rayforce.c:1184-1201) */
static obj_p gather_table(obj_p t, const void *dids, int64_t m) {
    obj_p names = RFB_OBJ_LIST(t)[0], cols = RFB_OBJ_LIST(t)[1];
    obj_p out = G.host->vector(RFB_T_LIST, cols->len);
    if (!out || out->type == RFB_T_ERR) return G.host->err_limit();
    for (int64_t c = 0; c < cols->len; c++) RFB_OBJ_LIST(out)[c] = G.host->null_obj;
    for (int64_t c = 0; c < cols->len; c++) {
        obj_p v = gather_column(RFB_OBJ_LIST(cols)[c], dids, m);
        if (!v || v->type == RFB_T_ERR) { G.host->drop_obj(out); return v ? v : G.host->err_limit(); }
        RFB_OBJ_LIST(out)[c] = v;
    }
    obj_p res = G.host->vector(RFB_T_LIST, 2);
    if (!res || res->type == RFB_T_ERR) { G.host->drop_obj(out); return G.host->err_limit(); }
    RFB_OBJ_LIST(res)[0] = G.host->clone_obj(names);
    RFB_OBJ_LIST(res)[1] = out;
    res->type = RFB_T_TABLE;
    return res;
}

/* at_ids(obj, ids, len) (core/rayforce.c:1100-1201): obj[ids] for a fixed-width vector or a table of such columns.  `ids` is a bare
 * host array — normally the payload of a row-id vector this layer produced (ray_where, a sort permutation), whose HBM image is
 * then found by its address.  No bounds checks, like the reference (its callers check). */
obj_p rfb_at_ids(obj_p obj, const int64_t *ids, int64_t len) {
    if (!G.ready || !obj || !ids || len < 0) return NULL;
    int64_t rows = 0;
    const int tab = is_flat_table(obj, &rows);
    if (!tab && !is_vec(obj)) return NULL;             /* GUID / LIST / ENUM / parted: CPU body */
    if (too_small(len)) return NULL;
    call_scope_t sc = enter();
    obj_p res;
    void *di = dev_payload(ids, len, RFB_T_I64, 0);
    if (!di) { res = G.host->err_limit(); goto out; }
    res = tab ? gather_table(obj, di, len) : gather_column(obj, di, len);
out:
    leave(sc);
    return res;
}

#define RFB_ATTR_DISTINCT 1
#define RFB_ATTR_ASC 2
#define RFB_ATTR_DESC 4

/* ray_asc / ray_desc (core/order.c:74-244): the sorted VALUES of a fixed-width vector = stable key sort + gather, both on the
 * device; the result carries ATTR_ASC / ATTR_DESC and keeps ATTR_DISTINCT */
static obj_p sorted_values(obj_p x, int desc) {
    if (!G.ready || !is_vec(x) || x->type == RFB_T_SYMBOL) return NULL;            /* symbols order by their strings: CPU body */
    if (x->attrs & (RFB_ATTR_ASC | RFB_ATTR_DESC)) return NULL;                    /* clone / reverse shortcuts (core/order.c:79-83) */
    if (too_small(x->len)) return NULL;
    call_scope_t sc = enter();
    obj_p res;
    void *dx = dev_column(x), *dp = dev_temp((size_t)(x->len > 0 ? x->len : 1) * 8);
    if (!dx || !dp) { res = G.host->err_limit(); goto out; }
    int rc = rfb_sort_dev(G.ctx, x->type, dx, x->len, desc, (int64_t *)dp);
    if (rc) { res = status_to_obj(rc); goto out; }
    res = gather_column(x, dp, x->len);
    if (res && res->type != RFB_T_ERR) res->attrs |= (uint8_t)((desc ? RFB_ATTR_DESC : RFB_ATTR_ASC) | (x->attrs & RFB_ATTR_DISTINCT));
out:
    leave(sc);
    return res;
}
obj_p rfb_ray_asc(obj_p x) { return sorted_values(x, 0); }
obj_p rfb_ray_desc(obj_p x) { return sorted_values(x, 1); }

/* ray_xasc / ray_xdesc (core/order.c:246-420): a table ordered by one column (y = symbol atom) or by several (y = symbol vector:
 * one stable sort per key column from the last to the first, each on the column as reordered so far — the same composition of
 * permutations, all on the device), then every column gathered by the final permutation. */
static obj_p sorted_table(obj_p x, obj_p y, int desc) {
    if (!G.ready || !y) return NULL;
    int64_t rows = 0;
    if (!is_flat_table(x, &rows)) return NULL;
    const int atom = y->type == -RFB_T_SYMBOL;
    if (!atom && !(y->type == RFB_T_SYMBOL && y->len >= 1 && y->len <= 16)) return NULL;
    if (too_small(rows) || rows == 0) return NULL;
    obj_p names = RFB_OBJ_LIST(x)[0], cols = RFB_OBJ_LIST(x)[1], key[16];
    const int64_t nk = atom ? 1 : y->len;
    for (int64_t k = 0; k < nk; k++) {
        const int64_t sym = atom ? y->i64 : ((const int64_t *)RFB_OBJ_PAYLOAD(y))[k];
        key[k] = NULL;
        for (int64_t c = 0; c < names->len; c++)
            if (((const int64_t *)RFB_OBJ_PAYLOAD(names))[c] == sym) { key[k] = RFB_OBJ_LIST(cols)[c]; break; }
        if (!key[k] || key[k]->type == RFB_T_SYMBOL) return NULL;                  /* unknown column / symbol keys: CPU body */
        if (atom && key[k]->attrs != 0) return NULL;                               /* ray_iasc's attribute shortcuts */
    }
    call_scope_t sc = enter();
    obj_p res;
    void *perm = NULL;
    int rc = RFB_OK;
    for (int64_t k = nk - 1; k >= 0 && !rc; k--) {
        void *dk = dev_column(key[k]), *local = dev_temp((size_t)rows * 8);
        if (!dk || !local) { res = G.host->err_limit(); goto out; }
        if (!perm) {                                   /* first pass: the column as it stands */
            rc = rfb_sort_dev(G.ctx, key[k]->type, dk, rows, desc, (int64_t *)local);
            perm = local;
        } else {                                       /* the column reordered so far, its stable order, composed: perm = perm[local] */
            void *re = dev_temp((size_t)rows * type_size(key[k]->type)), *np = dev_temp((size_t)rows * 8);
            if (!re || !np) { res = G.host->err_limit(); goto out; }
            rc = rfb_gather_dev(G.ctx, key[k]->type, dk, (const int64_t *)perm, rows, re);
            if (!rc) rc = rfb_sort_dev(G.ctx, key[k]->type, re, rows, desc, (int64_t *)local);
            if (!rc) rc = rfb_gather_dev(G.ctx, RFB_T_I64, perm, (const int64_t *)local, rows, np);
            perm = np;
        }
    }
    if (rc) { res = status_to_obj(rc); goto out; }
    res = gather_table(x, perm, rows);
out:
    leave(sc);
    return res;
}
obj_p rfb_ray_xasc(obj_p x, obj_p y) { return sorted_table(x, y, 0); }
obj_p rfb_ray_xdesc(obj_p x, obj_p y) { return sorted_table(x, y, 1); }

/* ------------------------------------------------------------------ aggr_first / aggr_last, index_group_list */

/* the reference's pool_split_by_mem (core/pool.c:450-478): the number of worker chunks aggr_map cuts `rows` rows into */
static int64_t aggr_chunks(int64_t rows, int64_t groups, int width) {
    const int64_t threads = G.host->executors ? G.host->executors() : 1;
    if (rows < 16384 || rows <= threads) return 1;
    const int64_t mem = groups * width;
    if (mem > (64ll << 20)) return 1;
    if (mem > 0 && (64ll << 20) / mem < threads) return (64ll << 20) / mem < 1 ? 1 : (64ll << 20) / mem;
    return threads;
}

static obj_p first_last(int last, obj_p val, obj_p index) {
    if (!G.ready || !is_vec(val) || !index || index->type != RFB_T_LIST || index->len != 7) return NULL;
    obj_p *ix = RFB_OBJ_LIST(index);
    if (!ix[0] || ix[0]->type != -RFB_T_I64 || ix[0]->i64 != RFB_INDEX_IDS || !ix[1] || ix[1]->type != -RFB_T_I64) return NULL;
    obj_p gids = ix[2], filter = ix[5], firsts = ix[6];
    if (!gids || gids->type != RFB_T_I64) return NULL;
    const int filtered = !is_null_obj(filter);
    if (filtered && filter->type != RFB_T_I64) return NULL;
    /* aggr_first without first_ids takes the partial path (first NON-null value per chunk, core/aggr.c:394-439): CPU body */
    if (!last && (is_null_obj(firsts) || firsts->type != RFB_T_I64)) return NULL;
    if (last && (val->type == RFB_T_B8 || val->type == RFB_T_U8)) return G.host->err_type();      /* core/aggr.c:904-1075: no case */
    const int64_t groups = ix[1]->i64, len = gids->len;
    if (too_small(len)) return NULL;
    call_scope_t sc = enter();
    obj_p res;
    void *dv = dev_column(val), *dg = dev_column(gids), *df = filtered ? dev_column(filter) : NULL;
    void *dout = dev_temp((size_t)(groups > 0 ? groups : 1) * 8);
    if (!dv || !dg || (filtered && !df) || !dout) { res = G.host->err_limit(); goto out; }
    int rc = last ? rfb_aggr_last_dev(G.ctx, val->type, dv, (const int64_t *)df, (const int64_t *)dg, len, groups,
                                      aggr_chunks(len, groups, type_size(val->type)), dout)
                  : rfb_aggr_dev(G.ctx, RFB_A_FIRST, val->type, dv, (const int64_t *)df, (const int64_t *)dg, len, groups, dout);
    if (rc) { res = status_to_obj(rc); goto out; }
    res = to_host_vector(val->type, groups, dout);
out:
    leave(sc);
    return res;
}
obj_p rfb_aggr_first(obj_p v, obj_p i) { return first_last(0, v, i); }
obj_p rfb_aggr_last(obj_p v, obj_p i) { return first_last(1, v, i); }

/* index_group_list (core/index.c:2731-2793): group rows by the tuple of 2..8 I64-kind key columns -> the reference's 7-element
 * index [IDS, groups, group_ids, null, null, filter, first_ids], groups numbered by first occurrence (the reference's own order
 * on one worker; with several its radix path numbers them by partition, SURVEY Q9 — results are compared as sets there). */
obj_p rfb_index_group_list(obj_p keys, obj_p filter) {
    if (!G.ready || !keys || keys->type != RFB_T_LIST || keys->len < 2 || keys->len > 8) return NULL;   /* 1 column: index_group (wrapped itself) */
    const int filtered = !is_null_obj(filter);
    if (filtered && filter->type != RFB_T_I64) return NULL;
    obj_p *kc = RFB_OBJ_LIST(keys);
    for (int64_t c = 0; c < keys->len; c++)
        if (!is_key_vec(kc[c]) || kc[c]->len != kc[0]->len) return NULL;
    const int64_t len = filtered ? filter->len : kc[0]->len;
    if (too_small(len) || len == 0) return NULL;
    call_scope_t sc = enter();
    obj_p res = NULL, gids = NULL, firsts = NULL;
    rfb_group_info_t info;
    const void *dk[8];
    for (int64_t c = 0; c < keys->len; c++) { dk[c] = dev_column(kc[c]); if (!dk[c]) { res = G.host->err_limit(); goto out; } }
    void *df = filtered ? dev_column(filter) : NULL, *dg = dev_temp((size_t)len * 8), *dfi = dev_temp((size_t)len * 8);
    if ((filtered && !df) || !dg || !dfi) { res = G.host->err_limit(); goto out; }
    int rc = rfb_group_keys_i64_dev(G.ctx, (int)keys->len, (const int64_t *const *)dk, (const int64_t *)df, len, (int64_t *)dg, (int64_t *)dfi, &info);
    if (rc) { res = status_to_obj(rc); goto out; }
    gids = to_host_vector(RFB_T_I64, len, dg);
    firsts = to_host_vector(RFB_T_I64, info.groups, dfi);
    res = G.host->vector(RFB_T_LIST, 7);
    if (!res || res->type == RFB_T_ERR || !gids || !firsts || gids->type == RFB_T_ERR || firsts->type == RFB_T_ERR) {
        if (res && res->type != RFB_T_ERR) { res->len = 0; G.host->drop_obj(res); }
        if (gids && gids->type != RFB_T_ERR) G.host->drop_obj(gids);
        if (firsts && firsts->type != RFB_T_ERR) G.host->drop_obj(firsts);
        res = G.host->err_limit();
        goto out;
    }
    RFB_OBJ_LIST(res)[0] = i64_atom(RFB_INDEX_IDS);
    RFB_OBJ_LIST(res)[1] = i64_atom(info.groups);
    RFB_OBJ_LIST(res)[2] = gids;
    RFB_OBJ_LIST(res)[3] = i64_atom(RFB_NULL_I64);
    RFB_OBJ_LIST(res)[4] = G.host->null_obj;
    RFB_OBJ_LIST(res)[5] = filtered ? G.host->clone_obj(filter) : G.host->null_obj;
    RFB_OBJ_LIST(res)[6] = firsts;
out:
    leave(sc);
    return res;
}

/* ------------------------------------------------------------------ fused query entry points */

/* Plugin mode (reference core/dynlib.c:153-218: `(loadfn "librfb200_ops.so" "rfb_where_lt_sum" 2)` from a STOCK reference
 * binary, no shim): nobody has called rfb_ops_init.  The first plugin call binds the host API itself from the symbols the
 * reference exports to plugins (rayforce.syms: vector, i64, clone_obj, drop_obj, __NULL_OBJ; err_* only when the binary was
 * linked with -rdynamic — otherwise an operand error answers the null object instead of the reference's error object). */
#include <dlfcn.h>
static rfb_host_api_t plugin_host;
static obj_p (*plugin_i64)(int64_t);
static obj_p (*plugin_err_type)(int8_t, int8_t, uint8_t, uint8_t);
static obj_p (*plugin_err_length)(uint8_t, uint8_t, uint8_t, uint8_t, int64_t, int64_t);
static obj_p (*plugin_err_limit)(int64_t);
static obj_p plugin_atom(int8_t type) {
    obj_p a = plugin_i64(0);                       /* a 16-byte atom; retype it (atoms of every numeric type share the layout) */
    if (a) a->type = (int8_t)-type;
    return a;
}
static obj_p plugin_e_type(void) { return plugin_err_type ? plugin_err_type(0, 0, 0, 0) : plugin_host.null_obj; }
static obj_p plugin_e_length(void) { return plugin_err_length ? plugin_err_length(0, 0, 0, 0, 0, 0) : plugin_host.null_obj; }
static obj_p plugin_e_limit(void) { return plugin_err_limit ? plugin_err_limit(0) : plugin_host.null_obj; }
static int plugin_bind(void) {
    if (G.ready) return 1;
    void *self = RTLD_DEFAULT;
    plugin_host.vector = (obj_p(*)(int8_t, int64_t))dlsym(self, "vector");
    plugin_i64 = (obj_p(*)(int64_t))dlsym(self, "i64");
    plugin_host.clone_obj = (obj_p(*)(obj_p))dlsym(self, "clone_obj");
    plugin_host.drop_obj = (void (*)(obj_p))dlsym(self, "drop_obj");
    plugin_host.null_obj = (obj_p)dlsym(self, "__NULL_OBJ");
    if (!plugin_host.vector || !plugin_i64 || !plugin_host.clone_obj || !plugin_host.drop_obj || !plugin_host.null_obj) {
        set_err("plugin mode: the host process does not export vector / i64 / clone_obj / drop_obj / __NULL_OBJ");
        return 0;
    }
    *(void **)&plugin_err_type = dlsym(self, "err_type");
    *(void **)&plugin_err_length = dlsym(self, "err_length");
    *(void **)&plugin_err_limit = dlsym(self, "err_limit");
    plugin_host.atom = plugin_atom;
    plugin_host.err_type = plugin_e_type;
    plugin_host.err_length = plugin_e_length;
    plugin_host.err_limit = plugin_e_limit;
    return rfb_ops_init(&plugin_host, 0) == 0;
}

static obj_p where_fold(int op, int what, obj_p pred, obj_p k, obj_p val) {
    if (!G.ready && !plugin_bind()) return plugin_host.null_obj;   /* no usable GPU: the null object (NULL when not even a host) */
    if (!G.ready || !is_vec(pred) || !is_vec(val) || !is_atom(k)) return G.ready ? G.host->err_type() : NULL;
    if (pred->len != val->len) return G.host->err_length();
    if (op < RFB_EQ || op > RFB_GE || what < F_SUM || what > F_AVG || !fold_type_ok(what, val->type)) return G.host->err_type();
    rfb_fold_t f;
    rfb_scalar_t ks = scalar_of(k);
    const int folds = (what == F_MIN || what == F_MAX) ? (RFB_F_MIN | RFB_F_MAX) : (RFB_F_SUM | RFB_F_CNT);
    int rc;
    if (G.scope_depth == 0) {
        /* one-shot: stream the host column(s) through the chunked copy/compute pipeline (no device residency needed) */
        if (G.mgpu && rfb_mgpu_devices(G.mgpu) > 1)
            rc = rfb_mgpu_filter_fold_host(G.mgpu, op, pred->type, RFB_OBJ_PAYLOAD(pred), &ks, folds, val->type, RFB_OBJ_PAYLOAD(val), pred->len, 0, &f, NULL);
        else
            rc = rfb_filter_fold_host(G.ctx, op, pred->type, RFB_OBJ_PAYLOAD(pred), &ks, folds, val->type, RFB_OBJ_PAYLOAD(val), pred->len, 0, &f, NULL);
    } else {
        void *dp = dev_column(pred), *dv = (val == pred) ? dp : dev_column(val);
        if (!dp || !dv) return G.host->err_limit();
        rc = rfb_filter_fold_dev(G.ctx, op, pred->type, dp, &ks, folds, val->type, dv, pred->len, &f);
    }
    if (rc) return status_to_obj(rc);
    return fold_result(what, val->type, &f);
}

obj_p rfb_where_lt_sum(obj_p col, obj_p k) { return where_fold(RFB_LT, F_SUM, col, k, col); }

obj_p rfb_where_fold(obj_p *args, int64_t n) {
    if (!G.ready && !plugin_bind()) return plugin_host.null_obj;
    if (n != 5 || !args[0] || !args[1] || args[0]->type != -RFB_T_I64 || args[1]->type != -RFB_T_I64) return G.host->err_type();
    return where_fold((int)args[0]->i64, (int)args[1]->i64, args[2], args[3], args[4]);
}
