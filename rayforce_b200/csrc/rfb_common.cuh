// rfb_common.cuh — shared device-side vocabulary for the sm_100a kernels: element kinds, the reference's null
// semantics as device functions (core/ops.h:63-197 of the reference), 128-bit streaming loads, block reductions,
// and the context object behind the C ABI (include/rfb200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rfb200.h"

typedef int16_t i16;
typedef int32_t i32;
typedef int64_t i64;
typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef double f64;

// ------------------------------------------------------------------ context (host side)

#define RFB_RESULT_SLOTS 1024
#define RFB_STAGE_BUFS 3
#define RFB_HOST_RING 4
#define RFB_HOST_RING_BYTES (16u << 20)

struct rfb_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;       // where every kernel of this context is enqueued
    bool own_stream;
    cudaStream_t copy_stream;  // host layer: H2D pipeline
    i64 launches;
    // small device scratch for reduction partials + ticket + result
    void *d_scratch;
    size_t scratch_bytes;
    void *h_result;            // mapped pinned host memory, RFB_RESULT_SLOTS rfb_fold_t slots: kernels write results here
    i64 *h_count;              // mapped pinned: small integer results (selection counts, group counts)
    int result_slot;           // slot the next fold launch reports into (host layer: one per chunk)
    void *result_override;     // when set (rfb_ctx_set_result_ptr) fold kernels report there instead (device-visible memory)
    // growable device workspace (sort / group temporaries)
    void *d_work;
    size_t work_bytes;
    // second growable buffer for callers that run workspace-using entry points themselves (fused multi-key keys, the
    // temporaries of the median / row-list pipelines): keeps cudaMalloc / cudaFree out of the per-call path
    void *d_aux;
    size_t aux_bytes;
    void *d_aux2;              // third level: composite entry points (asof join) that call aux-using entry points themselves
    size_t aux2_bytes;
    // host layer staging
    void *d_stage[2][RFB_STAGE_BUFS];  // [column][ring slot] device staging for chunked column shipping
    size_t stage_bytes;
    // pageable host memory: copier threads fill a ring of pinned buffers that the DMA engine drains (rfb_host.cu)
    void *h_ring[RFB_HOST_RING];
    cudaEvent_t ev_ring[RFB_HOST_RING];
    int ring_next;
    void *copy_pool;
    cudaEvent_t ev_copy[RFB_STAGE_BUFS], ev_kernel[RFB_STAGE_BUFS];
    // peer mailboxes (k_peer.cu): the one-shot all-reduce of fold results over NVLink P2P stores
    void *mbox;                // own mailbox (device memory, exported to the peers through CUDA IPC)
    void *mbox_peer[16];       // every rank's mailbox as mapped into this process ([rank] = own)
    int mbox_rank, mbox_world;
    unsigned long long mbox_seq;
    // group exchange buffers (k_peer.cu): every rank's partial (key, sum, count) lists, read by the peers over NVLink
    void *gx;                  // own buffer (device memory, exported through CUDA IPC): 2 halves x (header + 3 x capacity words)
    void *gx_peer[16];
    i64 gx_cap;
    int gx_rank, gx_world;
    unsigned long long gx_seq;
};

void rfb_set_error(const char *fmt, ...);
int rfb_cuda_fail(cudaError_t e, const char *what, const char *file, int line);
int rfb_ensure_work(rfb_ctx_t *ctx, size_t bytes, void **out);
int rfb_ensure_aux(rfb_ctx_t *ctx, size_t bytes, void **out);
int rfb_ensure_aux2(rfb_ctx_t *ctx, size_t bytes, void **out);
// host<->device copies that pick the fast route for the memory they are given (rfb_host.cu): pinned -> one async DMA;
// pageable -> copier threads through the pinned ring.  `stream`: where the DMA is enqueued.
int rfb_copy_h2d(rfb_ctx_t *ctx, void *dst_dev, const void *src_host, size_t bytes, cudaStream_t stream);
int rfb_copy_d2h(rfb_ctx_t *ctx, void *dst_host, const void *src_dev, size_t bytes, cudaStream_t stream);
void rfb_copy_shutdown(rfb_ctx_t *ctx);
void rfb_peer_mailbox_release(rfb_ctx_t *ctx);   // k_peer.cu
// launch-only entry points of k_fold.cu: the result lands in h_result[result_slot] once the stream drains
int rfb_fold_launch(rfb_ctx_t *ctx, int folds, int type, const void *x, i64 n);
int rfb_filter_fold_launch(rfb_ctx_t *ctx, int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k, int folds,
                           int val_type, const void *val, i64 n);

// k_stats.cu: grouped median / deviation behind rfb_aggr_dev(RFB_A_MED / RFB_A_DEV)
int rfb_aggr_med_launch(rfb_ctx_t *ctx, int val_type, const void *val, const int64_t *filter, const int64_t *group_ids, int64_t len, int64_t groups, double *out);
int rfb_aggr_stddev_launch(rfb_ctx_t *ctx, int val_type, const void *val, const int64_t *filter, const int64_t *group_ids, int64_t len, int64_t groups, double *out);

// run-time tuning knobs: read from the environment ONCE (first use), never on the per-call path; rfb_options_reload() re-reads them
struct rfb_options_t { int loaded; int group_strategy; i64 part_min_rows; int accum_tma; int sort_algo; };
const rfb_options_t *rfb_options();

// k_fused_group.cu: per-group integer sums + counts over dense group ids through the narrow partitioned passes
size_t rfb_narrow_sums_bytes(i64 n);
int rfb_narrow_sums(rfb_ctx_t *ctx, const i64 *gid, const i64 *val, i64 n, i64 groups, void *work, u64 *sum, u64 *cnt, u32 *has_null,
                    bool nulls_count, bool *done);

#define RFB_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) return rfb_cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define RFB_CHECK_LAUNCH(ctx)                                                   \
    do {                                                                        \
        (ctx)->launches++;                                                      \
        cudaError_t _e = cudaGetLastError();                                    \
        if (_e != cudaSuccess) return rfb_cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define RFB_ARG(cond, msg)                          \
    do {                                            \
        if (!(cond)) {                              \
            rfb_set_error("bad argument: %s", msg); \
            return RFB_ERR_ARG;                     \
        }                                           \
    } while (0)

// ------------------------------------------------------------------ element kinds

enum Kind { K_NONE = 0, K_U8 = 1, K_I16 = 2, K_I32 = 3, K_I64 = 4, K_F64 = 5 };

static inline __host__ __device__ int rfb_kind_of(int type) {
    switch (type) {
        case RFB_B8: case RFB_U8: return K_U8;
        case RFB_I16: return K_I16;
        case RFB_I32: case RFB_DATE: case RFB_TIME: return K_I32;
        case RFB_I64: case RFB_SYMBOL: case RFB_TIMESTAMP: return K_I64;
        case RFB_F64: return K_F64;
        default: return K_NONE;
    }
}
static inline __host__ __device__ int rfb_type_size(int type) {
    switch (rfb_kind_of(type)) {
        case K_U8: return 1; case K_I16: return 2; case K_I32: return 4; case K_I64: case K_F64: return 8;
        default: return 0;
    }
}

// ------------------------------------------------------------------ null semantics (reference core/ops.h)

#define NULL_I16 RFB_NULL_I16
#define NULL_I32 RFB_NULL_I32
#define NULL_I64 RFB_NULL_I64

__host__ __device__ __forceinline__ u64 f64_bits(f64 x) {
#ifdef __CUDA_ARCH__
    return (u64)__double_as_longlong(x);
#else
    u64 u; memcpy(&u, &x, 8); return u;
#endif
}
__host__ __device__ __forceinline__ f64 bits_f64(u64 u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    f64 d; memcpy(&d, &u, 8); return d;
#endif
}
// ISNANF64 (core/ops.h:63-70): by bit pattern
__host__ __device__ __forceinline__ bool isnan64(f64 x) { return (f64_bits(x) & 0x7FFFFFFFFFFFFFFFULL) > 0x7FF0000000000000ULL; }
__host__ __device__ __forceinline__ f64 null_f64() { return bits_f64(0x7FF8000000000000ULL); }

template <typename T> struct Elem;
template <> struct Elem<u8>  { static constexpr int kind = K_U8;  typedef i64 acc_t;
    __device__ __forceinline__ static bool is_null(u8) { return false; }
    __device__ __forceinline__ static u8 null() { return 0; } };
template <> struct Elem<i16> { static constexpr int kind = K_I16; typedef i64 acc_t;
    __device__ __forceinline__ static bool is_null(i16 v) { return v == NULL_I16; }
    __device__ __forceinline__ static i16 null() { return NULL_I16; } };
template <> struct Elem<i32> { static constexpr int kind = K_I32; typedef i64 acc_t;
    __device__ __forceinline__ static bool is_null(i32 v) { return v == NULL_I32; }
    __device__ __forceinline__ static i32 null() { return NULL_I32; } };
template <> struct Elem<i64> { static constexpr int kind = K_I64; typedef i64 acc_t;
    __device__ __forceinline__ static bool is_null(i64 v) { return v == NULL_I64; }
    __device__ __forceinline__ static i64 null() { return NULL_I64; } };
template <> struct Elem<f64> { static constexpr int kind = K_F64; typedef f64 acc_t;
    __device__ __forceinline__ static bool is_null(f64 v) { return isnan64(v); }
    __device__ __forceinline__ static f64 null() { return null_f64(); } };

// comparisons: integers compare as plain values (a null is the smallest value); doubles put NaN below everything
// and NaN == NaN (core/ops.h:74-127)
template <int OP> __device__ __forceinline__ bool cmp_int(i64 a, i64 b) {
    if (OP == RFB_EQ) return a == b;
    if (OP == RFB_NE) return a != b;
    if (OP == RFB_LT) return a < b;
    if (OP == RFB_GT) return a > b;
    if (OP == RFB_LE) return a <= b;
    return a >= b;
}
__device__ __forceinline__ bool eq_f64(f64 a, f64 b) { return isnan64(a) ? isnan64(b) : (isnan64(b) ? false : a == b); }
__device__ __forceinline__ bool lt_f64(f64 a, f64 b) { return isnan64(a) ? !isnan64(b) : (isnan64(b) ? false : a < b); }
__device__ __forceinline__ bool gt_f64(f64 a, f64 b) { return isnan64(b) ? !isnan64(a) : (isnan64(a) ? false : a > b); }
template <int OP> __device__ __forceinline__ bool cmp_flt(f64 a, f64 b) {
    if (OP == RFB_EQ) return eq_f64(a, b);
    if (OP == RFB_NE) return !eq_f64(a, b);
    if (OP == RFB_LT) return lt_f64(a, b);
    if (OP == RFB_GT) return gt_f64(a, b);
    if (OP == RFB_LE) return !gt_f64(a, b);
    return !lt_f64(a, b);
}

// null-preserving widenings (core/ops.h:218-277)
__device__ __forceinline__ i64 widen_i64(u8 v) { return (i64)v; }
__device__ __forceinline__ i64 widen_i64(i16 v) { return v == NULL_I16 ? NULL_I64 : (i64)v; }
__device__ __forceinline__ i64 widen_i64(i32 v) { return v == NULL_I32 ? NULL_I64 : (i64)v; }
__device__ __forceinline__ i64 widen_i64(i64 v) { return v; }
__device__ __forceinline__ f64 widen_f64(u8 v) { return (f64)v; }
__device__ __forceinline__ f64 widen_f64(i16 v) { return v == NULL_I16 ? null_f64() : (f64)v; }
__device__ __forceinline__ f64 widen_f64(i32 v) { return v == NULL_I32 ? null_f64() : (f64)v; }
__device__ __forceinline__ f64 widen_f64(i64 v) { return v == NULL_I64 ? null_f64() : (f64)v; }
__device__ __forceinline__ f64 widen_f64(f64 v) { return v; }

// order-preserving u64 key of a double (core/sort.c:266-285): NaN -> 0, negatives bit-flipped, others sign set
__host__ __device__ __forceinline__ u64 f64_sort_key(f64 v) {
    u64 u = f64_bits(v);
    if ((u & 0x7FFFFFFFFFFFFFFFULL) > 0x7FF0000000000000ULL) return 0;
    return (u & 0x8000000000000000ULL) ? ~u : (u | 0x8000000000000000ULL);
}

// ------------------------------------------------------------------ 128-bit streaming loads

// One 16-byte read-only load that does not allocate in L1: the columns are streamed once.
struct __align__(16) vec16 { u64 lo, hi; };
__device__ __forceinline__ vec16 ld_stream16(const void *p) {
    vec16 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream16(void *p, vec16 v) {
    asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.lo), "l"(v.hi) : "memory");
}
template <typename T> __device__ __forceinline__ T ld_stream(const T *p) { return __ldcs(p); }

// a 16-byte vector viewed as 16/sizeof(T) elements of T
template <typename T> union Vec16 {
    static constexpr int N = 16 / (int)sizeof(T);
    vec16 raw;
    T e[N];
    __device__ __forceinline__ Vec16() {}
};

// ------------------------------------------------------------------ warp / block reductions (fixed trees => deterministic)

__device__ __forceinline__ i64 shfl_down(i64 v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ f64 shfl_down(f64 v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

struct OpAdd { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a + b; } };
struct OpAddWrap { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return (i64)((u64)a + (u64)b); } };
// null-skipping min/max (core/ops.h:178-188)
struct OpMinI { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return a == NULL_I64 ? b : (b == NULL_I64 ? a : (a < b ? a : b)); } };
struct OpMaxI { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return a == NULL_I64 ? b : (b == NULL_I64 ? a : (a > b ? a : b)); } };
struct OpMinF { __device__ __forceinline__ f64 operator()(f64 a, f64 b) const { return isnan64(a) ? b : (isnan64(b) ? a : (a < b ? a : b)); } };
struct OpMaxF { __device__ __forceinline__ f64 operator()(f64 a, f64 b) const { return isnan64(a) ? b : (isnan64(b) ? a : (a > b ? a : b)); } };

template <typename T, typename Op> __device__ __forceinline__ T warp_reduce(T v, Op op) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = op(v, shfl_down(v, d));
    return v;  // valid in lane 0
}

// Block-wide reduce for blockDim.x <= 1024; result valid in thread 0.  `smem` holds 32 T's.
template <typename T, typename Op> __device__ __forceinline__ T block_reduce(T v, Op op, T identity, T *smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    v = warp_reduce(v, op);
    __syncthreads();  // smem reuse across consecutive calls
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nwarps ? smem[lane] : identity;
        v = warp_reduce(v, op);
    }
    return v;
}

// ------------------------------------------------------------------ predicates against a constant

// Operand types the reference's comparison matrix accepts for vector operands (core/cmp.c:77-258): any mix of the plain
// numeric types I16/I32/I64/F64, a temporal/symbol type with itself, or DATE against TIMESTAMP (the date is converted to
// nanoseconds, core/cmp.c:243-257, core/ops.h:264).  B8/U8 compare as atoms only there: a type error at this layer.
#define RFB_NANOS_FROM_DAY 86400000000000LL
static inline bool rfb_cmp_types_ok(int xt, int yt) {
    auto plain = [](int t) { return t == RFB_I16 || t == RFB_I32 || t == RFB_I64 || t == RFB_F64; };
    if (plain(xt) && plain(yt)) return true;
    if ((xt == RFB_DATE && yt == RFB_TIMESTAMP) || (xt == RFB_TIMESTAMP && yt == RFB_DATE)) return true;
    return xt == yt && (xt == RFB_DATE || xt == RFB_TIME || xt == RFB_TIMESTAMP || xt == RFB_SYMBOL);
}
// multiplier that brings an operand of type t into the comparison unit of the pair (1 except DATE vs TIMESTAMP)
static inline i64 rfb_cmp_scale(int t, int other) { return (t == RFB_DATE && other == RFB_TIMESTAMP) ? RFB_NANOS_FROM_DAY : 1; }

// predicate = unsigned range test on an order-preserving 64-bit key, optionally negated.
// All six comparison operators against a constant reduce to it (see make_pred_range).
struct PredRange {
    u64 lo, span;
    u32 negate;
    u64 lo_int;  // lo ^ 2^63: integer columns test (x - lo_int) <= span directly, (x ^ S) - lo == x - (lo ^ S) mod 2^64
};

__host__ __device__ __forceinline__ u64 key_of_i64(i64 x) { return (u64)x ^ 0x8000000000000000ULL; }
// doubles: -0.0 == +0.0 must hold for comparisons (unlike the sort key), NaN is below everything and equals NaN
__host__ __device__ __forceinline__ u64 key_of_f64(f64 x) { return f64_sort_key(x == 0.0 ? 0.0 : x); }

template <typename P> __device__ __forceinline__ u64 pred_key(P x) {
    if constexpr (Elem<P>::kind == K_F64) return key_of_f64(x);
    else return key_of_i64(widen_i64(x));
}

__device__ __forceinline__ bool pred_test(u64 key, const PredRange &pr) { return ((key - pr.lo) <= pr.span) != (bool)pr.negate; }

// scalar -> i64 (null-preserving) or f64
static inline bool scalar_as_i64(const rfb_scalar_t *k, i64 *out) {
    switch (rfb_kind_of(k->type)) {
        case K_U8: *out = k->v.u8; return true;
        case K_I16: *out = k->v.i16 == NULL_I16 ? NULL_I64 : (i64)k->v.i16; return true;
        case K_I32: *out = k->v.i32 == NULL_I32 ? NULL_I64 : (i64)k->v.i32; return true;
        case K_I64: *out = k->v.i64; return true;
        default: return false;
    }
}
static inline bool scalar_as_f64(const rfb_scalar_t *k, f64 *out) {
    i64 t;
    if (rfb_kind_of(k->type) == K_F64) { *out = k->v.f64; return true; }
    if (!scalar_as_i64(k, &t)) return false;
    *out = (rfb_kind_of(k->type) != K_U8 && t == NULL_I64) ? null_f64() : (f64)t;
    return true;
}

// OP(x, k)  <=>  key(x) in [lo, lo+span] (xor negate)
static inline PredRange make_pred_range(int op, u64 kk) {
    PredRange pr;
    const u64 MAXK = ~0ULL;
    pr.negate = 0;
    switch (op) {
        case RFB_EQ: pr.lo = kk; pr.span = 0; break;
        case RFB_NE: pr.lo = kk; pr.span = 0; pr.negate = 1; break;
        case RFB_LE: pr.lo = 0; pr.span = kk; break;
        case RFB_GE: pr.lo = kk; pr.span = MAXK - kk; break;
        case RFB_LT: if (kk == 0) { pr.lo = 0; pr.span = MAXK; pr.negate = 1; } else { pr.lo = 0; pr.span = kk - 1; } break;
        default /*GT*/: if (kk == MAXK) { pr.lo = 0; pr.span = MAXK; pr.negate = 1; } else { pr.lo = kk + 1; pr.span = MAXK - kk - 1; } break;
    }
    pr.lo_int = pr.lo ^ 0x8000000000000000ULL;
    return pr;
}

// Build the range test for `OP(column, k)`; the constant is widened into the column's comparison domain (integers compare
// as i64 with null = minimum, anything against F64 compares as f64).  Returns false for unsupported type mixes.
static inline bool rfb_make_pred(int op, int col_type, const rfb_scalar_t *k, PredRange *pr) {
    const int ck = rfb_kind_of(col_type);
    if (op < RFB_EQ || op > RFB_GE || !ck || !k) return false;
    if (!rfb_cmp_types_ok(col_type, k->type) || rfb_cmp_scale(col_type, k->type) != 1 || rfb_cmp_scale(k->type, col_type) != 1) return false;  // unit-converting pairs go through rfb_cmp_dev
    if (ck == K_F64) {
        f64 kv;
        if (!scalar_as_f64(k, &kv)) return false;
        *pr = make_pred_range(op, key_of_f64(kv));
        return true;
    }
    i64 kv;
    if (!scalar_as_i64(k, &kv)) return false;   // integer column vs F64 constant: handled by the mask path (rfb_cmp_dev)
    *pr = make_pred_range(op, key_of_i64(kv));
    return true;
}
static inline bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

// splitmix64: the synthetic-column generator shared with the oracle (oracle/rf_oracle.c rfo_splitmix64)
__host__ __device__ __forceinline__ u64 splitmix64(u64 seed, u64 i) {
    u64 z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline int rfb_grid_for(const rfb_ctx_t *ctx, i64 work_items, int per_block, int blocks_per_sm) {
    i64 need = (work_items + per_block - 1) / per_block;
    i64 cap = (i64)ctx->sm_count * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}
