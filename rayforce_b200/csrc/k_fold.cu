// k_fold.cu — streaming scan + fold kernels (sm_100a).
//
//   rfb_fold_dev          ray_sum / ray_min / ray_max / ray_cnt        (reference core/math.c:1785-2045)
//   rfb_filter_fold_dev   select {(fold v) from t where (cmp p k)} fused into one pass: replaces
//                         ray_lt -> ray_where -> filter_collect -> ray_sum (core/cmp.c:335, core/ops.c:255,
//                         core/rayforce.c:1100, core/math.c:1874-1890)
//   rfb_fma_fold_dev      (fold (+ (* a b) c)) over three F64 columns
//   rfb_gather_fold_dev   ray_sum(MAPFILTER[col, ids]) without materialising the gather
//
// Shape of every kernel: persistent grid (SM count x resident CTAs), each CTA walks interleaved tiles; every thread
// issues all of its 128-bit L1-bypassing loads for a tile before consuming them (>= 128 B in flight per thread,
// ~128 KB per SM) so HBM latency is covered by memory-level parallelism; per-thread accumulators -> warp shuffle
// tree -> shared-memory tree -> one partial per CTA; the last CTA to finish (atomic ticket) folds the partials in
// index order and writes the result straight into mapped pinned host memory.  All trees are fixed, so fp64 results
// are run-to-run deterministic.  HBM roofline: algorithmic bytes = sizeof(elem) per row per distinct column.
#include "rfb_common.cuh"

namespace {

constexpr int THREADS = 256;       // gather / fma kernels
constexpr int BLOCKS_PER_SM = 4;

enum { FS_SUMCNT = RFB_F_SUM | RFB_F_CNT, FS_MINMAX = RFB_F_MIN | RFB_F_MAX, FS_ALL = RFB_F_SUM | RFB_F_CNT | RFB_F_MIN | RFB_F_MAX };

struct Partial {
    i64 rows, nonnull;
    u64 sum, mn, mx;  // bit patterns of i64 or f64 depending on the value kind
    u64 comp;         // f64 sums: bits of the compensation (low-order) term
};

// fp64 sums are carried as an unevaluated pair (hi, lo) and every addition is an error-free TwoSum (Knuth): the final
// hi + lo is the exact sum of the inputs rounded ONCE (barring catastrophic cancellation), i.e. within 1 ULP of the true
// sum no matter how the rows are split over threads, CTAs, chunks or GPUs — which is what makes the north_star's
// "within 1 ULP" meaningful against a CPU path whose own summation order is unspecified.  ~7 flops per row: free next
// to 8 bytes of HBM traffic.
struct dd { f64 hi, lo; };
__device__ __forceinline__ void two_sum_acc(f64 &hi, f64 &lo, f64 x) {
    const f64 t = __dadd_rn(hi, x);
    const f64 bp = __dsub_rn(t, hi);
    lo = __dadd_rn(lo, __dadd_rn(__dsub_rn(hi, __dsub_rn(t, bp)), __dsub_rn(x, bp)));
    hi = t;
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
    two_sum_acc(a.hi, a.lo, b.hi);
    a.lo = __dadd_rn(a.lo, b.lo);
    return a;
}
__device__ __forceinline__ dd dd_block_reduce(dd v, f64 *smem /* 64 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        dd o;
        o.hi = __shfl_down_sync(0xffffffffu, v.hi, d);
        o.lo = __shfl_down_sync(0xffffffffu, v.lo, d);
        v = dd_add(v, o);
    }
    __syncthreads();
    if (lane == 0) { smem[2 * warp] = v.hi; smem[2 * warp + 1] = v.lo; }
    __syncthreads();
    if (warp == 0) {
        v.hi = lane < nwarps ? smem[2 * lane] : 0.0;
        v.lo = lane < nwarps ? smem[2 * lane + 1] : 0.0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            dd o;
            o.hi = __shfl_down_sync(0xffffffffu, v.hi, d);
            o.lo = __shfl_down_sync(0xffffffffu, v.lo, d);
            v = dd_add(v, o);
        }
    }
    return v;  // valid in thread 0
}

// Per-thread accumulators.  Row counters are 32-bit (a thread sees far fewer than 2^31 rows; the launcher checks) because
// the scan kernels are bound by the INT32 pipe, not by HBM, as soon as the per-row instruction count reaches ~12
// (profiles/r01_bw_probe.txt): every 64-bit add saved is bandwidth gained.
//   LEAN: the predicate has been adjusted on the host so that it never selects a null of the (same) column, so no null test
//         is needed and nonnull == selected rows.
template <typename V, int FOLDS, bool COUNT_ROWS, bool LEAN> struct Acc {
    typedef typename Elem<V>::acc_t A;
    static constexpr bool FLT = (Elem<V>::kind == K_F64);
    u32 rows, nonnull;
    A sum, mn, mx;
    A comp;  // f64 sums only: running compensation term
    __device__ __forceinline__ static A min_identity() { if constexpr (FLT) return (A)bits_f64(0x7FF0000000000000ULL); else return (A)RFB_INF_I64; }
    __device__ __forceinline__ static A max_identity() { if constexpr (FLT) return (A)bits_f64(0xFFF0000000000000ULL); else return (A)NULL_I64; }
    __device__ __forceinline__ void init() { rows = 0; nonnull = 0; sum = (A)0; comp = (A)0; mn = min_identity(); mx = max_identity(); }
    __device__ __forceinline__ static A widen(V v) { if constexpr (FLT) return (A)widen_f64(v); else return (A)widen_i64(v); }
    // fold one element; `sel` = row passed the predicate.  Nulls are skipped (FOLD_ADD*, MIN*, MAX*, CNT*: core/ops.h)
    __device__ __forceinline__ void take(V v, bool sel) {
        const bool ok = LEAN ? sel : (sel && !Elem<V>::is_null(v));
        if (COUNT_ROWS) rows += sel;
        nonnull += ok;
        const A w = widen(v);
        if (FOLDS & RFB_F_SUM) {
            if constexpr (FLT) two_sum_acc(sum, comp, ok ? w : (A)0);   // adding 0.0 is exact: skipped rows leave (sum, comp) alone
            else sum = (A)((u64)sum + (ok ? (u64)w : 0ULL));            // wraps mod 2^64 like the reference's plain C add
        }
        if (FOLDS & RFB_F_MIN) { const A c = ok ? w : min_identity(); mn = c < mn ? c : mn; }
        if (FOLDS & RFB_F_MAX) { const A c = ok ? w : max_identity(); mx = c > mx ? c : mx; }
    }
};

__device__ __forceinline__ u64 to_bits(i64 v) { return (u64)v; }
__device__ __forceinline__ u64 to_bits(f64 v) { return f64_bits(v); }
template <typename A> __device__ __forceinline__ A from_bits(u64 b);
template <> __device__ __forceinline__ i64 from_bits<i64>(u64 b) { return (i64)b; }
template <> __device__ __forceinline__ f64 from_bits<f64>(u64 b) { return bits_f64(b); }

struct MinPlain { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return b < a ? b : a; } };
struct MaxPlain { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return b > a ? b : a; } };

// CTA-level finish shared by all fold kernels: block tree -> partial -> last CTA folds partials -> host result.
// vkind: element kind of the value column (decides how sum/min/max are reported, rfb200.h rfb_fold_t).
// rows_override >= 0: the selected-row count is known to the host (no predicate) ; -1: report the counted rows;
// -2: rows were not counted (null-excluding fast path) -> reported as -1.
template <typename A, int FOLDS>
__device__ __forceinline__ void finish_fold(i64 rows, i64 nonnull, A sum, A comp, A mn, A mx, A min_id, A max_id, int vkind, i64 rows_override,
                                            Partial *partials, u32 *ticket, rfb_fold_t *out) {
    constexpr bool FLT = (sizeof(A) == 8) && (A(0.5) != A(0));  // true for f64, false for i64
    __shared__ u64 red_smem[64];
    __shared__ bool is_last;
    rows = block_reduce<i64>(rows, OpAddWrap(), 0, (i64 *)red_smem);
    nonnull = block_reduce<i64>(nonnull, OpAddWrap(), 0, (i64 *)red_smem);
    if (FOLDS & RFB_F_SUM) {
        if constexpr (FLT) { dd r = dd_block_reduce(dd{(f64)sum, (f64)comp}, (f64 *)red_smem); sum = (A)r.hi; comp = (A)r.lo; }
        else sum = (A)block_reduce<i64>((i64)sum, OpAddWrap(), 0, (i64 *)red_smem);
    }
    if (FOLDS & RFB_F_MIN) mn = block_reduce<A>(mn, MinPlain(), min_id, (A *)red_smem);
    if (FOLDS & RFB_F_MAX) mx = block_reduce<A>(mx, MaxPlain(), max_id, (A *)red_smem);
    if (threadIdx.x == 0) {
        Partial p;
        p.rows = rows; p.nonnull = nonnull; p.sum = to_bits(sum); p.mn = to_bits(mn); p.mx = to_bits(mx); p.comp = to_bits(comp);
        partials[blockIdx.x] = p;
        __threadfence();
        const u32 t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last CTA: fixed-order fold of the per-CTA partials
    i64 r = 0, nn = 0; A s = (A)0, sc = (A)0, lo = min_id, hi = max_id;
    for (u32 b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        const Partial p = partials[b];
        r += p.rows; nn += p.nonnull;
        if (FOLDS & RFB_F_SUM) {
            if constexpr (FLT) { dd t = dd_add(dd{(f64)s, (f64)sc}, dd{bits_f64(p.sum), bits_f64(p.comp)}); s = (A)t.hi; sc = (A)t.lo; }
            else s = (A)((u64)s + p.sum);
        }
        if (FOLDS & RFB_F_MIN) { A c = from_bits<A>(p.mn); lo = c < lo ? c : lo; }
        if (FOLDS & RFB_F_MAX) { A c = from_bits<A>(p.mx); hi = c > hi ? c : hi; }
    }
    r = block_reduce<i64>(r, OpAddWrap(), 0, (i64 *)red_smem);
    nn = block_reduce<i64>(nn, OpAddWrap(), 0, (i64 *)red_smem);
    if (FOLDS & RFB_F_SUM) {
        if constexpr (FLT) { dd t = dd_block_reduce(dd{(f64)s, (f64)sc}, (f64 *)red_smem); s = (A)t.hi; sc = (A)t.lo; }
        else s = (A)block_reduce<i64>((i64)s, OpAddWrap(), 0, (i64 *)red_smem);
    }
    if (FOLDS & RFB_F_MIN) lo = block_reduce<A>(lo, MinPlain(), min_id, (A *)red_smem);
    if (FOLDS & RFB_F_MAX) hi = block_reduce<A>(hi, MaxPlain(), max_id, (A *)red_smem);
    if (threadIdx.x == 0) {
        rfb_fold_t res;
        res.rows = rows_override >= 0 ? rows_override : (rows_override == -2 ? -1 : r);
        res.nonnull = nn;
        res.sum_i64 = 0; res.sum_f64 = 0.0; res.min_i64 = res.max_i64 = 0; res.min_f64 = res.max_f64 = 0.0; res.sum_f64_err = 0.0;
        if constexpr (FLT) {
            res.sum_f64 = __dadd_rn((f64)s, (f64)sc);                                        // the one rounding
            res.sum_f64_err = __dadd_rn(__dsub_rn((f64)s, res.sum_f64), (f64)sc);            // what it dropped (for further merging)
            res.min_f64 = nn ? (f64)lo : null_f64();
            res.max_f64 = nn ? (f64)hi : null_f64();
        } else {
            i64 si = (i64)s, nul = NULL_I64;
            if (vkind == K_I32) { si = (i64)(i32)(u32)(u64)si; nul = (i64)NULL_I32; }  // I32/TIME sums wrap in 32 bits
            if (vkind == K_I16) nul = (i64)NULL_I16;
            if (vkind == K_U8) nul = 0;
            res.sum_i64 = si;
            res.min_i64 = nn ? (i64)lo : nul;
            res.max_i64 = nn ? (i64)hi : nul;
        }
        *out = res;
        *ticket = 0;  // re-arm for the next launch on this stream
        __threadfence_system();
    }
}

// ------------------------------------------------------------------ scan + fold over one or two columns

// Launch shape per fold set.  sum/count kernels are the bandwidth-critical ones: 512 threads x 4 CTAs = all 2048 thread
// slots of the SM (<= 32 registers), 4 x 16 B in flight per thread (128 KB per SM); the min/max and all-folds kernels
// carry more live state and keep 256 threads x 4 CTAs x 8 loads.
#ifndef SC_THREADS   // overridable for tuning sweeps (tools/sweep_scan_cfg.sh)
#define SC_THREADS 1024
#define SC_BPS 2
#define SC_LOADS 4
#endif
template <int FOLDS> struct ScanCfg {
    static constexpr int THREADS = (FOLDS == FS_SUMCNT) ? SC_THREADS : 256;
    static constexpr int BPS = (FOLDS == FS_SUMCNT) ? SC_BPS : 4;
    static constexpr int LOADS = (FOLDS == FS_SUMCNT) ? SC_LOADS : 8;   // 16-byte loads in flight per thread
};

// integer predicate test with the bias folded into the constant: (x ^ S) - lo == x - (lo ^ S)  (mod 2^64)
template <typename P> __device__ __forceinline__ bool pred_sel(P x, const PredRange &pr) {
    if constexpr (Elem<P>::kind == K_F64) return pred_test(key_of_f64(x), pr);
    else return (((u64)widen_i64(x) - pr.lo_int) <= pr.span) != (bool)pr.negate;
}

template <typename P, typename V, int FOLDS, bool HAS_PRED, bool SAME, bool LEAN>
__global__ void __launch_bounds__(ScanCfg<FOLDS>::THREADS, ScanCfg<FOLDS>::BPS)
k_scan_fold(const P *__restrict__ pred, PredRange pr, const V *__restrict__ val, i64 n, i64 chunks, int vkind,
            Partial *partials, u32 *ticket, rfb_fold_t *out) {
    constexpr int THREADS_ = ScanCfg<FOLDS>::THREADS;
    constexpr bool TWO = HAS_PRED && !SAME;
    constexpr int SZ_MIN = TWO ? (sizeof(P) < sizeof(V) ? sizeof(P) : sizeof(V)) : sizeof(V);
    constexpr int RPT = 16 / SZ_MIN;                       // rows per thread-chunk
    constexpr int NV_V = RPT * (int)sizeof(V) / 16;         // 16-byte loads per chunk, value column
    constexpr int NV_P = TWO ? RPT * (int)sizeof(P) / 16 : 0;
    constexpr int UNROLL = (ScanCfg<FOLDS>::LOADS / (NV_V + NV_P)) > 0 ? ScanCfg<FOLDS>::LOADS / (NV_V + NV_P) : 1;
    constexpr int TILE = THREADS_ * UNROLL;                 // chunks per tile
    constexpr bool COUNT_ROWS = HAS_PRED && !LEAN;

    Acc<V, FOLDS, COUNT_ROWS, LEAN> acc;
    acc.init();

    // full tiles only: no bounds checks, no predicated loads in the hot loop
    const i64 tiles = chunks / TILE;
    for (i64 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const i64 c0 = tile * TILE + threadIdx.x;
        Vec16<V> vv[UNROLL][NV_V];
        Vec16<P> pv[UNROLL][TWO ? NV_P : 1];
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
            const i64 c = c0 + (i64)j * THREADS_;
            const char *vb = (const char *)val + c * (RPT * sizeof(V));
#pragma unroll
            for (int q = 0; q < NV_V; q++) vv[j][q].raw = ld_stream16(vb + 16 * q);
            if (TWO) {
                const char *pb = (const char *)pred + c * (RPT * sizeof(P));
#pragma unroll
                for (int q = 0; q < NV_P; q++) pv[j][q].raw = ld_stream16(pb + 16 * q);
            }
        }
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
#pragma unroll
            for (int r = 0; r < RPT; r++) {
                const V v = vv[j][r / Vec16<V>::N].e[r % Vec16<V>::N];
                bool sel = true;
                if (HAS_PRED) {
                    if (SAME) sel = pred_sel<P>(*reinterpret_cast<const P *>(&v), pr);
                    else sel = pred_sel<P>(pv[j][r / Vec16<P>::N].e[r % Vec16<P>::N], pr);
                }
                acc.take(v, sel);
            }
        }
    }
    // the remaining rows: less than one tile (and everything, when a pointer is not 16-byte aligned: chunks == 0)
    for (i64 r = tiles * TILE * RPT + (i64)blockIdx.x * THREADS_ + threadIdx.x; r < n; r += (i64)gridDim.x * THREADS_) {
        const V v = ld_stream(val + r);
        bool sel = true;
        if (HAS_PRED) sel = pred_sel<P>(SAME ? *reinterpret_cast<const P *>(&v) : ld_stream(pred + r), pr);
        acc.take(v, sel);
    }
    typedef typename Acc<V, FOLDS, COUNT_ROWS, LEAN>::A A;
    finish_fold<A, FOLDS>((i64)acc.rows, (i64)acc.nonnull, acc.sum, acc.comp, acc.mn, acc.mx, Acc<V, FOLDS, COUNT_ROWS, LEAN>::min_identity(),
                          Acc<V, FOLDS, COUNT_ROWS, LEAN>::max_identity(), vkind, HAS_PRED ? (LEAN ? -2 : -1) : n, partials, ticket, out);
}

// ------------------------------------------------------------------ (fold (+ (* a b) c)) over three F64 columns

template <int FOLDS>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_fma_fold(const f64 *__restrict__ a, const f64 *__restrict__ b, const f64 *__restrict__ c, i64 n, i64 chunks,
           Partial *partials, u32 *ticket, rfb_fold_t *out) {
    constexpr int UNROLL = 2;  // 3 columns x 2 loads = 6 x 16 B in flight per thread
    constexpr int TILE = THREADS * UNROLL;
    Acc<f64, FOLDS, false, false> acc;
    acc.init();
    // MULF64 then ADDF64 (core/ops.h:155,164): NaN in -> NaN out; the products/sums of non-NaN values are plain IEEE
    // ops (no fused multiply-add: the reference materialises a*b, rounding it, before adding c).
    auto eval = [](f64 x, f64 y, f64 z) -> f64 {
        const f64 m = (isnan64(x) || isnan64(y)) ? null_f64() : __dmul_rn(x, y);
        return (isnan64(m) || isnan64(z)) ? null_f64() : __dadd_rn(m, z);
    };
    const i64 tiles = (chunks + TILE - 1) / TILE;
    for (i64 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        Vec16<f64> va[UNROLL], vb[UNROLL], vc[UNROLL];
        const i64 c0 = tile * TILE + threadIdx.x;
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
            const i64 ch = c0 + (i64)j * THREADS;
            if (ch < chunks) {
                va[j].raw = ld_stream16(a + 2 * ch);
                vb[j].raw = ld_stream16(b + 2 * ch);
                vc[j].raw = ld_stream16(c + 2 * ch);
            }
        }
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
            const i64 ch = c0 + (i64)j * THREADS;
            if (ch < chunks) {
                acc.take(eval(va[j].e[0], vb[j].e[0], vc[j].e[0]), true);
                acc.take(eval(va[j].e[1], vb[j].e[1], vc[j].e[1]), true);
            }
        }
    }
    for (i64 r = chunks * 2 + (i64)blockIdx.x * THREADS + threadIdx.x; r < n; r += (i64)gridDim.x * THREADS)
        acc.take(eval(ld_stream(a + r), ld_stream(b + r), ld_stream(c + r)), true);
    finish_fold<f64, FOLDS>(0, (i64)acc.nonnull, acc.sum, acc.comp, acc.mn, acc.mx, Acc<f64, FOLDS, false, false>::min_identity(),
                            Acc<f64, FOLDS, false, false>::max_identity(), K_F64, n, partials, ticket, out);
}

// ------------------------------------------------------------------ fold through a selection vector (MAPFILTER)

template <typename V, int FOLDS>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_gather_fold(const V *__restrict__ col, const i64 *__restrict__ ids, i64 m, int vkind, Partial *partials, u32 *ticket,
              rfb_fold_t *out) {
    Acc<V, FOLDS, false, false> acc;
    acc.init();
    constexpr int UNROLL = 4;
    const i64 stride = (i64)gridDim.x * THREADS;
    i64 i = (i64)blockIdx.x * THREADS + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < m; i += UNROLL * stride) {
        i64 id[UNROLL];
        V v[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; j++) id[j] = ld_stream(ids + i + j * stride);
#pragma unroll
        for (int j = 0; j < UNROLL; j++) v[j] = __ldg(col + id[j]);
#pragma unroll
        for (int j = 0; j < UNROLL; j++) acc.take(v[j], true);
    }
    for (; i < m; i += stride) acc.take(__ldg(col + ld_stream(ids + i)), true);
    typedef typename Acc<V, FOLDS, false, false>::A A;
    finish_fold<A, FOLDS>(0, (i64)acc.nonnull, acc.sum, acc.comp, acc.mn, acc.mx, Acc<V, FOLDS, false, false>::min_identity(),
                          Acc<V, FOLDS, false, false>::max_identity(), vkind, m, partials, ticket, out);
}

// ------------------------------------------------------------------ host-side dispatch

inline int foldset_of(int folds) {
    folds &= ~RFB_F_ROWS;
    if (!(folds & ~FS_SUMCNT)) return FS_SUMCNT;
    if (!(folds & ~FS_MINMAX)) return FS_MINMAX;
    return FS_ALL;
}

struct Scratch {
    Partial *partials;
    u32 *ticket;
    rfb_fold_t *result;  // mapped pinned host memory
};
inline Scratch scratch_of(rfb_ctx_t *ctx) {
    Scratch s;
    s.ticket = (u32 *)ctx->d_scratch;
    s.partials = (Partial *)((char *)ctx->d_scratch + 256);
    s.result = ctx->result_override ? (rfb_fold_t *)ctx->result_override : (rfb_fold_t *)ctx->h_result + ctx->result_slot;
    return s;
}


template <typename P, typename V, int FOLDS, bool HAS_PRED, bool SAME, bool LEAN>
int launch_scan_fold(rfb_ctx_t *ctx, const void *pred, PredRange pr, const void *val, i64 n, int vkind) {
    constexpr bool TWO = HAS_PRED && !SAME;
    constexpr int SZ_MIN = TWO ? (sizeof(P) < sizeof(V) ? sizeof(P) : sizeof(V)) : sizeof(V);
    constexpr int RPT = 16 / SZ_MIN;
    constexpr int THREADS_ = ScanCfg<FOLDS>::THREADS;
    const bool vec_ok = aligned16(val) && (!TWO || aligned16(pred));
    const i64 chunks = vec_ok ? n / RPT : 0;
    const i64 work = vec_ok ? (chunks + ScanCfg<FOLDS>::LOADS - 1) / ScanCfg<FOLDS>::LOADS + 1 : n;  // thread-work estimate for small grids
    const int grid = rfb_grid_for(ctx, work, THREADS_, ScanCfg<FOLDS>::BPS);
    if (n / ((i64)grid * THREADS_) >= (1ll << 30)) {   // per-thread row counters are 32-bit
        rfb_set_error("fold: column of %lld rows is beyond the per-launch limit", (long long)n);
        return RFB_ERR_ARG;
    }
    Scratch s = scratch_of(ctx);
    k_scan_fold<P, V, FOLDS, HAS_PRED, SAME, LEAN><<<grid, THREADS_, 0, ctx->stream>>>(
        (const P *)pred, pr, (const V *)val, n, chunks, vkind, s.partials, s.ticket, s.result);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

template <typename P, typename V, bool HAS_PRED, bool SAME>
int dispatch_foldset(rfb_ctx_t *ctx, int fs, bool lean, const void *pred, PredRange pr, const void *val, i64 n, int vkind) {
    switch (fs) {
        case FS_SUMCNT:
            if constexpr (HAS_PRED && SAME && Elem<V>::kind != K_F64) {
                if (lean) return launch_scan_fold<P, V, FS_SUMCNT, HAS_PRED, SAME, true>(ctx, pred, pr, val, n, vkind);
            }
            return launch_scan_fold<P, V, FS_SUMCNT, HAS_PRED, SAME, false>(ctx, pred, pr, val, n, vkind);
        case FS_MINMAX: return launch_scan_fold<P, V, FS_MINMAX, HAS_PRED, SAME, false>(ctx, pred, pr, val, n, vkind);
        default: return launch_scan_fold<P, V, FS_ALL, HAS_PRED, SAME, false>(ctx, pred, pr, val, n, vkind);
    }
}

template <typename P>
int dispatch_val(rfb_ctx_t *ctx, int fs, bool lean, const void *pred, PredRange pr, int vkind, const void *val, i64 n, bool same) {
    switch (vkind) {
        case K_I32:
            if (same && Elem<P>::kind == K_I32) return dispatch_foldset<i32, i32, true, true>(ctx, fs, lean, pred, pr, val, n, vkind);
            return dispatch_foldset<P, i32, true, false>(ctx, fs, false, pred, pr, val, n, vkind);
        case K_I64:
            if (same && Elem<P>::kind == K_I64) return dispatch_foldset<i64, i64, true, true>(ctx, fs, lean, pred, pr, val, n, vkind);
            return dispatch_foldset<P, i64, true, false>(ctx, fs, false, pred, pr, val, n, vkind);
        case K_F64:
            if (same && Elem<P>::kind == K_F64) return dispatch_foldset<f64, f64, true, true>(ctx, fs, false, pred, pr, val, n, vkind);
            return dispatch_foldset<P, f64, true, false>(ctx, fs, false, pred, pr, val, n, vkind);
        default:
            rfb_set_error("filter+fold: unsupported value type");
            return RFB_ERR_TYPE;
    }
}

// Make the predicate reject the column's own null (biased key 0) so the kernel can drop the per-row null test:
// sum/nonnull are unchanged, only the count of selected NULL rows is lost (hence not done when RFB_F_ROWS is requested).
bool exclude_null_key(PredRange *pr) {
    if (!pr->negate) {
        if (pr->lo != 0) return true;               // the null key is outside the selected range already
        if (pr->span == 0) return false;            // selects exactly the nulls
        pr->lo = 1; pr->span -= 1;
    } else {
        if (pr->lo == 0) return true;               // the null key is inside the rejected range already
        if (pr->lo != 1) return false;
        pr->lo = 0; pr->span += 1;                  // grow the rejected range over key 0
    }
    pr->lo_int = pr->lo ^ 0x8000000000000000ULL;
    return true;
}

int wait_result(rfb_ctx_t *ctx, rfb_fold_t *out) {
    if (!out) return RFB_OK;  // asynchronous form: the caller collects with rfb_fold_result()
    if (ctx->result_override) { rfb_set_error("fold results are redirected (rfb_ctx_set_result_ptr): pass out == NULL"); return RFB_ERR_ARG; }
    RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out, (rfb_fold_t *)ctx->h_result + ctx->result_slot, sizeof(rfb_fold_t));
    return RFB_OK;
}

}  // namespace

// launch only (result lands in ctx->h_result after the stream drains); used by the host-layer pipeline too
int rfb_fold_launch(rfb_ctx_t *ctx, int folds, int type, const void *x, i64 n) {
    const int fs = foldset_of(folds), vk = rfb_kind_of(type);
    PredRange pr = {0, 0, 0, 0};
    if (type == RFB_B8 || type == RFB_SYMBOL) { rfb_set_error("fold: unsupported type %d", type); return RFB_ERR_TYPE; }
    switch (vk) {
        case K_U8: return dispatch_foldset<u8, u8, false, false>(ctx, fs, false, nullptr, pr, x, n, vk);
        case K_I16: return dispatch_foldset<i16, i16, false, false>(ctx, fs, false, nullptr, pr, x, n, vk);
        case K_I32: return dispatch_foldset<i32, i32, false, false>(ctx, fs, false, nullptr, pr, x, n, vk);
        case K_I64: return dispatch_foldset<i64, i64, false, false>(ctx, fs, false, nullptr, pr, x, n, vk);
        case K_F64: return dispatch_foldset<f64, f64, false, false>(ctx, fs, false, nullptr, pr, x, n, vk);
        default: rfb_set_error("fold: unsupported type %d", type); return RFB_ERR_TYPE;
    }
}

int rfb_filter_fold_launch(rfb_ctx_t *ctx, int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k, int folds,
                           int val_type, const void *val, i64 n) {
    const int fs = foldset_of(folds), pk = rfb_kind_of(pred_type), vk = rfb_kind_of(val_type);
    if (cmp_op < RFB_EQ || cmp_op > RFB_GE) { rfb_set_error("bad comparison op %d", cmp_op); return RFB_ERR_ARG; }
    const bool same = (pred == val) && (pk == vk);
    if (!rfb_cmp_types_ok(pred_type, k->type)) { rfb_set_error("filter+fold: unsupported comparison types %d, %d", pred_type, k->type); return RFB_ERR_TYPE; }
    if (pk == K_I32 || pk == K_I64) {
        i64 kv;
        if (!scalar_as_i64(k, &kv)) { rfb_set_error("filter+fold: integer column vs non-integer constant"); return RFB_ERR_TYPE; }
        PredRange pr = make_pred_range(cmp_op, key_of_i64(kv));
        const bool lean = same && fs == FS_SUMCNT && !(folds & RFB_F_ROWS) && exclude_null_key(&pr);
        return pk == K_I32 ? dispatch_val<i32>(ctx, fs, lean, pred, pr, vk, val, n, same)
                           : dispatch_val<i64>(ctx, fs, lean, pred, pr, vk, val, n, same);
    }
    if (pk == K_F64) {
        f64 kv;
        if (!scalar_as_f64(k, &kv)) { rfb_set_error("filter+fold: bad constant type"); return RFB_ERR_TYPE; }
        PredRange pr = make_pred_range(cmp_op, key_of_f64(kv));
        return dispatch_val<f64>(ctx, fs, false, pred, pr, vk, val, n, same);
    }
    rfb_set_error("filter+fold: unsupported predicate column type %d", pred_type);
    return RFB_ERR_TYPE;
}

extern "C" int rfb_fold_dev(rfb_ctx_t *ctx, int folds, int type, const void *x, int64_t n, rfb_fold_t *out) {
    RFB_ARG(ctx && n >= 0 && (x || n == 0), "rfb_fold_dev");
    int rc = rfb_fold_launch(ctx, folds, type, x, n);
    if (rc) return rc;
    return wait_result(ctx, out);
}

extern "C" int rfb_filter_fold_dev(rfb_ctx_t *ctx, int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k,
                                   int folds, int val_type, const void *val, int64_t n, rfb_fold_t *out) {
    RFB_ARG(ctx && k && n >= 0 && ((pred && val) || n == 0), "rfb_filter_fold_dev");
    int rc = rfb_filter_fold_launch(ctx, cmp_op, pred_type, pred, k, folds, val_type, val, n);
    if (rc) return rc;
    return wait_result(ctx, out);
}

extern "C" int rfb_fma_fold_dev(rfb_ctx_t *ctx, int folds, const double *a, const double *b, const double *c, int64_t n,
                                rfb_fold_t *out) {
    RFB_ARG(ctx && n >= 0 && ((a && b && c) || n == 0), "rfb_fma_fold_dev");
    const bool vec_ok = aligned16(a) && aligned16(b) && aligned16(c);
    const i64 chunks = vec_ok ? n / 2 : 0;
    const int grid = rfb_grid_for(ctx, vec_ok ? chunks / 2 + 1 : n, THREADS, BLOCKS_PER_SM);
    Scratch s = scratch_of(ctx);
    switch (foldset_of(folds)) {
        case FS_SUMCNT: k_fma_fold<FS_SUMCNT><<<grid, THREADS, 0, ctx->stream>>>(a, b, c, n, chunks, s.partials, s.ticket, s.result); break;
        case FS_MINMAX: k_fma_fold<FS_MINMAX><<<grid, THREADS, 0, ctx->stream>>>(a, b, c, n, chunks, s.partials, s.ticket, s.result); break;
        default: k_fma_fold<FS_ALL><<<grid, THREADS, 0, ctx->stream>>>(a, b, c, n, chunks, s.partials, s.ticket, s.result); break;
    }
    RFB_CHECK_LAUNCH(ctx);
    return wait_result(ctx, out);
}

template <typename V> static int gather_fold_t(rfb_ctx_t *ctx, int fs, const void *col, const i64 *ids, i64 m, int vk) {
    const int grid = rfb_grid_for(ctx, m / 4 + 1, THREADS, BLOCKS_PER_SM);
    Scratch s = scratch_of(ctx);
    switch (fs) {
        case FS_SUMCNT: k_gather_fold<V, FS_SUMCNT><<<grid, THREADS, 0, ctx->stream>>>((const V *)col, ids, m, vk, s.partials, s.ticket, s.result); break;
        case FS_MINMAX: k_gather_fold<V, FS_MINMAX><<<grid, THREADS, 0, ctx->stream>>>((const V *)col, ids, m, vk, s.partials, s.ticket, s.result); break;
        default: k_gather_fold<V, FS_ALL><<<grid, THREADS, 0, ctx->stream>>>((const V *)col, ids, m, vk, s.partials, s.ticket, s.result); break;
    }
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

extern "C" int rfb_gather_fold_dev(rfb_ctx_t *ctx, int folds, int type, const void *col, const int64_t *ids, int64_t m,
                                   rfb_fold_t *out) {
    RFB_ARG(ctx && m >= 0 && ((col && ids) || m == 0), "rfb_gather_fold_dev");
    const int fs = foldset_of(folds), vk = rfb_kind_of(type);
    int rc;
    if (type == RFB_B8 || type == RFB_SYMBOL) { rfb_set_error("fold: unsupported type %d", type); return RFB_ERR_TYPE; }
    switch (vk) {
        case K_U8: rc = gather_fold_t<u8>(ctx, fs, col, ids, m, vk); break;
        case K_I16: rc = gather_fold_t<i16>(ctx, fs, col, ids, m, vk); break;
        case K_I32: rc = gather_fold_t<i32>(ctx, fs, col, ids, m, vk); break;
        case K_I64: rc = gather_fold_t<i64>(ctx, fs, col, ids, m, vk); break;
        case K_F64: rc = gather_fold_t<f64>(ctx, fs, col, ids, m, vk); break;
        default: rfb_set_error("fold: unsupported type %d", type); return RFB_ERR_TYPE;
    }
    if (rc) return rc;
    return wait_result(ctx, out);
}

// ------------------------------------------------------------------ compound predicates: (and p1 p2 ..) / (or p1 p2 ..)  + fold
//
// `where: (and (< x k1) (>= y k2) ...)` in the reference evaluates every conjunct to a 1 B/row mask, combines them
// (core/logic.c:34-110), materialises ids and gathers.  Here up to 4 range predicates over 8-byte
// columns (I64-kind or F64) are tested in the same pass that folds the value column: (npred + 1) x 8 B per row, nothing
// materialised.  A predicate column that is also the value column is still loaded once.

namespace {

constexpr int MAX_PREDS = 4;
struct MultiPred {
    const u64 *col[MAX_PREDS];
    PredRange pr[MAX_PREDS];
    u32 is_f64[MAX_PREDS];
    int npred;
    int conj;  // 1 = and, 0 = or
};

__device__ __forceinline__ bool mp_test(u64 raw, const PredRange &pr, u32 is_f64) {
    const u64 key = is_f64 ? key_of_f64(bits_f64(raw)) : (raw ^ 0x8000000000000000ULL);
    return pred_test(key, pr);
}

template <typename V, int FOLDS, int NP>
__global__ void __launch_bounds__(256, 4)
k_multi_filter_fold(MultiPred mp, const V *__restrict__ val, i64 n, i64 chunks, int vkind, Partial *partials, u32 *ticket, rfb_fold_t *out) {
    constexpr int T = 256, UNROLL = (NP >= 3) ? 1 : 2;   // (NP + 1) x UNROLL 16-byte loads in flight per thread
    constexpr int TILE = T * UNROLL;
    Acc<V, FOLDS, true, false> acc;
    acc.init();
    const i64 tiles = chunks / TILE;
    for (i64 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        Vec16<V> vv[UNROLL];
        Vec16<u64> pv[UNROLL][NP];
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
            const i64 c = tile * TILE + (i64)j * T + threadIdx.x;
            vv[j].raw = ld_stream16((const char *)val + c * 16);
#pragma unroll
            for (int q = 0; q < NP; q++) pv[j][q].raw = ld_stream16((const char *)mp.col[q] + c * 16);
        }
#pragma unroll
        for (int j = 0; j < UNROLL; j++)
#pragma unroll
            for (int r = 0; r < 2; r++) {
                bool sel = mp.conj != 0;
#pragma unroll
                for (int q = 0; q < NP; q++) {
                    const bool t = mp_test(pv[j][q].e[r], mp.pr[q], mp.is_f64[q]);
                    sel = mp.conj ? (sel && t) : (sel || t);
                }
                acc.take(vv[j].e[r], sel);
            }
    }
    for (i64 r = tiles * TILE * 2 + (i64)blockIdx.x * T + threadIdx.x; r < n; r += (i64)gridDim.x * T) {
        bool sel = mp.conj != 0;
#pragma unroll
        for (int q = 0; q < NP; q++) {
            const bool t = mp_test(ld_stream(mp.col[q] + r), mp.pr[q], mp.is_f64[q]);
            sel = mp.conj ? (sel && t) : (sel || t);
        }
        acc.take(ld_stream(val + r), sel);
    }
    typedef typename Acc<V, FOLDS, true, false>::A A;
    finish_fold<A, FOLDS>((i64)acc.rows, (i64)acc.nonnull, acc.sum, acc.comp, acc.mn, acc.mx, Acc<V, FOLDS, true, false>::min_identity(),
                          Acc<V, FOLDS, true, false>::max_identity(), vkind, -1, partials, ticket, out);
}

template <typename V, int FOLDS>
int launch_multi(rfb_ctx_t *ctx, const MultiPred &mp, const void *val, i64 n, int vkind) {
    bool vec_ok = aligned16(val);
    for (int q = 0; q < mp.npred; q++) vec_ok = vec_ok && aligned16(mp.col[q]);
    const i64 chunks = vec_ok ? n / 2 : 0;
    const int grid = rfb_grid_for(ctx, vec_ok ? chunks / 2 + 1 : n, 256, 4);
    if (n / ((i64)grid * 256) >= (1ll << 30)) { rfb_set_error("fold: column of %lld rows is beyond the per-launch limit", (long long)n); return RFB_ERR_ARG; }
    Scratch s = scratch_of(ctx);
    switch (mp.npred) {
        case 1: k_multi_filter_fold<V, FOLDS, 1><<<grid, 256, 0, ctx->stream>>>(mp, (const V *)val, n, chunks, vkind, s.partials, s.ticket, s.result); break;
        case 2: k_multi_filter_fold<V, FOLDS, 2><<<grid, 256, 0, ctx->stream>>>(mp, (const V *)val, n, chunks, vkind, s.partials, s.ticket, s.result); break;
        case 3: k_multi_filter_fold<V, FOLDS, 3><<<grid, 256, 0, ctx->stream>>>(mp, (const V *)val, n, chunks, vkind, s.partials, s.ticket, s.result); break;
        default: k_multi_filter_fold<V, FOLDS, 4><<<grid, 256, 0, ctx->stream>>>(mp, (const V *)val, n, chunks, vkind, s.partials, s.ticket, s.result); break;
    }
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

template <typename V>
int launch_multi_fs(rfb_ctx_t *ctx, int fs, const MultiPred &mp, const void *val, i64 n, int vkind) {
    switch (fs) {
        case FS_SUMCNT: return launch_multi<V, FS_SUMCNT>(ctx, mp, val, n, vkind);
        case FS_MINMAX: return launch_multi<V, FS_MINMAX>(ctx, mp, val, n, vkind);
        default: return launch_multi<V, FS_ALL>(ctx, mp, val, n, vkind);
    }
}

}  // namespace

extern "C" int rfb_multi_filter_fold_dev(rfb_ctx_t *ctx, int npred, const rfb_pred_t *preds, int conjunction, int folds, int val_type,
                                         const void *val, int64_t n, rfb_fold_t *out) {
    RFB_ARG(ctx && preds && npred >= 1 && npred <= MAX_PREDS && n >= 0 && (val || n == 0), "rfb_multi_filter_fold_dev");
    MultiPred mp;
    memset(&mp, 0, sizeof(mp));
    mp.npred = npred;
    mp.conj = conjunction ? 1 : 0;
    for (int q = 0; q < npred; q++) {
        const int pk = rfb_kind_of(preds[q].type);
        if (pk != K_I64 && pk != K_F64) { rfb_set_error("compound filter: predicate columns must be 8-byte (I64-kind or F64), got type %d", preds[q].type); return RFB_ERR_TYPE; }
        if (!rfb_make_pred(preds[q].op, preds[q].type, &preds[q].k, &mp.pr[q])) { rfb_set_error("compound filter: unsupported comparison %d on types %d, %d", preds[q].op, preds[q].type, preds[q].k.type); return RFB_ERR_TYPE; }
        RFB_ARG(preds[q].col || n == 0, "rfb_multi_filter_fold_dev: predicate column");
        mp.col[q] = (const u64 *)preds[q].col;
        mp.is_f64[q] = pk == K_F64;
    }
    const int fs = foldset_of(folds), vk = rfb_kind_of(val_type);
    int rc;
    if (vk == K_I64) rc = launch_multi_fs<i64>(ctx, fs, mp, val, n, vk);
    else if (vk == K_F64) rc = launch_multi_fs<f64>(ctx, fs, mp, val, n, vk);
    else { rfb_set_error("compound filter: value column must be I64-kind or F64, got type %d", val_type); return RFB_ERR_TYPE; }
    if (rc) return rc;
    return wait_result(ctx, out);
}

extern "C" int rfb_fold_result(rfb_ctx_t *ctx, rfb_fold_t *out) {
    RFB_ARG(ctx && out, "rfb_fold_result");
    return wait_result(ctx, out);
}
