// rfb_map.cuh — shared pieces of the element-wise kernels (k_map.cu, k_binop_typed.cu): the two-input streaming map and the
// reference's scalar arithmetic (core/ops.h:125-197) as device functions.
#pragma once
#include "rfb_common.cuh"

// one case of the reference's arithmetic type matrix (binop_matrix.inc; k_binop_typed.cu holds the table and the kernels)
struct rfb_bincase_t { signed char op, form, xt, yt, lt, rt, ot, mt, fam, vt; int line; };
const rfb_bincase_t *rfb_binop_case(int op, int form, int xt, int yt);
int rfb_binop_matrix_dev(rfb_ctx_t *ctx, const rfb_bincase_t &c, const void *x, i64 xn, const rfb_scalar_t *xs, const void *y, i64 yn,
                         const rfb_scalar_t *ys, void *out);

namespace {


constexpr int THREADS = 256;
constexpr int BLOCKS_PER_SM = 4;

template <int BYTES> struct RawVec;
template <> struct RawVec<16> { typedef vec16 type; };
template <> struct RawVec<8> { typedef u64 type; };
template <> struct RawVec<4> { typedef u32 type; };
template <> struct RawVec<2> { typedef unsigned short type; };
template <> struct RawVec<1> { typedef u8 type; };

// R consecutive elements of T moved with one load/store of R*sizeof(T) bytes
template <typename T, int R> union Pack {
    typename RawVec<R * (int)sizeof(T)>::type raw;
    T e[R];
    __device__ __forceinline__ Pack() {}
};
template <typename T, int R> __device__ __forceinline__ void ld_pack(Pack<T, R> &p, const T *src) {
    if constexpr (R * sizeof(T) == 16) p.raw = ld_stream16(src);
    else p.raw = __ldcs(reinterpret_cast<const typename RawVec<R * (int)sizeof(T)>::type *>(src));
}
template <typename T, int R> __device__ __forceinline__ void st_pack(T *dst, const Pack<T, R> &p) {
    if constexpr (R * sizeof(T) == 16) st_stream16(dst, p.raw);
    else __stcs(reinterpret_cast<typename RawVec<R * (int)sizeof(T)>::type *>(dst), p.raw);
}

template <int A, int B> struct MaxI { static constexpr int v = A > B ? A : B; };
template <bool C, typename A, typename B> struct Sel { typedef A type; };
template <typename A, typename B> struct Sel<false, A, B> { typedef B type; };

// Generic two-input map.  F: out = f(x, y).  XA / YA: that side is an atom (broadcast), passed by value.
template <typename X, typename Y, typename O, bool XA, bool YA, typename F>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_map2(const X *__restrict__ x, X xa, const Y *__restrict__ y, Y ya, O *__restrict__ out, i64 n, bool vec_ok, F f) {
    constexpr int WX = XA ? 1 : (int)sizeof(X), WY = YA ? 1 : (int)sizeof(Y);
    constexpr int R = 16 / MaxI<MaxI<WX, WY>::v, (int)sizeof(O)>::v;   // rows per lane
    constexpr int UNROLL = 4;
    constexpr i64 TILE = (i64)THREADS * UNROLL * R;                     // rows per tile
    const i64 nvec = vec_ok ? (n / TILE) * TILE : 0;
    for (i64 base = (i64)blockIdx.x * TILE; base < nvec; base += (i64)gridDim.x * TILE) {
        typedef typename Sel<XA, u8, X>::type XV;  // an atom side needs no registers: its pack degenerates to bytes
        typedef typename Sel<YA, u8, Y>::type YV;
        Pack<XV, R> px[UNROLL];
        Pack<YV, R> py[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
            const i64 r = base + ((i64)j * THREADS + threadIdx.x) * R;
            if constexpr (!XA) ld_pack<X, R>(px[j], x + r);
            if constexpr (!YA) ld_pack<Y, R>(py[j], y + r);
        }
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
            const i64 r = base + ((i64)j * THREADS + threadIdx.x) * R;
            Pack<O, R> po;
#pragma unroll
            for (int e = 0; e < R; e++) {
                X a; Y b;
                if constexpr (XA) a = xa; else a = px[j].e[e];
                if constexpr (YA) b = ya; else b = py[j].e[e];
                po.e[e] = f(a, b);
            }
            st_pack<O, R>(out + r, po);
        }
    }
    // tail (and everything when a pointer is not 16-byte aligned)
    for (i64 r = nvec + (i64)blockIdx.x * THREADS + threadIdx.x; r < n; r += (i64)gridDim.x * THREADS)
        out[r] = f(XA ? xa : ld_stream(x + r), YA ? ya : ld_stream(y + r));
}

template <typename X, typename O, typename F>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_map1(const X *__restrict__ x, O *__restrict__ out, i64 n, bool vec_ok, F f) {
    constexpr int R = 16 / MaxI<(int)sizeof(X), (int)sizeof(O)>::v;
    constexpr int UNROLL = 8;
    constexpr i64 TILE = (i64)THREADS * UNROLL * R;
    const i64 nvec = vec_ok ? (n / TILE) * TILE : 0;
    for (i64 base = (i64)blockIdx.x * TILE; base < nvec; base += (i64)gridDim.x * TILE) {
        Pack<X, R> px[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; j++) ld_pack<X, R>(px[j], x + base + ((i64)j * THREADS + threadIdx.x) * R);
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
            Pack<O, R> po;
#pragma unroll
            for (int e = 0; e < R; e++) po.e[e] = f(px[j].e[e]);
            st_pack<O, R>(out + base + ((i64)j * THREADS + threadIdx.x) * R, po);
        }
    }
    for (i64 r = nvec + (i64)blockIdx.x * THREADS + threadIdx.x; r < n; r += (i64)gridDim.x * THREADS) out[r] = f(ld_stream(x + r));
}


template <typename X, typename Y, typename O, bool XA, bool YA, typename F>
int launch_map2(rfb_ctx_t *ctx, const void *x, X xa, const void *y, Y ya, void *out, i64 n, F f) {
    if (n == 0) return RFB_OK;
    const bool vec_ok = (XA || aligned16(x)) && (YA || aligned16(y)) && aligned16(out);
    const int grid = rfb_grid_for(ctx, n, THREADS * 8, BLOCKS_PER_SM);
    k_map2<X, Y, O, XA, YA, F><<<grid, THREADS, 0, ctx->stream>>>((const X *)x, xa, (const Y *)y, ya, (O *)out, n, vec_ok, f);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

// null-propagating scalar ops in each computation type (core/ops.h:153-177)
__device__ __forceinline__ i64 eucl_div64(i64 x, i64 y) {
    if (y == -1) return (i64)(0 - (u64)x);
    const i64 q = x / y, r = x - q * y;
    return q - ((((x < 0) != (y < 0)) && r != 0) ? 1 : 0);
}
__device__ __forceinline__ i32 eucl_div32(i32 x, i32 y) {
    if (y == -1) return (i32)(0 - (u32)x);
    const i32 q = x / y, r = x - q * y;
    return q - ((((x < 0) != (y < 0)) && r != 0) ? 1 : 0);
}
__device__ __forceinline__ i32 op_i32(int op, i32 x, i32 y) {
    if (x == NULL_I32 || y == NULL_I32) return NULL_I32;
    switch (op) {
        case RFB_ADD: return (i32)((u32)x + (u32)y);
        case RFB_SUB: return (i32)((u32)x - (u32)y);
        case RFB_MUL: return (i32)((u32)x * (u32)y);
        case RFB_DIV: return y == 0 ? NULL_I32 : eucl_div32(x, y);
        case RFB_XBAR: {  // XBARI32 (core/ops.h:193-194): C truncating division of the shifted value
            if (y == 0) return NULL_I32;
            const i32 t = x < 0 ? (i32)((u32)x + 1u - (u32)y) : x;
            return (i32)((u32)(y == -1 ? (i32)(0u - (u32)t) : t / y) * (u32)y);
        }
        default: return y == 0 ? NULL_I32 : (i32)((u32)x - (u32)eucl_div32(x, y) * (u32)y);
    }
}
__device__ __forceinline__ i64 op_i64(int op, i64 x, i64 y) {
    if (x == NULL_I64 || y == NULL_I64) return NULL_I64;
    switch (op) {
        case RFB_ADD: return (i64)((u64)x + (u64)y);
        case RFB_SUB: return (i64)((u64)x - (u64)y);
        case RFB_MUL: return (i64)((u64)x * (u64)y);
        case RFB_DIV: return y == 0 ? NULL_I64 : eucl_div64(x, y);
        case RFB_XBAR: {  // XBARI64 (core/ops.h:195-196)
            if (y == 0) return NULL_I64;
            const i64 t = x < 0 ? (i64)((u64)x + 1ULL - (u64)y) : x;
            return (i64)((u64)(y == -1 ? (i64)(0ULL - (u64)t) : t / y) * (u64)y);
        }
        default: return y == 0 ? NULL_I64 : (i64)((u64)x - (u64)eucl_div64(x, y) * (u64)y);
    }
}
__device__ __forceinline__ i64 f64_to_i64(f64 x);
// plain IEEE ops, never contracted into FMAs: the reference materialises every intermediate
__device__ __forceinline__ f64 op_f64(int op, f64 x, f64 y) {
    if (isnan64(x) || isnan64(y)) return null_f64();
    switch (op) {
        case RFB_ADD: return __dadd_rn(x, y);
        case RFB_SUB: return __dsub_rn(x, y);
        case RFB_MUL: return __dmul_rn(x, y);
        case RFB_DIV: return y == 0.0 ? null_f64() : floor(__ddiv_rn(x, y));
        case RFB_XBAR: {  // XBARF64 = FLOORF64(x / y) * y (core/ops.h:197,191): floor through an (i64) cast
            if (y == 0.0) return null_f64();   // the compiled reference yields NaN (inf * 0) for a zero bucket width
            const f64 q = __ddiv_rn(x, y);
            if (isnan64(q)) return null_f64();
            const f64 t = (f64)f64_to_i64(q);
            return __dmul_rn((q < 0.0 && t != q) ? __dsub_rn(t, 1.0) : t, y);
        }
        default: return y == 0.0 ? null_f64() : __dsub_rn(x, __dmul_rn(floor(__ddiv_rn(x, y)), y));
    }
}
__device__ __forceinline__ f64 op_fdiv(bool left_is_int, f64 x, f64 y) {
    if (left_is_int) {  // FDIVI64 applied to converted doubles (core/ops.h:173): null test against (double)INT64_MIN
        const f64 nul = -9223372036854775808.0;
        if (y == 0.0 || x == nul || y == nul || isnan64(y)) return null_f64();
        return __ddiv_rn(x, y);
    }
    if (y == 0.0 || isnan64(x) || isnan64(y)) return null_f64();
    return __ddiv_rn(x, y);
}
// f64 -> integer with the x86 cvttsd2si behaviour the reference compiles to: NaN / out of range -> INT_MIN (= null)
__device__ __forceinline__ i64 f64_to_i64(f64 x) {
    if (isnan64(x) || !(x > -9223372036854775808.0 && x < 9223372036854775808.0)) return NULL_I64;
    return (i64)x;
}
__device__ __forceinline__ i32 f64_to_i32(f64 x) {
    if (isnan64(x) || !(x > -2147483649.0 && x < 2147483648.0)) return NULL_I32;
    return (i32)x;
}
__device__ __forceinline__ i32 i64_to_i32(i64 x) { return x == NULL_I64 ? NULL_I32 : (i32)x; }

}  // namespace
