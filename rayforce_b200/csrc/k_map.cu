// k_map.cu — element-wise column kernels (sm_100a): one output element per row, no cross-row dependence.
//
//   rfb_cmp_dev       ray_eq/ne/lt/gt/le/ge -> cmp_map                 (reference core/cmp.c:35-68 loops, :335-683 driver)
//   rfb_binop_dev     ray_add/sub/mul/div/fdiv/mod -> binop_map        (core/math.c:55-90 loops, :251-1782 matrix, :2280-2345)
//   rfb_unop_f64_dev  ray_round/floor/ceil -> unop_map                 (core/math.c:2047-2117, core/ops.h:190-192)
//
// All three are pure HBM streams (cmp: 8+8 B in, 1 B out per row; binop: 8+8 in, 8 out).  Shape: each thread owns
// "lanes" of R = 16 / sizeof(widest operand) consecutive rows; lane j of thread t in a tile covers rows
// (tile*TILE + j*THREADS + t)*R .. +R, so every load instruction of a warp touches one contiguous 512-byte span (and the
// narrower operand a contiguous 512/k span).  All loads of a tile are issued before the first use (UNROLL vectors in
// flight per operand per thread) and bypass L1 allocation.  Grid = SM count x resident CTAs, grid-stride over tiles.
#include "rfb_map.cuh"

namespace {

// ------------------------------------------------------------------ comparisons

// order-preserving u64 keys: integers compare as plain values (null = smallest); doubles put NaN below
// everything, NaN == NaN and -0.0 == +0.0 (core/ops.h:74-127)
__host__ __device__ __forceinline__ u64 ckey_i64(i64 v) { return (u64)v ^ 0x8000000000000000ULL; }
__host__ __device__ __forceinline__ u64 ckey_f64(f64 v) { return f64_sort_key(v == 0.0 ? 0.0 : v); }

// `scale` converts units (DATE days -> TIMESTAMP nanoseconds), null-preserving like date_to_timestamp (core/ops.h:264)
template <bool FLT, typename T> __device__ __forceinline__ u64 ckey(T v, i64 scale) {
    if constexpr (FLT || Elem<T>::kind == K_F64) return ckey_f64(widen_f64(v));
    else {
        const i64 w = widen_i64(v);
        return ckey_i64(w == NULL_I64 ? w : w * scale);
    }
}

// 3-bit truth table indexed by sign(a - b) + 1: bit0 a<b, bit1 a==b, bit2 a>b
inline u32 cmp_lut(int op) {
    switch (op) {
        case RFB_EQ: return 0b010; case RFB_NE: return 0b101; case RFB_LT: return 0b001;
        case RFB_GT: return 0b100; case RFB_LE: return 0b011; default: return 0b110;
    }
}
inline int mirror_op(int op) {  // x OP y  <=>  y OP' x
    switch (op) { case RFB_LT: return RFB_GT; case RFB_GT: return RFB_LT; case RFB_LE: return RFB_GE; case RFB_GE: return RFB_LE; default: return op; }
}

template <bool FLT, typename X, typename Y> struct CmpVV {
    u32 lut;
    i64 sx, sy;
    __device__ __forceinline__ u8 operator()(X a, Y b) const {
        const u64 ka = ckey<FLT>(a, sx), kb = ckey<FLT>(b, sy);
        const int idx = (ka > kb) - (ka < kb) + 1;
        return (u8)((lut >> idx) & 1u);
    }
};
// vector vs atom: the atom is already a key
template <bool FLT, typename X> struct CmpVA {
    u32 lut;
    u64 kb;
    i64 sx;
    __device__ __forceinline__ u8 operator()(X a, u8) const {
        const u64 ka = ckey<FLT>(a, sx);
        const int idx = (ka > kb) - (ka < kb) + 1;
        return (u8)((lut >> idx) & 1u);
    }
};

template <bool FLT, typename X>
int cmp_va(rfb_ctx_t *ctx, int op, const void *x, i64 n, u64 kb, i64 sx, u8 *mask) {
    CmpVA<FLT, X> f{cmp_lut(op), kb, sx};
    return launch_map2<X, u8, u8, false, true>(ctx, x, X(), nullptr, (u8)0, mask, n, f);
}
template <typename X>
int cmp_va_x(rfb_ctx_t *ctx, int op, bool flt, const void *x, i64 n, u64 kb, i64 sx, u8 *mask) {
    return flt ? cmp_va<true, X>(ctx, op, x, n, kb, sx, mask) : cmp_va<false, X>(ctx, op, x, n, kb, sx, mask);
}

template <typename X, typename Y>
int cmp_vv(rfb_ctx_t *ctx, int op, const void *x, const void *y, i64 n, i64 sx, i64 sy, u8 *mask) {
    constexpr bool FLT = Elem<X>::kind == K_F64 || Elem<Y>::kind == K_F64;
    CmpVV<FLT, X, Y> f{cmp_lut(op), sx, sy};
    return launch_map2<X, Y, u8, false, false>(ctx, x, X(), y, Y(), mask, n, f);
}
template <typename X>
int cmp_vv_y(rfb_ctx_t *ctx, int op, const void *x, int ky, const void *y, i64 n, i64 sx, i64 sy, u8 *mask) {
    switch (ky) {
        case K_I16: return cmp_vv<X, i16>(ctx, op, x, y, n, sx, sy, mask);
        case K_I32: return cmp_vv<X, i32>(ctx, op, x, y, n, sx, sy, mask);
        case K_I64: return cmp_vv<X, i64>(ctx, op, x, y, n, sx, sy, mask);
        default: return cmp_vv<X, f64>(ctx, op, x, y, n, sx, sy, mask);
    }
}

// scalar -> widened value (host)
bool scalar_i64(const rfb_scalar_t *s, i64 *out) {
    switch (rfb_kind_of(s->type)) {
        case K_U8: *out = s->v.u8; return true;
        case K_I16: *out = s->v.i16 == NULL_I16 ? NULL_I64 : (i64)s->v.i16; return true;
        case K_I32: *out = s->v.i32 == NULL_I32 ? NULL_I64 : (i64)s->v.i32; return true;
        case K_I64: *out = s->v.i64; return true;
        default: return false;
    }
}
bool scalar_f64(const rfb_scalar_t *s, f64 *out) {
    if (rfb_kind_of(s->type) == K_F64) { *out = s->v.f64; return true; }
    i64 t;
    if (!scalar_i64(s, &t)) return false;
    *out = (rfb_kind_of(s->type) != K_U8 && t == NULL_I64) ? null_f64() : (f64)t;
    return true;
}

}  // namespace

extern "C" int rfb_cmp_dev(rfb_ctx_t *ctx, int op, int xt, const void *x, int64_t xn, const rfb_scalar_t *xs, int yt,
                           const void *y, int64_t yn, const rfb_scalar_t *ys, uint8_t *mask) {
    RFB_ARG(ctx && op >= RFB_EQ && op <= RFB_GE, "rfb_cmp_dev: op");
    RFB_ARG((xn >= 0 ? (x || xn == 0) : xs != nullptr) && (yn >= 0 ? (y || yn == 0) : ys != nullptr), "rfb_cmp_dev: operands");
    if (!rfb_cmp_types_ok(xt, yt)) { rfb_set_error("cmp: unsupported operand types %d, %d", xt, yt); return RFB_ERR_TYPE; }
    if (xn >= 0 && yn >= 0 && xn != yn) { rfb_set_error("cmp: vector lengths differ (%lld vs %lld)", (long long)xn, (long long)yn); return RFB_ERR_LENGTH; }
    const int kx = rfb_kind_of(xt), ky = rfb_kind_of(yt);
    const bool flt = (kx == K_F64 || ky == K_F64);
    const i64 sx = rfb_cmp_scale(xt, yt), sy = rfb_cmp_scale(yt, xt);
    if (xn < 0 && yn < 0) {  // atom vs atom: one byte, computed by the same kernel on a 1-element broadcast
        RFB_ARG(mask, "rfb_cmp_dev: mask");
        rfb_set_error("cmp: both operands are atoms; the operator layer folds constants on the host");
        return RFB_ERR_ARG;
    }
    if (xn >= 0 && yn >= 0) {
        RFB_ARG(mask || xn == 0, "rfb_cmp_dev: mask");
        switch (kx) {
            case K_I16: return cmp_vv_y<i16>(ctx, op, x, ky, y, xn, sx, sy, mask);
            case K_I32: return cmp_vv_y<i32>(ctx, op, x, ky, y, xn, sx, sy, mask);
            case K_I64: return cmp_vv_y<i64>(ctx, op, x, ky, y, xn, sx, sy, mask);
            default: return cmp_vv_y<f64>(ctx, op, x, ky, y, xn, sx, sy, mask);
        }
    }
    // vector vs atom (atom on the left: mirror the operator)
    const bool atom_left = xn < 0;
    const void *v = atom_left ? y : x;
    const i64 n = atom_left ? yn : xn;
    const int kv = atom_left ? ky : kx;
    const rfb_scalar_t *s = atom_left ? xs : ys;
    const int vop = atom_left ? mirror_op(op) : op;
    RFB_ARG(mask || n == 0, "rfb_cmp_dev: mask");
    u64 kb;
    if (flt) { f64 d; if (!scalar_f64(s, &d)) return RFB_ERR_TYPE; kb = ckey_f64(d); }
    else {
        i64 d;
        if (!scalar_i64(s, &d)) return RFB_ERR_TYPE;
        const i64 sa = atom_left ? sx : sy;
        kb = ckey_i64(d == NULL_I64 ? d : d * sa);
    }
    const i64 sv = atom_left ? sy : sx;
    switch (kv) {
        case K_I16: return cmp_va_x<i16>(ctx, vop, flt, v, n, kb, sv, mask);
        case K_I32: return cmp_va_x<i32>(ctx, vop, flt, v, n, kb, sv, mask);
        case K_I64: return cmp_va_x<i64>(ctx, vop, flt, v, n, kb, sv, mask);
        default: return cmp_va_x<f64>(ctx, vop, flt, v, n, kb, sv, mask);
    }
}

// ------------------------------------------------------------------ arithmetic

namespace {

template <typename M> struct Conv;  // widen an operand into computation type M
template <> struct Conv<i32> { template <typename T> __device__ __forceinline__ static i32 of(T v) { return (i32)v; } };
template <> struct Conv<i64> { template <typename T> __device__ __forceinline__ static i64 of(T v) { return widen_i64(v); } };
template <> struct Conv<f64> { template <typename T> __device__ __forceinline__ static f64 of(T v) { return widen_f64(v); } };

// X, Y operand element types; M computation type; O output type
template <typename X, typename Y, typename M, typename O> struct BinOp {
    int op;
    bool left_is_int;
    __device__ __forceinline__ O operator()(X a, Y b) const {
        if constexpr (Elem<M>::kind == K_F64) {
            const f64 x = Conv<f64>::of(a), y = Conv<f64>::of(b);
            if (op == RFB_FDIV) return (O)op_fdiv(left_is_int, x, y);
            const f64 r = op_f64(op, x, y);
            if constexpr (Elem<O>::kind == K_F64) return r;
            else if constexpr (Elem<O>::kind == K_I64) return f64_to_i64(r);
            else return f64_to_i32(r);
        } else if constexpr (Elem<M>::kind == K_I64) {
            const i64 r = op_i64(op, Conv<i64>::of(a), Conv<i64>::of(b));
            if constexpr (Elem<O>::kind == K_I64) return r;
            else if constexpr (Elem<O>::kind == K_I32) return i64_to_i32(r);
            else return (O)r;
        } else {
            return (O)op_i32(op, (i32)a, (i32)b);
        }
    }
};

// ---- i64 column (/ | % | xbar) i64 ATOM: the divisor is a launch constant, so the 64-bit hardware division (a ~100-instruction
// emulation on the integer pipe: 5.7 ms per 1e9 rows, issue-bound) becomes one 64x64 -> high-64 multiply, a subtract, an add and
// two shifts (the round-up magic number of Granlund & Montgomery in its branch-free form, computed once on the host with a
// 128-bit division).  Signs are peeled off first; the quotient semantics stay the reference's: `/` and `%` floor (EUCL_DIV,
// core/ops.h:165-171), xbar truncates its shifted operand (XBARI64, core/ops.h:195-196).
struct DivMagic { u64 m; int more; };            // |y| >= 2: q = t >> more with t = ((n - hi(m*n)) >> 1) + hi(m*n)
__device__ __forceinline__ u64 udiv_magic(u64 n, const DivMagic &g) {
    const u64 q = __umul64hi(g.m, n);
    return (((n - q) >> 1) + q) >> g.more;
}
static DivMagic div_magic_of(u64 d) {            // d >= 2
    DivMagic g;
    const int fl = 63 - __builtin_clzll(d);
    if ((d & (d - 1)) == 0) { g.m = 0; g.more = fl - 1; return g; }            // power of two: t = n >> 1, then >> (log2 - 1)
    const unsigned __int128 num = (unsigned __int128)1 << (64 + fl);
    u64 pm = (u64)(num / d);
    const u64 rem = (u64)(num % d);
    pm += pm;
    const u64 twice = rem + rem;
    if (twice >= d || twice < rem) pm += 1;
    g.m = pm + 1;
    g.more = fl;
    return g;
}
struct DivConstI64 {                             // y != 0, y != NULL, |y| >= 2
    int op;
    i64 y;
    u64 ay;
    DivMagic g;
    __device__ __forceinline__ i64 operator()(i64 x, i64) const {
        if (x == NULL_I64) return NULL_I64;
        if (op == RFB_XBAR) {
            const i64 t = x < 0 ? (i64)((u64)x + 1ULL - (u64)y) : x;
            const u64 at = t < 0 ? 0ULL - (u64)t : (u64)t;
            const u64 q = udiv_magic(at, g);
            const i64 sq = ((t < 0) != (y < 0)) ? (i64)(0ULL - q) : (i64)q;    // truncating quotient
            return (i64)((u64)sq * (u64)y);
        }
        const u64 ax = x < 0 ? 0ULL - (u64)x : (u64)x;
        const u64 q = udiv_magic(ax, g), r = ax - q * ay;
        const i64 fq = ((x < 0) != (y < 0)) ? (i64)(0ULL - q - (r != 0 ? 1ULL : 0ULL)) : (i64)q;   // floor quotient
        return op == RFB_DIV ? fq : (i64)((u64)x - (u64)fq * (u64)y);
    }
};

// the same for the 32-bit operator family (DIVI32 / MODI32 / XBARI32: TIME / DATE / I32 columns by a constant — `xbar time 60000`
// is THE temporal bucketing operation): |t| < 2^31 is widened, divided with the 64-bit magic number and narrowed; the 32-bit
// wrap-around of `x + 1 - y` inside xbar happens before the widening, as in the reference's int arithmetic
struct DivConstI32 {                             // y != 0, y != NULL_I32, |y| >= 2
    int op;
    i32 y;
    u64 ay;
    DivMagic g;
    __device__ __forceinline__ i32 operator()(i32 x, i32) const {
        if (x == NULL_I32) return NULL_I32;
        if (op == RFB_XBAR) {
            const i32 t = x < 0 ? (i32)((u32)x + 1u - (u32)y) : x;
            const u64 at = t < 0 ? 0ULL - (u64)(i64)t : (u64)t;
            const u64 q = udiv_magic(at, g);
            const i32 sq = ((t < 0) != (y < 0)) ? (i32)(0u - (u32)q) : (i32)(u32)q;      // truncating quotient
            return (i32)((u32)sq * (u32)y);
        }
        const u64 ax = x < 0 ? 0ULL - (u64)(i64)x : (u64)x;
        const u64 q = udiv_magic(ax, g), r = ax - q * ay;
        const i32 fq = ((x < 0) != (y < 0)) ? (i32)(0u - (u32)q - (r != 0 ? 1u : 0u)) : (i32)(u32)q;   // floor quotient
        return op == RFB_DIV ? fq : (i32)((u32)x - (u32)fq * (u32)y);
    }
};

// a 32-bit column by an I64 atom (`xbar time 60000`: the literal is an i64): the reference widens the column (i32_to_i64 keeps
// nullness), works in the 64-bit family and narrows the result (i64_to_time ...: nullness, then an (i32) cast)
struct DivConstI64Narrow {
    DivConstI64 f;
    __device__ __forceinline__ i32 operator()(i32 x, i32) const {
        const i64 r = f(x == NULL_I32 ? NULL_I64 : (i64)x, 0);
        return r == NULL_I64 ? NULL_I32 : (i32)r;
    }
};

// ... and when that I64 atom fits 32 bits (it nearly always does: `xbar time 60000`), every intermediate of the reference's 64-bit
// arithmetic on a widened 32-bit operand stays below 2^32 in magnitude, so the division runs on a 32-bit magic number (one
// 32-bit multiply-high) and the narrowing (i32) cast of the 64-bit result equals the wrapped 32-bit product
struct DivMagic32 { u32 m; int more; };
__device__ __forceinline__ u32 udiv_magic32(u32 n, const DivMagic32 &g) {
    const u32 q = __umulhi(g.m, n);
    return (((n - q) >> 1) + q) >> g.more;
}
static DivMagic32 div_magic32_of(u32 d) {        // 2 <= d < 2^31
    DivMagic32 g;
    const int fl = 31 - __builtin_clz(d);
    if ((d & (d - 1)) == 0) { g.m = 0; g.more = fl - 1; return g; }
    const u64 num = 1ULL << (32 + fl);
    u32 pm = (u32)(num / d);
    const u32 rem = (u32)(num % d);
    pm += pm;
    const u32 twice = rem + rem;
    if (twice >= d || twice < rem) pm += 1;
    g.m = pm + 1;
    g.more = fl;
    return g;
}
struct DivConstI64Narrow32 {                     // y: an I64 atom with 2 <= |y| < 2^31
    int op;
    i64 y;
    u32 ay;
    DivMagic32 g;
    __device__ __forceinline__ i32 operator()(i32 x, i32) const {
        if (x == NULL_I32) return NULL_I32;
        if (op == RFB_XBAR) {
            const i64 t = x < 0 ? (i64)x + 1 - y : (i64)x;                  // |t| < 2^32
            const u32 at = (u32)(t < 0 ? -t : t);
            const u32 q = udiv_magic32(at, g);
            const u32 sq = ((t < 0) != (y < 0)) ? 0u - q : q;                // truncating quotient, mod 2^32
            return (i32)(sq * (u32)y);
        }
        const u32 ax = (u32)(x < 0 ? -(i64)x : (i64)x);
        const u32 q = udiv_magic32(ax, g), r = ax - q * ay;
        const u32 fq = ((x < 0) != (y < 0)) ? 0u - q - (r != 0 ? 1u : 0u) : q;   // floor quotient, mod 2^32
        return op == RFB_DIV ? (i32)fq : (i32)((u32)x - fq * (u32)y);
    }
};

// result typing: the per-case macro arguments of core/math.c:251-1782 / infer_*_type core/math.c:92-223
bool binop_types(int op, int xt, int yt, int *mt, int *ot) {
    const bool okx = (xt == RFB_I32 || xt == RFB_I64 || xt == RFB_F64), oky = (yt == RFB_I32 || yt == RFB_I64 || yt == RFB_F64);
    if (!okx || !oky) return false;
    const bool anyf = (xt == RFB_F64 || yt == RFB_F64), any64 = (xt == RFB_I64 || yt == RFB_I64);
    const int wide = anyf ? RFB_F64 : any64 ? RFB_I64 : RFB_I32;
    switch (op) {
        case RFB_ADD: case RFB_SUB: case RFB_MUL: case RFB_XBAR: *mt = wide; *ot = wide; return true;   // xbar: infer_xbar_type core/math.c:225-249
        case RFB_DIV: *mt = wide; *ot = xt; return true;                    // keeps the LEFT operand's type
        case RFB_FDIV: *mt = RFB_F64; *ot = RFB_F64; return true;
        case RFB_MOD: *mt = wide; *ot = anyf ? RFB_F64 : yt; return true;   // integer % integer: RIGHT operand's type
        default: return false;
    }
}

template <typename T> T scalar_as(const rfb_scalar_t *s) {
    if (!s) return T();
    switch (rfb_kind_of(s->type)) {
        case K_I32: return (T)s->v.i32;
        case K_I64: return (T)s->v.i64;
        case K_F64: return (T)s->v.f64;
        default: return T();
    }
}

template <typename X, typename Y, typename M, typename O>
int binop_form(rfb_ctx_t *ctx, int op, bool left_is_int, const void *x, i64 xn, const rfb_scalar_t *xs, const void *y,
               i64 yn, const rfb_scalar_t *ys, void *out) {
    BinOp<X, Y, M, O> f{op, left_is_int};
    const i64 n = xn >= 0 ? xn : yn;
    if (xn >= 0 && yn >= 0) return launch_map2<X, Y, O, false, false>(ctx, x, X(), y, Y(), out, n, f);
    if (xn >= 0) return launch_map2<X, Y, O, false, true>(ctx, x, X(), nullptr, scalar_as<Y>(ys), out, n, f);
    return launch_map2<X, Y, O, true, false>(ctx, nullptr, scalar_as<X>(xs), y, Y(), out, n, f);
}

template <typename X, typename Y, typename M>
int binop_out(rfb_ctx_t *ctx, int op, int ot, bool lii, const void *x, i64 xn, const rfb_scalar_t *xs, const void *y, i64 yn,
              const rfb_scalar_t *ys, void *out) {
    switch (ot) {
        case RFB_I32: return binop_form<X, Y, M, i32>(ctx, op, lii, x, xn, xs, y, yn, ys, out);
        case RFB_I64: return binop_form<X, Y, M, i64>(ctx, op, lii, x, xn, xs, y, yn, ys, out);
        default: return binop_form<X, Y, M, f64>(ctx, op, lii, x, xn, xs, y, yn, ys, out);
    }
}

template <typename X, typename Y>
int binop_xy(rfb_ctx_t *ctx, int op, int mt, int ot, bool lii, const void *x, i64 xn, const rfb_scalar_t *xs, const void *y,
             i64 yn, const rfb_scalar_t *ys, void *out) {
    constexpr int kx = Elem<X>::kind, ky = Elem<Y>::kind;
    // only the (M, O) pairs binop_types can produce for this (X, Y) are instantiated
    if (mt == RFB_F64) {
        if constexpr (kx == K_F64 || ky == K_F64) return binop_out<X, Y, f64>(ctx, op, ot, lii, x, xn, xs, y, yn, ys, out);
        else return binop_form<X, Y, f64, f64>(ctx, op, lii, x, xn, xs, y, yn, ys, out);  // fdiv of two integer columns
    }
    if (mt == RFB_I64) {
        if constexpr (kx != K_F64 && ky != K_F64 && (kx == K_I64 || ky == K_I64)) {
            if (ot == RFB_I32) return binop_form<X, Y, i64, i32>(ctx, op, lii, x, xn, xs, y, yn, ys, out);
            return binop_form<X, Y, i64, i64>(ctx, op, lii, x, xn, xs, y, yn, ys, out);
        }
    }
    if constexpr (kx == K_I32 && ky == K_I32) return binop_form<i32, i32, i32, i32>(ctx, op, lii, x, xn, xs, y, yn, ys, out);
    rfb_set_error("binop: internal type dispatch");
    return RFB_ERR_TYPE;
}

template <typename X>
int binop_x(rfb_ctx_t *ctx, int op, int mt, int ot, int yt, bool lii, const void *x, i64 xn, const rfb_scalar_t *xs,
            const void *y, i64 yn, const rfb_scalar_t *ys, void *out) {
    switch (yt) {
        case RFB_I32: return binop_xy<X, i32>(ctx, op, mt, ot, lii, x, xn, xs, y, yn, ys, out);
        case RFB_I64: return binop_xy<X, i64>(ctx, op, mt, ot, lii, x, xn, xs, y, yn, ys, out);
        default: return binop_xy<X, f64>(ctx, op, mt, ot, lii, x, xn, xs, y, yn, ys, out);
    }
}

}  // namespace

namespace {
inline bool plain_num(int t) { return t == RFB_I32 || t == RFB_I64 || t == RFB_F64; }
inline bool is_i64_like(int t) { return t == RFB_I64 || t == RFB_TIMESTAMP; }
inline bool is_i32_like(int t) { return t == RFB_I32 || t == RFB_DATE || t == RFB_TIME; }
}  // namespace

// result vector type per operand form (0 vector-vector, 1 vector-atom, 2 atom-vector) over the reference's full type matrix
extern "C" int rfb_binop_type_form(int op, int form, int xt, int yt) {
    if (plain_num(xt) && plain_num(yt)) return rfb_binop_type(op, xt, yt);
    const rfb_bincase_t *c = rfb_binop_case(op, form, xt, yt);
    return c ? c->vt : RFB_ERR_TYPE;
}

extern "C" int rfb_binop_type(int op, int xt, int yt) {
    int mt, ot;
    return binop_types(op, xt, yt, &mt, &ot) ? ot : RFB_ERR_TYPE;
}

extern "C" int rfb_binop_dev(rfb_ctx_t *ctx, int op, int xt, const void *x, int64_t xn, const rfb_scalar_t *xs, int yt,
                             const void *y, int64_t yn, const rfb_scalar_t *ys, void *out) {
    RFB_ARG(ctx, "rfb_binop_dev: ctx");
    int mt, ot;
    if (!(plain_num(xt) && plain_num(yt))) {       // the full type matrix
        if (xn < 0 && yn < 0) { rfb_set_error("binop: both operands are atoms; the operator layer folds constants on the host"); return RFB_ERR_ARG; }
        const rfb_bincase_t *c = rfb_binop_case(op, xn >= 0 ? (yn >= 0 ? 0 : 1) : 2, xt, yt);
        if (!c) { rfb_set_error("binop %d: unsupported operand types %d, %d", op, xt, yt); return RFB_ERR_TYPE; }
        if (xn >= 0 && yn >= 0 && xn != yn) { rfb_set_error("binop: vector lengths differ (%lld vs %lld)", (long long)xn, (long long)yn); return RFB_ERR_LENGTH; }
        RFB_ARG((xn >= 0 ? (x || xn == 0) : xs != nullptr) && (yn >= 0 ? (y || yn == 0) : ys != nullptr), "rfb_binop_dev: operands");
        RFB_ARG(out || (xn >= 0 ? xn : yn) == 0, "rfb_binop_dev: out");
        if ((xn < 0 && xs->type != xt) || (yn < 0 && ys->type != yt)) { rfb_set_error("binop: scalar type tag does not match operand type"); return RFB_ERR_ARG; }
        if (c->form == 1 && c->fam == RFB_I64 && is_i64_like(c->lt) && is_i64_like(c->rt) && is_i64_like(c->mt) && is_i64_like(c->ot) &&
            (op == RFB_DIV || op == RFB_MOD || op == RFB_XBAR)) {
            const i64 d = ys->v.i64;   // timestamp xbar / div / mod by a constant: the same 64-bit kernel as i64 by an atom
            if (d != 0 && d != NULL_I64 && d != 1 && d != -1) {
                const u64 ad = d < 0 ? 0ULL - (u64)d : (u64)d;
                DivConstI64 f{op, d, ad, div_magic_of(ad)};
                return launch_map2<i64, i64, i64, false, true>(ctx, x, i64(), nullptr, d, out, xn, f);
            }
        }
        if (c->form == 1 && c->fam == RFB_I32 && is_i32_like(c->lt) && is_i32_like(c->mt) && is_i32_like(c->ot) &&
            (is_i32_like(c->rt) || is_i64_like(c->rt)) && (op == RFB_DIV || op == RFB_MOD || op == RFB_XBAR)) {
            // the atom as the kernel would see it: rt_to_mt keeps nullness and narrows (i64_to_time ..., core/ops.h:247-252)
            const i32 d = is_i32_like(c->rt) ? ys->v.i32 : (ys->v.i64 == NULL_I64 ? NULL_I32 : (i32)ys->v.i64);
            if (d != 0 && d != NULL_I32 && d != 1 && d != -1) {
                const u64 ad = d < 0 ? 0ULL - (u64)(i64)d : (u64)d;
                DivConstI32 f{op, d, ad, div_magic_of(ad)};
                return launch_map2<i32, i32, i32, false, true>(ctx, x, i32(), nullptr, d, out, xn, f);
            }
        }
        if (c->form == 1 && c->fam == RFB_I64 && is_i32_like(c->lt) && c->mt == RFB_I64 && is_i64_like(c->rt) && is_i32_like(c->ot) &&
            (op == RFB_DIV || op == RFB_MOD || op == RFB_XBAR)) {
            const i64 d = ys->v.i64;
            if (d != 0 && d != NULL_I64 && d != 1 && d != -1) {
                const u64 ad = d < 0 ? 0ULL - (u64)d : (u64)d;
                if (ad < (1ULL << 31)) {
                    DivConstI64Narrow32 f32{op, d, (u32)ad, div_magic32_of((u32)ad)};
                    return launch_map2<i32, i32, i32, false, true>(ctx, x, i32(), nullptr, 0, out, xn, f32);
                }
                DivConstI64Narrow f{DivConstI64{op, d, ad, div_magic_of(ad)}};
                return launch_map2<i32, i32, i32, false, true>(ctx, x, i32(), nullptr, 0, out, xn, f);
            }
        }
        return rfb_binop_matrix_dev(ctx, *c, x, xn, xs, y, yn, ys, out);
    }
    if (!binop_types(op, xt, yt, &mt, &ot)) { rfb_set_error("binop %d: unsupported operand types %d, %d", op, xt, yt); return RFB_ERR_TYPE; }
    if (xn >= 0 && yn >= 0 && xn != yn) { rfb_set_error("binop: vector lengths differ (%lld vs %lld)", (long long)xn, (long long)yn); return RFB_ERR_LENGTH; }
    if (xn < 0 && yn < 0) { rfb_set_error("binop: both operands are atoms; the operator layer folds constants on the host"); return RFB_ERR_ARG; }
    RFB_ARG((xn >= 0 ? (x || xn == 0) : xs != nullptr) && (yn >= 0 ? (y || yn == 0) : ys != nullptr), "rfb_binop_dev: operands");
    RFB_ARG(out || (xn >= 0 ? xn : yn) == 0, "rfb_binop_dev: out");
    if ((xn < 0 && xs->type != xt) || (yn < 0 && ys->type != yt)) { rfb_set_error("binop: scalar type tag does not match operand type"); return RFB_ERR_ARG; }
    const bool lii = xt != RFB_F64;
    if (xn >= 0 && yn < 0 && xt == RFB_I64 && yt == RFB_I64 && (op == RFB_DIV || op == RFB_MOD || op == RFB_XBAR)) {
        const i64 d = ys->v.i64;
        if (d != 0 && d != NULL_I64 && d != 1 && d != -1) {       // 0 / null / +-1: the generic kernel (nulls, identity, negation)
            const u64 ad = d < 0 ? 0ULL - (u64)d : (u64)d;
            DivConstI64 f{op, d, ad, div_magic_of(ad)};
            return launch_map2<i64, i64, i64, false, true>(ctx, x, i64(), nullptr, d, out, xn, f);
        }
    }
    if (xn >= 0 && yn < 0 && xt == RFB_I32 && yt == RFB_I32 && (op == RFB_DIV || op == RFB_MOD || op == RFB_XBAR)) {
        const i32 d = ys->v.i32;
        if (d != 0 && d != NULL_I32 && d != 1 && d != -1) {
            const u64 ad = d < 0 ? 0ULL - (u64)(i64)d : (u64)d;
            DivConstI32 f{op, d, ad, div_magic_of(ad)};
            return launch_map2<i32, i32, i32, false, true>(ctx, x, i32(), nullptr, d, out, xn, f);
        }
    }
    switch (xt) {
        case RFB_I32: return binop_x<i32>(ctx, op, mt, ot, yt, lii, x, xn, xs, y, yn, ys, out);
        case RFB_I64: return binop_x<i64>(ctx, op, mt, ot, yt, lii, x, xn, xs, y, yn, ys, out);
        default: return binop_x<f64>(ctx, op, mt, ot, yt, lii, x, xn, xs, y, yn, ys, out);
    }
}

// ------------------------------------------------------------------ round / floor / ceil

namespace {
// core/ops.h:190-192: the reference rounds through an (i64) cast
__device__ __forceinline__ f64 trunc_via_i64(f64 v) { return (f64)f64_to_i64(v); }
__device__ __forceinline__ f64 floor_ref(f64 v) {
    const f64 t = trunc_via_i64(v);
    return (v < 0.0 && t != v) ? __dsub_rn(t, 1.0) : t;
}
struct UnopF64 {
    int op;
    __device__ __forceinline__ f64 operator()(f64 v) const {
        if (isnan64(v)) return null_f64();
        switch (op) {
            case RFB_ROUND: return v >= 0.0 ? trunc_via_i64(__dadd_rn(v, 0.5)) : trunc_via_i64(__dsub_rn(v, 0.5));
            case RFB_FLOOR: return floor_ref(v);
            default: return -floor_ref(-v);
        }
    }
};
}  // namespace

extern "C" int rfb_unop_f64_dev(rfb_ctx_t *ctx, int op, const double *x, int64_t n, double *out) {
    RFB_ARG(ctx && n >= 0 && ((x && out) || n == 0), "rfb_unop_f64_dev");
    if (op < RFB_ROUND || op > RFB_CEIL) { rfb_set_error("unop: unknown op %d", op); return RFB_ERR_ARG; }
    if (n == 0) return RFB_OK;
    const bool vec_ok = aligned16(x) && aligned16(out);
    const int grid = rfb_grid_for(ctx, n, THREADS * 16, BLOCKS_PER_SM);
    UnopF64 f{op};
    k_map1<f64, f64, UnopF64><<<grid, THREADS, 0, ctx->stream>>>(x, out, n, vec_ok, f);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

// ------------------------------------------------------------------ mask logic: and / or / not  

namespace {
struct MaskAnd { __device__ __forceinline__ u8 operator()(u8 a, u8 b) const { return (u8)((a != 0) & (b != 0)); } };
struct MaskOr { __device__ __forceinline__ u8 operator()(u8 a, u8 b) const { return (u8)((a != 0) | (b != 0)); } };
struct MaskNot { __device__ __forceinline__ u8 operator()(u8 a) const { return (u8)(a == 0); } };
}  // namespace

// and_op_partial / or_op_partial (reference core/logic.c:34-86): mask[i] = mask[i] && next[i]; the right side may be one
// broadcast byte (bn == -1).  ray_not (core/order.c:422-443): out[i] = !x[i].
extern "C" int rfb_mask_logic_dev(rfb_ctx_t *ctx, int op, const uint8_t *a, int64_t n, const uint8_t *b, int64_t bn, uint8_t bs,
                                  uint8_t *out) {
    RFB_ARG(ctx && n >= 0 && ((a && out) || n == 0), "rfb_mask_logic_dev");
    if (n == 0) return RFB_OK;
    if (op == RFB_M_NOT) {
        const bool vec_ok = aligned16(a) && aligned16(out);
        k_map1<u8, u8, MaskNot><<<rfb_grid_for(ctx, n, THREADS * 128, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(a, out, n, vec_ok, MaskNot());
        RFB_CHECK_LAUNCH(ctx);
        return RFB_OK;
    }
    if (op != RFB_M_AND && op != RFB_M_OR) { rfb_set_error("mask logic: unknown op %d", op); return RFB_ERR_ARG; }
    if (bn >= 0 && bn != n) { rfb_set_error("mask logic: vector lengths differ"); return RFB_ERR_LENGTH; }
    RFB_ARG(bn < 0 || b, "rfb_mask_logic_dev: right operand");
    if (bn >= 0) {
        if (op == RFB_M_AND) return launch_map2<u8, u8, u8, false, false>(ctx, a, (u8)0, b, (u8)0, out, n, MaskAnd());
        return launch_map2<u8, u8, u8, false, false>(ctx, a, (u8)0, b, (u8)0, out, n, MaskOr());
    }
    if (op == RFB_M_AND) return launch_map2<u8, u8, u8, false, true>(ctx, a, (u8)0, nullptr, bs, out, n, MaskAnd());
    return launch_map2<u8, u8, u8, false, true>(ctx, a, (u8)0, nullptr, bs, out, n, MaskOr());
}
