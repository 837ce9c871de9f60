/*
 * rf_oracle.c — scalar CPU restatement of the RayforceDB columnar hot path.  TEST INFRASTRUCTURE ONLY
 * (see rf_oracle.h).  Compile: gcc -shared -fPIC -O2 -fsigned-char -fno-fast-math rf_oracle.c -lm
 *
 * Parity status: PINNED — checked against golden vectors transcribed from the reference's tests
 * (tests/golden/reference_kat.json) and differentially against the reference compiled from source
 * (oracle/_ref/librayforce_ref.so) by tests/test_oracle_golden.py and tests/test_oracle_vs_reference.py.
 *
 * All citations are paths inside the reference tree (commit 2151d51d).
 */
#include "rf_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef int16_t i16;
typedef int32_t i32;
typedef int64_t i64;
typedef uint64_t u64;
typedef uint8_t u8;
typedef double f64;

/* ------------------------------------------------------------------ scalar semantics (core/ops.h:63-197) */

/* core/ops.h:63-70: NaN by bit pattern, not by x != x (the reference builds with -funsafe-math-optimizations) */
static inline int isnan64(f64 x) {
    u64 u;
    memcpy(&u, &x, 8);
    return (u & 0x7FF0000000000000ULL) == 0x7FF0000000000000ULL && (u & 0x000FFFFFFFFFFFFFULL) != 0;
}
static inline f64 null_f64(void) {
    const u64 u = 0x7FF8000000000000ULL; /* a quiet NaN; tests compare NaNs as a class */
    f64 d;
    memcpy(&d, &u, 8);
    return d;
}

/* wrap-around integer arithmetic: the reference relies on plain C signed ops compiled -O3 */
static inline i64 wadd64(i64 a, i64 b) { return (i64)((u64)a + (u64)b); }
static inline i64 wsub64(i64 a, i64 b) { return (i64)((u64)a - (u64)b); }
static inline i64 wmul64(i64 a, i64 b) { return (i64)((u64)a * (u64)b); }
static inline i32 wadd32(i32 a, i32 b) { return (i32)((uint32_t)a + (uint32_t)b); }
static inline i32 wsub32(i32 a, i32 b) { return (i32)((uint32_t)a - (uint32_t)b); }
static inline i32 wmul32(i32 a, i32 b) { return (i32)((uint32_t)a * (uint32_t)b); }

/* core/ops.h:218-277 widening/narrowing keeps nullness */
static inline i64 i32_to_i64(i32 x) { return x == RFO_NULL_I32 ? RFO_NULL_I64 : (i64)x; }
static inline f64 i32_to_f64(i32 x) { return x == RFO_NULL_I32 ? null_f64() : (f64)x; }
static inline f64 i64_to_f64(i64 x) { return x == RFO_NULL_I64 ? null_f64() : (f64)x; }
static inline i32 i64_to_i32(i64 x) { return x == RFO_NULL_I64 ? RFO_NULL_I32 : (i32)x; }
static inline i64 i16_to_i64(i16 x) { return x == RFO_NULL_I16 ? RFO_NULL_I64 : (i64)x; }
/* f64 -> int: NaN -> null, otherwise C truncation.  Out-of-range casts are UB in C; the x86 cvttsd2si the
 * reference compiles to yields INT_MIN ("integer indefinite") which is also the null sentinel. */
static inline i64 f64_to_i64(f64 x) {
    if (isnan64(x)) return RFO_NULL_I64;
    if (!(x > -9223372036854775808.0 && x < 9223372036854775808.0)) return RFO_NULL_I64;
    return (i64)x;
}
static inline i32 f64_to_i32(f64 x) {
    if (isnan64(x)) return RFO_NULL_I32;
    if (!(x > -2147483649.0 && x < 2147483648.0)) return RFO_NULL_I32;
    return (i32)x;
}

/* core/ops.h:165-168: floor division / modulo built on C truncating ops */
static inline i64 eucl_div64(i64 x, i64 y) {
    if (y == -1) return (i64)(0 - (u64)x); /* avoids the INT_MIN / -1 trap; same value mod 2^64 */
    return (x / y) - ((((x < 0) != (y < 0)) && (x % y != 0)) ? 1 : 0);
}
static inline i64 eucl_mod64(i64 x, i64 y) { return wsub64(x, wmul64(eucl_div64(x, y), y)); }
static inline i32 eucl_div32(i32 x, i32 y) {
    if (y == -1) return (i32)(0 - (uint32_t)x);
    return (x / y) - ((((x < 0) != (y < 0)) && (x % y != 0)) ? 1 : 0);
}
static inline i32 eucl_mod32(i32 x, i32 y) { return wsub32(x, wmul32(eucl_div32(x, y), y)); }

int rfo_type_size(int type) {
    switch (type) {
        case RFO_B8: case RFO_U8: return 1;
        case RFO_I16: return 2;
        case RFO_I32: case RFO_DATE: case RFO_TIME: return 4;
        case RFO_I64: case RFO_SYMBOL: case RFO_TIMESTAMP: case RFO_F64: return 8;
        default: return 0;
    }
}

/* storage class of a type code */
enum { K_NONE = 0, K_U8, K_I16, K_I32, K_I64, K_F64 };
static int kind_of(int type) {
    switch (type) {
        case RFO_B8: case RFO_U8: return K_U8;
        case RFO_I16: return K_I16;
        case RFO_I32: case RFO_DATE: case RFO_TIME: return K_I32;
        case RFO_I64: case RFO_SYMBOL: case RFO_TIMESTAMP: return K_I64;
        case RFO_F64: return K_F64;
        default: return K_NONE;
    }
}

/* ------------------------------------------------------------------ predicate scan (core/cmp.c) */

/* core/ops.h:74-127.  Integers compare as plain values (a null is just the smallest value, Q4);
 * doubles order NaN below everything and NaN == NaN. */
static inline int cmp_i64(int op, i64 a, i64 b) {
    switch (op) {
        case RFO_EQ: return a == b; case RFO_NE: return a != b; case RFO_LT: return a < b;
        case RFO_GT: return a > b;  case RFO_LE: return a <= b; default: return a >= b;
    }
}
static inline int eq_f64(f64 a, f64 b) { return isnan64(a) ? isnan64(b) : isnan64(b) ? 0 : a == b; }
static inline int lt_f64(f64 a, f64 b) { return isnan64(a) ? !isnan64(b) : isnan64(b) ? 0 : a < b; }
static inline int gt_f64(f64 a, f64 b) { return isnan64(b) ? !isnan64(a) : isnan64(a) ? 0 : a > b; }
static inline int cmp_f64(int op, f64 a, f64 b) {
    switch (op) {
        case RFO_EQ: return eq_f64(a, b); case RFO_NE: return !eq_f64(a, b); case RFO_LT: return lt_f64(a, b);
        case RFO_GT: return gt_f64(a, b); case RFO_LE: return !gt_f64(a, b); default: return !lt_f64(a, b);
    }
}

/* load element i of a (vector or atom) operand, widened to i64 / f64 with null mapping */
static inline i64 ld_as_i64(int kind, const void *p, i64 i) {
    switch (kind) {
        case K_U8: return (i64)((const u8 *)p)[i];
        case K_I16: return i16_to_i64(((const i16 *)p)[i]);
        case K_I32: return i32_to_i64(((const i32 *)p)[i]);
        default: return ((const i64 *)p)[i];
    }
}
static inline f64 ld_as_f64(int kind, const void *p, i64 i) {
    switch (kind) {
        case K_U8: return (f64)((const u8 *)p)[i];
        case K_I16: { i16 v = ((const i16 *)p)[i]; return v == RFO_NULL_I16 ? null_f64() : (f64)v; }
        case K_I32: return i32_to_f64(((const i32 *)p)[i]);
        case K_I64: return i64_to_f64(((const i64 *)p)[i]);
        default: return ((const f64 *)p)[i];
    }
}

/* core/cmp.c:77-332: mixed-width integer operands are promoted (null-preserving) to the wider type, any F64 side
 * promotes both to F64.  Comparing same-width integers needs no conversion, and widening is monotone with the null
 * sentinel mapping to the wider null sentinel (still the minimum), so promoting everything to i64 is equivalent. */
int64_t rfo_cmp(int op, int xt, const void *x, int64_t xn, int yt, const void *y, int64_t yn, uint8_t *out) {
    int kx = kind_of(xt), ky = kind_of(yt);
    if (!kx || !ky || op < RFO_EQ || op > RFO_GE) return RFO_ERR_TYPE;
    /* the type matrix core/cmp.c:77-258: plain numerics I16/I32/I64/F64 mix freely; DATE/TIME/TIMESTAMP/SYMBOL only with
     * themselves; B8/U8 are atom-only (:78-85), so any vector form is a type error.  DATE<->TIMESTAMP (:243-257, unit
     * conversion) is not modelled. */
    {
        int px = (xt == RFO_I16 || xt == RFO_I32 || xt == RFO_I64 || xt == RFO_F64);
        int py = (yt == RFO_I16 || yt == RFO_I32 || yt == RFO_I64 || yt == RFO_F64);
        int same_special = (xt == yt) && (xt == RFO_DATE || xt == RFO_TIME || xt == RFO_TIMESTAMP || xt == RFO_SYMBOL);
        int atoms_u8 = (xn < 0 && yn < 0) && (kx == K_U8 && ky == K_U8);
        int date_ts = (xt == RFO_DATE && yt == RFO_TIMESTAMP) || (xt == RFO_TIMESTAMP && yt == RFO_DATE);
        if (!((px && py) || same_special || atoms_u8 || date_ts)) return RFO_ERR_TYPE;
    }
    /* DATE vs TIMESTAMP: the date side goes through date_to_timestamp (core/ops.h:264, core/cmp.c:243-257) */
    const i64 NANOS_FROM_DAY = 86400000000000LL;
    i64 sx = (xt == RFO_DATE && yt == RFO_TIMESTAMP) ? NANOS_FROM_DAY : 1, sy = (yt == RFO_DATE && xt == RFO_TIMESTAMP) ? NANOS_FROM_DAY : 1;
    if (xn >= 0 && yn >= 0 && xn != yn) return RFO_ERR_LENGTH; /* core/cmp.c:625-627 */
    i64 n = xn >= 0 ? xn : (yn >= 0 ? yn : 1);
    int use_f = (kx == K_F64 || ky == K_F64);
    for (i64 i = 0; i < n; i++) {
        i64 ix = xn >= 0 ? i : 0, iy = yn >= 0 ? i : 0;
        if (use_f) { out[i] = (u8)cmp_f64(op, ld_as_f64(kx, x, ix), ld_as_f64(ky, y, iy)); continue; }
        i64 a = ld_as_i64(kx, x, ix), b = ld_as_i64(ky, y, iy);
        if (a != RFO_NULL_I64) a = wmul64(a, sx);
        if (b != RFO_NULL_I64) b = wmul64(b, sy);
        out[i] = (u8)cmp_i64(op, a, b);
    }
    return n;
}

/* ------------------------------------------------------------------ selection vector (core/ops.c:255-273) */
int64_t rfo_where(const uint8_t *mask, int64_t n, int64_t *ids) {
    i64 j = 0;
    for (i64 i = 0; i < n; i++)
        if (mask[i]) ids[j++] = i;
    return j;
}

/* ------------------------------------------------------------------ gather (core/rayforce.c:1036-1098) */
int rfo_at_ids(int type, const void *col, const int64_t *ids, int64_t m, void *out) {
    switch (kind_of(type)) {
        case K_U8: for (i64 i = 0; i < m; i++) ((u8 *)out)[i] = ((const u8 *)col)[ids[i]]; return RFO_OK;
        case K_I16: for (i64 i = 0; i < m; i++) ((i16 *)out)[i] = ((const i16 *)col)[ids[i]]; return RFO_OK;
        case K_I32: for (i64 i = 0; i < m; i++) ((i32 *)out)[i] = ((const i32 *)col)[ids[i]]; return RFO_OK;
        case K_I64: for (i64 i = 0; i < m; i++) ((i64 *)out)[i] = ((const i64 *)col)[ids[i]]; return RFO_OK;
        case K_F64: for (i64 i = 0; i < m; i++) ((f64 *)out)[i] = ((const f64 *)col)[ids[i]]; return RFO_OK;
        default: return RFO_ERR_TYPE;
    }
}

/* ------------------------------------------------------------------ ungrouped folds (core/math.c:1785-2045) */

/* core/ops.h:183-188: null-skipping min / max */
static inline i64 min_i64(i64 a, i64 b) { return a == RFO_NULL_I64 ? b : b == RFO_NULL_I64 ? a : (a < b ? a : b); }
static inline i64 max_i64(i64 a, i64 b) { return a == RFO_NULL_I64 ? b : b == RFO_NULL_I64 ? a : (a > b ? a : b); }
static inline i32 min_i32(i32 a, i32 b) { return a == RFO_NULL_I32 ? b : b == RFO_NULL_I32 ? a : (a < b ? a : b); }
static inline i32 max_i32(i32 a, i32 b) { return a == RFO_NULL_I32 ? b : b == RFO_NULL_I32 ? a : (a > b ? a : b); }
static inline i16 min_i16(i16 a, i16 b) { return a == RFO_NULL_I16 ? b : b == RFO_NULL_I16 ? a : (a < b ? a : b); }
static inline i16 max_i16(i16 a, i16 b) { return a == RFO_NULL_I16 ? b : b == RFO_NULL_I16 ? a : (a > b ? a : b); }
static inline f64 min_f64(f64 a, f64 b) { return isnan64(a) ? b : isnan64(b) ? a : (a < b ? a : b); }
static inline f64 max_f64(f64 a, f64 b) { return isnan64(a) ? b : isnan64(b) ? a : (a > b ? a : b); }

static i64 count_nonnull(int type, const void *x, i64 n) { /* core/math.c:1785-1835, CNT* core/ops.h:148-152 */
    i64 c = 0;
    switch (kind_of(type)) {
        case K_U8: return n;
        case K_I16: for (i64 i = 0; i < n; i++) c += ((const i16 *)x)[i] != RFO_NULL_I16; return c;
        case K_I32: for (i64 i = 0; i < n; i++) c += ((const i32 *)x)[i] != RFO_NULL_I32; return c;
        case K_I64: for (i64 i = 0; i < n; i++) c += ((const i64 *)x)[i] != RFO_NULL_I64; return c;
        case K_F64: for (i64 i = 0; i < n; i++) c += !isnan64(((const f64 *)x)[i]); return c;
        default: return -1;
    }
}

int rfo_fold(int op, int type, const void *x, int64_t n, void *out, int *out_type) {
    int k = kind_of(type);
    memset(out, 0, 8);
    if (!k) return RFO_ERR_TYPE;
    switch (op) {
        case RFO_COUNT: /* (count x) = length, nulls included: core/misc.c:43-85 -> ops_count core/ops.c:169 */
            *(i64 *)out = n; *out_type = RFO_I64; return RFO_OK;
        case RFO_CNT: /* non-null count, B8 unsupported: core/math.c:1785-1835 */
            if (type == RFO_B8 || type == RFO_SYMBOL) return RFO_ERR_TYPE;
            *(i64 *)out = count_nonnull(type, x, n); *out_type = RFO_I64; return RFO_OK;
        case RFO_SUM: /* core/math.c:1837-1871: U8,I16 -> i64; I32/TIME stay 32-bit; DATE/TIMESTAMP/B8 -> type error */
            switch (type) {
                case RFO_U8: { i64 s = 0; for (i64 i = 0; i < n; i++) s += ((const u8 *)x)[i];
                               *(i64 *)out = s; *out_type = RFO_I64; return RFO_OK; }
                case RFO_I16: { i64 s = 0;
                               for (i64 i = 0; i < n; i++) { i64 v = i16_to_i64(((const i16 *)x)[i]);
                                                             if (v != RFO_NULL_I64) s = wadd64(s, v); }
                               *(i64 *)out = s; *out_type = RFO_I64; return RFO_OK; }
                case RFO_I32: case RFO_TIME: { i32 s = 0;
                               for (i64 i = 0; i < n; i++) { i32 v = ((const i32 *)x)[i];
                                                             if (v != RFO_NULL_I32) s = wadd32(s, v); }
                               *(i32 *)out = s; *out_type = type; return RFO_OK; }
                case RFO_I64: { i64 s = 0;
                               for (i64 i = 0; i < n; i++) { i64 v = ((const i64 *)x)[i];
                                                             if (v != RFO_NULL_I64) s = wadd64(s, v); }
                               *(i64 *)out = s; *out_type = RFO_I64; return RFO_OK; }
                case RFO_F64: { f64 s = 0.0;
                               for (i64 i = 0; i < n; i++) { f64 v = ((const f64 *)x)[i];
                                                             if (!isnan64(v)) s = s + v; }
                               *(f64 *)out = s; *out_type = RFO_F64; return RFO_OK; }
                default: return RFO_ERR_TYPE;
            }
        case RFO_MIN: case RFO_MAX: { /* core/math.c:1909-2045: accumulator starts as the typed null */
            int mn = (op == RFO_MIN);
            *out_type = type;
            if (type == RFO_B8 || type == RFO_SYMBOL) return RFO_ERR_TYPE;
            switch (k) {
                case K_U8: { const u8 *p = x; u8 a = n > 0 ? p[0] : 0;
                             for (i64 i = 1; i < n; i++) a = mn ? (a < p[i] ? a : p[i]) : (a > p[i] ? a : p[i]);
                             *(u8 *)out = a; return RFO_OK; }
                case K_I16: { const i16 *p = x; i16 a = RFO_NULL_I16;
                             for (i64 i = 0; i < n; i++) a = mn ? min_i16(a, p[i]) : max_i16(a, p[i]);
                             *(i16 *)out = a; return RFO_OK; }
                case K_I32: { const i32 *p = x; i32 a = RFO_NULL_I32;
                             for (i64 i = 0; i < n; i++) a = mn ? min_i32(a, p[i]) : max_i32(a, p[i]);
                             *(i32 *)out = a; return RFO_OK; }
                case K_I64: { const i64 *p = x; i64 a = RFO_NULL_I64;
                             for (i64 i = 0; i < n; i++) a = mn ? min_i64(a, p[i]) : max_i64(a, p[i]);
                             *(i64 *)out = a; return RFO_OK; }
                default: { const f64 *p = x; f64 a = null_f64();
                             for (i64 i = 0; i < n; i++) a = mn ? min_f64(a, p[i]) : max_f64(a, p[i]);
                             *(f64 *)out = a; return RFO_OK; }
            }
        }
        case RFO_AVG: { /* core/math.c:2445-2490: sum / non-null count through FDIVI64 / FDIVF64 */
            u8 tmp[8]; int tt; f64 r;
            *out_type = RFO_F64;
            if (!(type == RFO_U8 || type == RFO_I16 || type == RFO_I32 || type == RFO_I64 || type == RFO_F64))
                return RFO_ERR_TYPE;
            rfo_fold(RFO_SUM, type, x, n, tmp, &tt);
            i64 c = (type == RFO_U8) ? n : count_nonnull(type, x, n);
            if (type == RFO_F64) {
                f64 s; memcpy(&s, tmp, 8);
                f64 cf = i64_to_f64(c);
                r = (cf == 0.0 || isnan64(s) || isnan64(cf)) ? null_f64() : s / cf;
            } else {
                i64 s;
                if (type == RFO_I32) { i32 s32; memcpy(&s32, tmp, 4); s = i32_to_i64(s32); }
                else memcpy(&s, tmp, 8);
                r = (c == 0 || s == RFO_NULL_I64 || c == RFO_NULL_I64) ? null_f64() : (f64)s / (f64)c;
            }
            *(f64 *)out = r; return RFO_OK;
        }
        default: return RFO_ERR_TYPE;
    }
}

double rfo_sum_f64_exact(const double *x, int64_t n) {
    __float128 s = 0;
    for (i64 i = 0; i < n; i++)
        if (!isnan64(x[i])) s += (__float128)x[i];
    return (double)s;
}

/* ------------------------------------------------------------------ element-wise arithmetic (core/math.c) */

/* For op and operand types (I32/I64/F64) return the computation type `mt` and the result type `ot`, following the
 * per-case macro arguments in core/math.c:251-1782 and infer_*_type core/math.c:92-223.  0 = unsupported. */
static int binop_types(int op, int xt, int yt, int *mt, int *ot) {
    int okx = (xt == RFO_I32 || xt == RFO_I64 || xt == RFO_F64), oky = (yt == RFO_I32 || yt == RFO_I64 || yt == RFO_F64);
    if (!okx || !oky) return 0;
    int anyf = (xt == RFO_F64 || yt == RFO_F64), any64 = (xt == RFO_I64 || yt == RFO_I64);
    int wide = anyf ? RFO_F64 : any64 ? RFO_I64 : RFO_I32; /* usual promotion */
    switch (op) {
        case RFO_ADD: case RFO_SUB: case RFO_MUL: *mt = wide; *ot = wide; return 1;
        case RFO_XBAR: *mt = wide; *ot = wide; return 1; /* infer_xbar_type core/math.c:225-249, matrix :1637-1700 */
        case RFO_DIV: /* result keeps the LEFT operand's type (infer_div_type), computed in the promoted type */
            *mt = wide; *ot = xt; return 1;
        case RFO_FDIV: *mt = RFO_F64; *ot = RFO_F64; return 1;
        case RFO_MOD: /* integer % integer takes the RIGHT operand's type; anything with F64 is F64 */
            *mt = wide; *ot = anyf ? RFO_F64 : yt; return 1;
        default: return 0;
    }
}
int rfo_binop_type(int op, int xt, int yt) {
    int mt, ot;
    return binop_types(op, xt, yt, &mt, &ot) ? ot : RFO_ERR_TYPE;
}

/* core/ops.h:153-177 — null-propagating scalar ops in each computation type */
static inline i32 op_i32(int op, i32 x, i32 y) {
    if (x == RFO_NULL_I32 || y == RFO_NULL_I32) return RFO_NULL_I32;
    switch (op) {
        case RFO_ADD: return wadd32(x, y); case RFO_SUB: return wsub32(x, y); case RFO_MUL: return wmul32(x, y);
        case RFO_DIV: return y == 0 ? RFO_NULL_I32 : eucl_div32(x, y);
        case RFO_XBAR: { /* XBARI32 core/ops.h:193-194 */
            if (y == 0) return RFO_NULL_I32;
            i32 t = x < 0 ? wsub32(wadd32(x, 1), y) : x;
            return wmul32(y == -1 ? (i32)(0 - (uint32_t)t) : t / y, y);
        }
        default: return y == 0 ? RFO_NULL_I32 : eucl_mod32(x, y);
    }
}
static inline i64 op_i64(int op, i64 x, i64 y) {
    if (x == RFO_NULL_I64 || y == RFO_NULL_I64) return RFO_NULL_I64;
    switch (op) {
        case RFO_ADD: return wadd64(x, y); case RFO_SUB: return wsub64(x, y); case RFO_MUL: return wmul64(x, y);
        case RFO_DIV: return y == 0 ? RFO_NULL_I64 : eucl_div64(x, y);
        case RFO_XBAR: { /* XBARI64 core/ops.h:195-196 */
            if (y == 0) return RFO_NULL_I64;
            i64 t = x < 0 ? wsub64(wadd64(x, 1), y) : x;
            return wmul64(y == -1 ? (i64)(0 - (u64)t) : t / y, y);
        }
        default: return y == 0 ? RFO_NULL_I64 : eucl_mod64(x, y);
    }
}
static inline f64 op_f64(int op, f64 x, f64 y) {
    if (isnan64(x) || isnan64(y)) return null_f64();
    switch (op) {
        case RFO_ADD: return x + y; case RFO_SUB: return x - y; case RFO_MUL: return x * y;
        case RFO_DIV: return y == 0.0 ? null_f64() : floor(x / y);                  /* DIVF64 / FEUCL_DIV */
        case RFO_XBAR: { /* XBARF64 = FLOORF64(x / y) * y, core/ops.h:197,191 */
            if (y == 0.0) return null_f64(); /* the compiled reference floors with the hardware rounding instruction: inf * 0 = NaN */
            f64 q = x / y;
            if (isnan64(q)) return null_f64();
            f64 t = (f64)f64_to_i64(q);
            return ((q < 0.0 && t != q) ? t - 1.0 : t) * y;
        }
        default: return y == 0.0 ? null_f64() : x - floor(x / y) * y;               /* MODF64 / FEUCL_MOD */
    }
}
/* fdiv: core/math.c:1365-1447.  Left operand I32/I64 uses FDIVI64 *applied to already-converted doubles*
 * (core/ops.h:173): its null test compares against (double)INT64_MIN; left F64 uses FDIVF64 (core/ops.h:174). */
static inline f64 op_fdiv(int left_is_int, f64 x, f64 y) {
    if (left_is_int) {
        const f64 nul = -9223372036854775808.0;
        if (y == 0 || x == nul || y == nul) return null_f64();
        return x / y;
    }
    if (y == 0.0 || isnan64(x) || isnan64(y)) return null_f64();
    return x / y;
}

/* ---- the full type matrix (U8 / I16 / B8 / DATE / TIME / TIMESTAMP operands): every case of ray_add_partial .. ray_xbar_partial
 * (core/math.c:251-1782) is  out[i] = mt_to_ot(OP(lt_to_mt(x[i]), rt_to_mt(y[i])))  (__BINOP_V_V / _V_A / _A_V, core/math.c:55-90);
 * binop_matrix.inc lists (lt, rt, ot, mt, OP family) per (operator, form, operand types) and the type binop_map gives the result
 * vector.  The conversions are the <from>_to_<to> helpers of core/ops.h:218-277, the operators core/ops.h:125-197. */
typedef struct { int8_t op, form, xt, yt, lt, rt, ot, mt, fam, vt; int line; } binop_case_t;
static const binop_case_t BINOP_CASES[] = {
#include "binop_matrix.inc"
};
static const binop_case_t *binop_case(int op, int form, int xt, int yt) {
    for (size_t i = 0; i < sizeof(BINOP_CASES) / sizeof(BINOP_CASES[0]); i++) {
        const binop_case_t *c = &BINOP_CASES[i];
        if (c->op == op && c->form == form && c->xt == xt && c->yt == yt) return c;
    }
    return NULL;
}
int rfo_binop_form(int op, int form, int xt, int yt) {
    const binop_case_t *c = binop_case(op, form, xt, yt);
    return c ? c->vt : RFO_ERR_TYPE;
}
typedef struct { i64 i; f64 f; } bval_t;   /* .i for every integer kind (the exact value of its C type), .f for F64 */
static int null_width(int kind) {          /* 0: no null (B8 / U8); else the width of the integer whose minimum is the null */
    switch (kind) {
        case RFO_I16: return 16;
        case RFO_I32: case RFO_DATE: case RFO_TIME: return 32;
        case RFO_I64: case RFO_TIMESTAMP: return 64;
        default: return 0;
    }
}
static int is_null_int(int kind, i64 v) {
    switch (null_width(kind)) {
        case 16: return v == RFO_NULL_I16;
        case 32: return v == RFO_NULL_I32;
        case 64: return v == RFO_NULL_I64;
        default: return 0;
    }
}
static bval_t bload(int kind, const void *p, i64 i) {
    bval_t v = {0, 0.0};
    switch (kind_of(kind)) {
        case K_U8: v.i = ((const u8 *)p)[i]; break;
        case K_I16: v.i = ((const i16 *)p)[i]; break;
        case K_I32: v.i = ((const i32 *)p)[i]; break;
        case K_I64: v.i = ((const i64 *)p)[i]; break;
        default: v.f = ((const f64 *)p)[i]; break;
    }
    return v;
}
static void bstore(int kind, void *p, i64 i, bval_t v) {
    switch (kind_of(kind)) {
        case K_U8: ((u8 *)p)[i] = (u8)v.i; break;
        case K_I16: ((i16 *)p)[i] = (i16)v.i; break;
        case K_I32: ((i32 *)p)[i] = (i32)v.i; break;
        case K_I64: ((i64 *)p)[i] = v.i; break;
        default: ((f64 *)p)[i] = v.f; break;
    }
}
#define NANOS_PER_DAY 86400000000000LL   /* core/temporal.h:39-40 */
#define NANOS_PER_MILLI 1000000LL
static bval_t bconv(int from, int to, bval_t v) {   /* <from>_to_<to>, core/ops.h:218-277 */
    bval_t r = {0, 0.0};
    if (from == to) return v;
    if (from == RFO_F64) {
        if (null_width(to) == 64) r.i = f64_to_i64(v.f); else r.i = f64_to_i32(v.f);
        return r;
    }
    if (from == RFO_B8) v.i = (v.i != 0);                                   /* b8_to_i64 */
    if (to == RFO_F64) { r.f = is_null_int(from, v.i) ? null_f64() : (f64)v.i; return r; }
    if (to == RFO_B8) { r.i = (v.i != 0 && v.i != RFO_NULL_I64); return r; } /* i64_to_b8 */
    if (is_null_int(from, v.i)) {
        switch (null_width(to)) {
            case 16: r.i = RFO_NULL_I16; return r;
            case 32: r.i = RFO_NULL_I32; return r;
            case 64: r.i = RFO_NULL_I64; return r;
            default: r.i = (u8)v.i; return r;
        }
    }
    if (from == RFO_DATE && to == RFO_TIMESTAMP) { r.i = wmul64(NANOS_PER_DAY, v.i); return r; }
    if (from == RFO_TIME && to == RFO_TIMESTAMP) { r.i = wmul64(NANOS_PER_MILLI, v.i); return r; }
    if (from == RFO_TIMESTAMP && to == RFO_DATE) { r.i = (i32)(v.i / NANOS_PER_DAY); return r; }
    if (from == RFO_TIMESTAMP && to == RFO_TIME) { r.i = (i32)(v.i % NANOS_PER_DAY / NANOS_PER_MILLI); return r; }
    switch (null_width(to)) {
        case 16: r.i = (i16)v.i; break;
        case 32: r.i = (i32)v.i; break;
        case 64: r.i = v.i; break;
        default: r.i = (u8)v.i; break;
    }
    return r;
}
static inline u8 op_u8(int op, u8 x, u8 y) {     /* core/ops.h:125-130: no nulls, division by zero gives 0 */
    switch (op) {
        case RFO_ADD: return (u8)(x + y); case RFO_SUB: return (u8)(x - y); case RFO_MUL: return (u8)(x * y);
        case RFO_DIV: return y == 0 ? 0 : (u8)(x / y);
        default: return y == 0 ? 0 : (u8)(x % y);
    }
}
static inline i16 op_i16(int op, i16 x, i16 y) { /* core/ops.h:136-140: computed in int, narrowed */
    if (x == RFO_NULL_I16 || y == RFO_NULL_I16) return RFO_NULL_I16;
    switch (op) {
        case RFO_ADD: return (i16)(x + y); case RFO_SUB: return (i16)(x - y); case RFO_MUL: return (i16)(x * y);
        case RFO_DIV: return y == 0 ? RFO_NULL_I16 : (i16)eucl_div32(x, y);
        default: return y == 0 ? RFO_NULL_I16 : (i16)eucl_mod32(x, y);
    }
}
static int64_t binop_matrix(const binop_case_t *c, const void *x, i64 xn, const void *y, i64 yn, void *out, int *out_type) {
    const i64 n = xn >= 0 ? xn : yn;
    *out_type = c->vt;
    for (i64 i = 0; i < n; i++) {
        const bval_t a = bconv(c->lt, c->mt, bload(c->lt, x, xn >= 0 ? i : 0));
        const bval_t b = bconv(c->rt, c->mt, bload(c->rt, y, yn >= 0 ? i : 0));
        bval_t r = {0, 0.0};
        if (c->mt == RFO_F64) {
            if (c->op == RFO_FDIV) r.f = op_fdiv(c->fam == RFO_I64, a.f, b.f);   /* FDIVI64 on converted doubles vs FDIVF64 */
            else r.f = op_f64(c->op, a.f, b.f);
        } else {
            switch (c->fam) {
                case RFO_U8: r.i = op_u8(c->op, (u8)a.i, (u8)b.i); break;
                case RFO_I16: r.i = op_i16(c->op, (i16)a.i, (i16)b.i); break;
                case RFO_I32: r.i = op_i32(c->op, (i32)a.i, (i32)b.i); break;
                default: r.i = op_i64(c->op, a.i, b.i); break;
            }
        }
        bstore(c->ot, out, i, bconv(c->mt, c->ot, r));
    }
    return n;
}
static int plain_num(int t) { return t == RFO_I32 || t == RFO_I64 || t == RFO_F64; }

int64_t rfo_binop(int op, int xt, const void *x, int64_t xn, int yt, const void *y, int64_t yn, void *out,
                  int *out_type) {
    int mt, ot;
    if (!(plain_num(xt) && plain_num(yt))) {
        if (xn < 0 && yn < 0) return RFO_ERR_TYPE; /* atom op atom: no vector form */
        const binop_case_t *c = binop_case(op, xn >= 0 ? (yn >= 0 ? 0 : 1) : 2, xt, yt);
        if (!c) return RFO_ERR_TYPE;
        if (xn >= 0 && yn >= 0 && xn != yn) return RFO_ERR_LENGTH;
        return binop_matrix(c, x, xn, y, yn, out, out_type);
    }
    if (!binop_types(op, xt, yt, &mt, &ot)) return RFO_ERR_TYPE;
    if (xn >= 0 && yn >= 0 && xn != yn) return RFO_ERR_LENGTH; /* core/math.c:2287-2289 */
    i64 n = xn >= 0 ? xn : (yn >= 0 ? yn : 1);
    int kx = kind_of(xt), ky = kind_of(yt);
    *out_type = ot;
    for (i64 i = 0; i < n; i++) {
        i64 ix = xn >= 0 ? i : 0, iy = yn >= 0 ? i : 0;
        if (op == RFO_FDIV) {
            ((f64 *)out)[i] = op_fdiv(xt != RFO_F64, ld_as_f64(kx, x, ix), ld_as_f64(ky, y, iy));
        } else if (mt == RFO_F64) {
            f64 r = op_f64(op, ld_as_f64(kx, x, ix), ld_as_f64(ky, y, iy));
            if (ot == RFO_F64) ((f64 *)out)[i] = r;
            else if (ot == RFO_I64) ((i64 *)out)[i] = f64_to_i64(r);
            else ((i32 *)out)[i] = f64_to_i32(r);
        } else if (mt == RFO_I64) {
            i64 r = op_i64(op, ld_as_i64(kx, x, ix), ld_as_i64(ky, y, iy));
            if (ot == RFO_I64) ((i64 *)out)[i] = r;
            else ((i32 *)out)[i] = i64_to_i32(r);
        } else {
            ((i32 *)out)[i] = op_i32(op, ((const i32 *)x)[ix], ((const i32 *)y)[iy]);
        }
    }
    return n;
}

/* core/ops.h:190-192.  The reference computes (i64_t) casts inside doubles; results are doubles. */
int rfo_unop_f64(int op, const double *x, int64_t n, double *out) {
    for (i64 i = 0; i < n; i++) {
        f64 v = x[i], r;
        if (isnan64(v)) { out[i] = null_f64(); continue; }
        switch (op) {
            case RFO_ROUND: r = (f64)(v >= 0.0 ? (i64)(v + 0.5) : (i64)(v - 0.5)); break;
            case RFO_FLOOR: r = (v < 0.0 && (f64)(i64)v != v) ? (f64)(i64)v - 1.0 : (f64)(i64)v; break;
            case RFO_CEIL: { f64 w = -v; f64 fl = (w < 0.0 && (f64)(i64)w != w) ? (f64)(i64)w - 1.0 : (f64)(i64)w;
                             r = -fl; break; }
            default: return RFO_ERR_TYPE;
        }
        out[i] = r;
    }
    return RFO_OK;
}

/* ------------------------------------------------------------------ group index (core/index.c, core/hash.c) */

static u64 fnv1a64(i64 key) { /* core/hash.c:530-542 */
    u64 h = 14695981039346656037ull;
    for (int i = 0; i < 8; i++) { h ^= (u8)((u64)key >> (i * 8)); h *= 1099511628211ull; }
    return h;
}
static int is_prime(i64 v) {
    if (v < 2) return 0;
    if (v % 2 == 0) return v == 2;
    for (i64 d = 3; d * d <= v; d += 2) if (v % d == 0) return 0;
    return 1;
}

int rfo_group_i64(const int64_t *keys, const int64_t *filter, int64_t len, int64_t *group_ids, int64_t *first_ids,
                  int64_t *hk, rfo_group_info_t *info) {
    memset(info, 0, sizeof(*info));
    if (len == 0) { /* core/index.c:408-409: empty scope, range 0 <= len 0 -> dense path with zero groups */
        info->min = info->max = RFO_NULL_I64; info->range = 0; info->dense = 1; info->index_type = RFO_INDEX_SHIFT;
        return RFO_OK;
    }
    /* scope: core/index.c:376-435 */
    i64 mn, mx;
    mn = mx = filter ? keys[filter[0]] : keys[0];
    for (i64 i = 0; i < len; i++) { i64 v = filter ? keys[filter[i]] : keys[i]; if (v < mn) mn = v; if (v > mx) mx = v; }
    info->min = mn; info->max = mx;
    info->range = (i64)((u64)mx - (u64)mn + 1); /* wraps like the reference's i64 arithmetic (core/index.c:434) */
    i64 groups = 0;
    if (info->range <= len && info->range > 0) {
        /* perfect hash: core/index.c:2013-2055.  (A wrapped negative range is <= len in the reference and would index
         * out of bounds there; we treat it as sparse — unreachable for the test inputs.) */
        i64 range = info->range;
        i64 *slot = hk ? hk : (i64 *)malloc((size_t)range * 8);
        for (i64 i = 0; i < range; i++) slot[i] = RFO_NULL_I64;
        for (i64 i = 0; i < len; i++) {
            i64 s = (filter ? keys[filter[i]] : keys[i]) - mn;
            if (slot[s] == RFO_NULL_I64) { slot[s] = groups; first_ids[groups] = i; groups++; }
            group_ids[i] = slot[s];
        }
        if (!hk) free(slot);
        info->dense = 1;
        info->index_type = range <= RFO_INDEX_SCOPE_LIMIT ? RFO_INDEX_SHIFT : RFO_INDEX_IDS; /* core/index.c:2063 */
    } else {
        /* open addressing, linear probing, FNV-1a, prime capacity >= len/0.75, empty marker NULL_I64:
         * core/hash.c:35-56,129-148; single-chunk path of core/index.c:1800-1835 (first-occurrence numbering).
         * The reference keeps no first_ids on this path (meta = NULL_OBJ, core/index.c:1975); we record them anyway. */
        i64 cap = (i64)ceil((f64)len / 0.75);
        while (!is_prime(cap)) cap++;
        i64 *tk = (i64 *)malloc((size_t)cap * 8), *tv = (i64 *)malloc((size_t)cap * 8);
        for (i64 i = 0; i < cap; i++) tk[i] = RFO_NULL_I64;
        for (i64 i = 0; i < len; i++) {
            i64 k = filter ? keys[filter[i]] : keys[i];
            i64 s = (i64)(fnv1a64(k) % (u64)cap);
            while (tk[s] != RFO_NULL_I64 && tk[s] != k) s = (s + 1) % cap;
            if (tk[s] == RFO_NULL_I64) { tk[s] = k; tv[s] = groups; first_ids[groups] = i; groups++; }
            group_ids[i] = tv[s];
        }
        free(tk); free(tv);
        info->dense = 0;
        info->index_type = RFO_INDEX_IDS;
    }
    info->groups = groups;
    return RFO_OK;
}

/* multi-key grouping: core/index.c:2731-2793 index_group_list (perfect-hash key fusion :2308-2424, or row hashes + radix
 * partitions :2556-2729).  Whatever the internal path, groups are numbered by first occurrence of the key TUPLE in
 * (filtered) row order (pinned against the reference's `select ... by: {a: a b: b}` in tests/test_oracle_vs_reference.py).
 * Restated with one open-addressing table over tuple hashes. */
int rfo_group_multi(int ncols, const int64_t *const *cols, const int64_t *filter, int64_t len, int64_t *group_ids,
                    int64_t *first_ids, int64_t *groups_out) {
    if (ncols < 1) return RFO_ERR_TYPE;
    i64 cap = 16;
    while (cap < 2 * len + 1) cap <<= 1;
    i64 *slot = (i64 *)malloc((size_t)cap * 8);   /* group id or -1 */
    for (i64 i = 0; i < cap; i++) slot[i] = -1;
    i64 groups = 0;
    for (i64 i = 0; i < len; i++) {
        const i64 row = filter ? filter[i] : i;
        u64 h = 0x9E3779B97F4A7C15ULL;
        for (int c = 0; c < ncols; c++) h = (h ^ fnv1a64(cols[c][row])) * 0xBF58476D1CE4E5B9ULL;
        i64 s = (i64)(h & (u64)(cap - 1));
        for (;;) {
            const i64 g = slot[s];
            if (g < 0) { slot[s] = groups; first_ids[groups] = i; group_ids[i] = groups; groups++; break; }
            const i64 frow = filter ? filter[first_ids[g]] : first_ids[g];
            int same = 1;
            for (int c = 0; c < ncols && same; c++) same = cols[c][frow] == cols[c][row];
            if (same) { group_ids[i] = g; break; }
            s = (s + 1) & (cap - 1);
        }
    }
    free(slot);
    *groups_out = groups;
    return RFO_OK;
}

/* ------------------------------------------------------------------ grouped aggregates (core/aggr.c) */

int rfo_aggr_last(int val_type, const void *val, const int64_t *filter, const int64_t *gid, int64_t len, int64_t groups,
                  int64_t nchunks, void *out, int *out_type);
int rfo_aggr(int op, int val_type, const void *val, const int64_t *filter, const int64_t *gid, int64_t len,
             int64_t groups, void *out, int *out_type) {
    int k = kind_of(val_type);
    if (!k) return RFO_ERR_TYPE;
#define ROW(i) (filter ? filter[i] : (i))
    switch (op) {
        case RFO_COUNT: { /* core/aggr.c:1317-1378: rows per group, nulls included */
            i64 *o = out; *out_type = RFO_I64;
            if (k == K_U8 || k == K_I16) return RFO_ERR_TYPE;
            for (i64 g = 0; g < groups; g++) o[g] = 0;
            for (i64 i = 0; i < len; i++) o[gid[i]]++;
            return RFO_OK;
        }
        case RFO_SUM: /* core/aggr.c:1078-1150: STICKY null (ADD*, not FOLD_ADD*), accumulator in the value type.
                       * The driver aggr_sum (:1107-1150) accepts I16, I64 and F64 only (I32/DATE/TIME exist only as parted). */
            *out_type = val_type;
            if (val_type == RFO_I64) {
                i64 *o = out; const i64 *v = val;
                for (i64 g = 0; g < groups; g++) o[g] = 0;
                for (i64 i = 0; i < len; i++) { i64 a = o[gid[i]], b = v[ROW(i)];
                    o[gid[i]] = (a == RFO_NULL_I64 || b == RFO_NULL_I64) ? RFO_NULL_I64 : wadd64(a, b); }
                return RFO_OK;
            }
            if (val_type == RFO_I16) {
                i16 *o = out; const i16 *v = val;
                for (i64 g = 0; g < groups; g++) o[g] = 0;
                for (i64 i = 0; i < len; i++) { i16 a = o[gid[i]], b = v[ROW(i)];
                    o[gid[i]] = (a == RFO_NULL_I16 || b == RFO_NULL_I16) ? RFO_NULL_I16 : (i16)((uint16_t)a + (uint16_t)b); }
                return RFO_OK;
            }
            if (k == K_F64) {
                f64 *o = out; const f64 *v = val;
                for (i64 g = 0; g < groups; g++) o[g] = 0.0;
                for (i64 i = 0; i < len; i++) { f64 a = o[gid[i]], b = v[ROW(i)];
                    o[gid[i]] = (isnan64(a) || isnan64(b)) ? null_f64() : a + b; }
                return RFO_OK;
            }
            return RFO_ERR_TYPE;
        case RFO_MIN: case RFO_MAX: { /* core/aggr.c:1152-1315: min starts at +INF, max at NULL; I32 unsupported */
            int mn = (op == RFO_MIN);
            *out_type = val_type;
            if (val_type == RFO_I64 || val_type == RFO_TIMESTAMP) {
                i64 *o = out; const i64 *v = val;
                for (i64 g = 0; g < groups; g++) o[g] = mn ? RFO_INF_I64 : RFO_NULL_I64;
                for (i64 i = 0; i < len; i++) { i64 *a = &o[gid[i]]; *a = mn ? min_i64(*a, v[ROW(i)]) : max_i64(*a, v[ROW(i)]); }
                return RFO_OK;
            }
            if (val_type == RFO_I16) {
                i16 *o = out; const i16 *v = val;
                for (i64 g = 0; g < groups; g++) o[g] = mn ? (i16)0x7FFF : RFO_NULL_I16;
                for (i64 i = 0; i < len; i++) { i16 *a = &o[gid[i]]; *a = mn ? min_i16(*a, v[ROW(i)]) : max_i16(*a, v[ROW(i)]); }
                return RFO_OK;
            }
            if (val_type == RFO_DATE || val_type == RFO_TIME) {
                i32 *o = out; const i32 *v = val;
                for (i64 g = 0; g < groups; g++) o[g] = mn ? RFO_INF_I32 : RFO_NULL_I32;
                for (i64 i = 0; i < len; i++) { i32 *a = &o[gid[i]]; *a = mn ? min_i32(*a, v[ROW(i)]) : max_i32(*a, v[ROW(i)]); }
                return RFO_OK;
            }
            if (val_type == RFO_F64) {
                f64 *o = out; const f64 *v = val;
                for (i64 g = 0; g < groups; g++) o[g] = mn ? (f64)INFINITY : null_f64();
                for (i64 i = 0; i < len; i++) { f64 *a = &o[gid[i]]; *a = mn ? min_f64(*a, v[ROW(i)]) : max_f64(*a, v[ROW(i)]); }
                return RFO_OK;
            }
            return RFO_ERR_TYPE;
        }
        case RFO_AVG: { /* core/aggr.c:1455-1875, 2013-2060: f64 sum of non-null values / non-null count */
            f64 *o = out; *out_type = RFO_F64;
            if (!(val_type == RFO_I16 || val_type == RFO_I32 || val_type == RFO_DATE || val_type == RFO_TIME ||
                  val_type == RFO_I64 || val_type == RFO_F64)) return RFO_ERR_TYPE;
            i64 *c = (i64 *)calloc((size_t)(groups > 0 ? groups : 1), 8);
            for (i64 g = 0; g < groups; g++) o[g] = 0.0;
            for (i64 i = 0; i < len; i++) {
                i64 r = ROW(i), g = gid[i];
                if (k == K_I64) { i64 v = ((const i64 *)val)[r]; if (v != RFO_NULL_I64) { o[g] += (f64)v; c[g]++; } }
                else if (k == K_I32) { i32 v = ((const i32 *)val)[r]; if (v != RFO_NULL_I32) { o[g] += (f64)v; c[g]++; } }
                else if (k == K_I16) { i16 v = ((const i16 *)val)[r]; if (v != RFO_NULL_I16) { o[g] += (f64)v; c[g]++; } }
                else { f64 v = ((const f64 *)val)[r]; if (!isnan64(v)) { o[g] += v; c[g]++; } }
            }
            for (i64 g = 0; g < groups; g++) o[g] = c[g] == 0 ? null_f64() : o[g] / (f64)c[g];
            free(c);
            return RFO_OK;
        }
        case RFO_MED: { /* core/aggr.c:2136-2246: collect the group's values (aggr_collect), sort them like ray_asc (nulls /
                         * NaN first, they stay in), median as f64; only I64/TIMESTAMP/F64 have a case, the rest gives nulls */
            f64 *o = out; *out_type = RFO_F64;
            i64 *rows = (i64 *)malloc((size_t)(len > 0 ? len : 1) * 8), *offs = (i64 *)malloc((size_t)(groups + 1) * 8);
            rfo_group_rows(gid, filter, len, groups, rows, offs);
            for (i64 g = 0; g < groups; g++) {
                i64 l = offs[g + 1] - offs[g];
                if (l == 0 || !(val_type == RFO_I64 || val_type == RFO_TIMESTAMP || val_type == RFO_F64)) { o[g] = null_f64(); continue; }
                u64 *v = (u64 *)malloc((size_t)l * 8);
                i64 *perm = (i64 *)malloc((size_t)l * 8);
                for (i64 j = 0; j < l; j++) v[j] = ((const u64 *)val)[rows[offs[g] + j]];
                rfo_sort(val_type, v, l, 0, perm);
                i64 mid = l / 2;
                if (val_type == RFO_F64) {
                    const f64 *f = (const f64 *)v;
                    o[g] = (l % 2 == 0) ? (f[perm[mid - 1]] + f[perm[mid]]) / 2.0 : f[perm[mid]];
                } else {
                    const i64 *x = (const i64 *)v;
                    o[g] = (l % 2 == 0) ? ((f64)x[perm[mid - 1]] + (f64)x[perm[mid]]) / 2.0 : (f64)x[perm[mid]];
                }
                free(v); free(perm);
            }
            free(rows); free(offs);
            return RFO_OK;
        }
        case RFO_DEV: { /* core/aggr.c:2250-2330 accumulation (f64 sum, sum of squares, non-null count), :2893-2906 formula */
            f64 *o = out; *out_type = RFO_F64;
            if (!(val_type == RFO_I16 || val_type == RFO_I32 || val_type == RFO_DATE || val_type == RFO_TIME ||
                  val_type == RFO_I64 || val_type == RFO_TIMESTAMP || val_type == RFO_F64)) return RFO_ERR_TYPE;
            f64 *s = (f64 *)calloc((size_t)(groups > 0 ? groups : 1), 8), *q = (f64 *)calloc((size_t)(groups > 0 ? groups : 1), 8);
            i64 *c = (i64 *)calloc((size_t)(groups > 0 ? groups : 1), 8);
            for (i64 i = 0; i < len; i++) {
                i64 r = ROW(i), g = gid[i];
                f64 v; int ok;
                if (k == K_I64) { i64 x = ((const i64 *)val)[r]; ok = x != RFO_NULL_I64; v = (f64)x; }
                else if (k == K_I32) { i32 x = ((const i32 *)val)[r]; ok = x != RFO_NULL_I32; v = (f64)x; }
                else if (k == K_I16) { i16 x = ((const i16 *)val)[r]; ok = x != RFO_NULL_I16; v = (f64)x; }
                else { v = ((const f64 *)val)[r]; ok = !isnan64(v); }
                if (ok) { s[g] += v; q[g] += v * v; c[g]++; }
            }
            for (i64 g = 0; g < groups; g++) {
                if (c[g] == 0) o[g] = null_f64();
                else if (c[g] == 1) o[g] = 0.0;
                else { f64 mean = s[g] / (f64)c[g], var = q[g] / (f64)c[g] - mean * mean; o[g] = var < 0.0 ? 0.0 : sqrt(var); }
            }
            free(s); free(q); free(c);
            return RFO_OK;
        }
        case RFO_FIRST: { /* aggr_first, core/aggr.c:441-577: the index carries first_ids (every index built on this path does), so
                           * the result is the value AT the group's first row, null or not, in the value's own type */
            int w = rfo_type_size(val_type);
            *out_type = val_type;
            char *seen = (char *)calloc((size_t)(groups > 0 ? groups : 1), 1);
            for (i64 i = 0; i < len; i++) {
                i64 g = gid[i];
                if (!seen[g]) { seen[g] = 1; memcpy((char *)out + (size_t)g * w, (const char *)val + (size_t)ROW(i) * w, (size_t)w); }
            }
            free(seen);
            return RFO_OK;
        }
        case RFO_LAST: return rfo_aggr_last(val_type, val, filter, gid, len, groups, 1, out, out_type);
        default: return RFO_ERR_TYPE;
    }
#undef ROW
}

/* aggr_last, core/aggr.c:851-1075: aggr_last_partial (:851-893) keeps, per group, the last NON-NULL value of its chunk of rows
 * (null when the chunk has none); aggr_map (:262-295) cuts the rows into `nchunks` runs of len / nchunks rows (the last takes the
 * remainder) and AGGR_COLLECT merges the partials in chunk order with `if (out == null) out = in` — a group's answer is its last
 * non-null value inside the FIRST chunk that has one.  nchunks = pool_split_by_mem(len, groups, width) (core/pool.c:450-478):
 * 1 below 16384 rows, else the executor count capped by 64 MiB / (groups * width).  No U8/B8 case; the partials are vectors of
 * the value's own type, so is the result. */
int rfo_aggr_last(int val_type, const void *val, const int64_t *filter, const int64_t *gid, int64_t len, int64_t groups,
                  int64_t nchunks, void *out, int *out_type) {
    int k = kind_of(val_type);
    if (!k || k == K_U8) return RFO_ERR_TYPE;
    int w = rfo_type_size(val_type);
    *out_type = val_type;
    if (nchunks < 1) nchunks = 1;
    i64 chunk = len / nchunks;
    char *has = (char *)calloc((size_t)(groups > 0 ? groups : 1), 1), *tmp_has = (char *)malloc((size_t)(groups > 0 ? groups : 1));
    char *tmp = (char *)malloc((size_t)(groups > 0 ? groups : 1) * (size_t)w);
    for (i64 g = 0; g < groups; g++) {   /* typed nulls */
        if (k == K_I64) ((i64 *)out)[g] = RFO_NULL_I64;
        else if (k == K_I32) ((i32 *)out)[g] = RFO_NULL_I32;
        else if (k == K_I16) ((i16 *)out)[g] = RFO_NULL_I16;
        else ((f64 *)out)[g] = null_f64();
    }
    for (i64 c = 0; c < nchunks; c++) {
        i64 lo = c * chunk, hi = c == nchunks - 1 ? len : lo + chunk;
        memset(tmp_has, 0, (size_t)(groups > 0 ? groups : 1));
        for (i64 i = lo; i < hi; i++) {
            i64 r = filter ? filter[i] : i, g = gid[i];
            int nul;
            if (k == K_I64) nul = ((const i64 *)val)[r] == RFO_NULL_I64;
            else if (k == K_I32) nul = ((const i32 *)val)[r] == RFO_NULL_I32;
            else if (k == K_I16) nul = ((const i16 *)val)[r] == RFO_NULL_I16;
            else nul = isnan64(((const f64 *)val)[r]);
            if (!nul) { tmp_has[g] = 1; memcpy(tmp + (size_t)g * w, (const char *)val + (size_t)r * w, (size_t)w); }
        }
        for (i64 g = 0; g < groups; g++)
            if (!has[g] && tmp_has[g]) { has[g] = 1; memcpy((char *)out + (size_t)g * w, tmp + (size_t)g * w, (size_t)w); }
    }
    free(has); free(tmp_has); free(tmp);
    return RFO_OK;
}

/* Parted aggregates: PARTED_MAP (core/aggr.c:183-260) and aggr_avg's parted branch (:2065-2127) with no filter.  Every
 * partition is aggregated as ONE group by the grouped partial (rfo_aggr with all group ids 0: sticky-null sum, +INF-initialised
 * min, null-initialised max); combine != 0 (groups == 1) folds the partition values with ADD (sticky) / MIN / MAX (null-skipping)
 * or sums the avg totals, else out holds one value per partition.  out: I64 / I32 / I16 / F64 entries of *out_type. */
int rfo_parted_aggr(int op, int val_type, int nparts, const void *const *parts, const int64_t *lens, int combine, void *out, int *out_type) {
    int k = kind_of(val_type);
    if (!(k == K_I16 || k == K_I32 || k == K_I64 || k == K_F64)) return RFO_ERR_TYPE;
    if (!(op == RFO_SUM || op == RFO_MIN || op == RFO_MAX || op == RFO_AVG)) return RFO_ERR_TYPE;
    if (op == RFO_AVG && val_type == RFO_TIMESTAMP) return RFO_ERR_TYPE;
    int w = op == RFO_AVG ? 8 : rfo_type_size(val_type);
    *out_type = op == RFO_AVG || k == K_F64 ? RFO_F64 : (k == K_I64 ? RFO_I64 : (k == K_I32 ? RFO_I32 : RFO_I16));
    f64 tsum = 0.0; i64 tcnt = 0;
    for (int p = 0; p < nparts; p++) {
        i64 n = lens[p];
        i64 *zeros = (i64 *)calloc((size_t)(n > 0 ? n : 1), 8);
        char v[8] = {0};
        int t;
        if (op == RFO_AVG) {       /* (f64 sum of non-nulls, count): aggr_avg_partial */
            f64 s = 0.0; i64 c = 0;
            for (i64 i = 0; i < n; i++) {
                if (k == K_I64) { i64 x = ((const i64 *)parts[p])[i]; if (x != RFO_NULL_I64) { s += (f64)x; c++; } }
                else if (k == K_I32) { i32 x = ((const i32 *)parts[p])[i]; if (x != RFO_NULL_I32) { s += (f64)x; c++; } }
                else if (k == K_I16) { i16 x = ((const i16 *)parts[p])[i]; if (x != RFO_NULL_I16) { s += (f64)x; c++; } }
                else { f64 x = ((const f64 *)parts[p])[i]; if (!isnan64(x)) { s += x; c++; } }
            }
            if (combine) { tsum += s; tcnt += c; }
            else { f64 a = c == 0 ? null_f64() : s / (f64)c; memcpy((char *)out + (size_t)p * 8, &a, 8); }
            free(zeros);
            continue;
        }
        /* the partial's value types: DATE/TIME partitions aggregate as their I32 payload, TIMESTAMP as I64 */
        if ((op == RFO_SUM && val_type == RFO_TIMESTAMP) || ((op == RFO_MIN || op == RFO_MAX) && val_type == RFO_I32)) { free(zeros); return RFO_ERR_TYPE; }
        int pt = val_type;
        if (op == RFO_SUM && k == K_I32) {   /* rfo_aggr has no I32 sum driver (the non-parted aggr_sum has none): i32 sticky wrap-around sum */
            i32 s = 0; int nul = 0;
            for (i64 i = 0; i < n; i++) { i32 x = ((const i32 *)parts[p])[i]; if (x == RFO_NULL_I32) nul = 1; else s = (i32)((uint32_t)s + (uint32_t)x); }
            if (nul) s = RFO_NULL_I32;
            memcpy(v, &s, 4);
        } else if (rfo_aggr(op, pt, parts[p], NULL, zeros, n, 1, v, &t) < 0) { free(zeros); return RFO_ERR_TYPE; }
        free(zeros);
        if (!combine) { memcpy((char *)out + (size_t)p * w, v, (size_t)w); continue; }
        if (p == 0) { memcpy(out, v, (size_t)w); continue; }
        if (k == K_F64) {
            f64 a, b; memcpy(&a, out, 8); memcpy(&b, v, 8);
            a = op == RFO_SUM ? ((isnan64(a) || isnan64(b)) ? null_f64() : a + b) : (op == RFO_MIN ? min_f64(a, b) : max_f64(a, b));
            memcpy(out, &a, 8);
        } else if (k == K_I64) {
            i64 a, b; memcpy(&a, out, 8); memcpy(&b, v, 8);
            a = op == RFO_SUM ? ((a == RFO_NULL_I64 || b == RFO_NULL_I64) ? RFO_NULL_I64 : wadd64(a, b)) : (op == RFO_MIN ? min_i64(a, b) : max_i64(a, b));
            memcpy(out, &a, 8);
        } else if (k == K_I32) {
            i32 a, b; memcpy(&a, out, 4); memcpy(&b, v, 4);
            a = op == RFO_SUM ? ((a == RFO_NULL_I32 || b == RFO_NULL_I32) ? RFO_NULL_I32 : (i32)((uint32_t)a + (uint32_t)b)) : (op == RFO_MIN ? min_i32(a, b) : max_i32(a, b));
            memcpy(out, &a, 4);
        } else {
            i16 a, b; memcpy(&a, out, 2); memcpy(&b, v, 2);
            a = op == RFO_SUM ? ((a == RFO_NULL_I16 || b == RFO_NULL_I16) ? RFO_NULL_I16 : (i16)((uint16_t)a + (uint16_t)b)) : (op == RFO_MIN ? min_i16(a, b) : max_i16(a, b));
            memcpy(out, &a, 2);
        }
    }
    if (op == RFO_AVG && combine) { f64 a = tcnt == 0 ? null_f64() : tsum / (f64)tcnt; memcpy(out, &a, 8); }
    return RFO_OK;
}

/* aggr_row / aggr_collect (core/aggr.c:3021-3136): AGGR_ITER pushes row x (= filter[i] or i) onto list gid[i], i ascending */
int rfo_group_rows(const int64_t *gid, const int64_t *filter, int64_t len, int64_t groups, int64_t *rows, int64_t *offsets) {
    for (i64 g = 0; g <= groups; g++) offsets[g] = 0;
    for (i64 i = 0; i < len; i++) offsets[gid[i] + 1]++;
    for (i64 g = 0; g < groups; g++) offsets[g + 1] += offsets[g];
    i64 *cur = (i64 *)malloc((size_t)(groups > 0 ? groups : 1) * 8);
    for (i64 g = 0; g < groups; g++) cur[g] = offsets[g];
    for (i64 i = 0; i < len; i++) rows[cur[gid[i]]++] = filter ? filter[i] : i;
    free(cur);
    return RFO_OK;
}

/* ray_med (core/math.c:2529-2626): vector cases exist for U8, I16, I64 only.  l = ray_cnt = NON-NULL count, but the sorted
 * column keeps its nulls (first): the reference indexes it with l anyway (its own null goldens are commented out,
 * tests/lang.c:2582-2585).  The two middle elements are added as integers before the division. */
int rfo_med(int type, const void *x, int64_t n, double *out) {
    if (!(type == RFO_U8 || type == RFO_I16 || type == RFO_I64)) return RFO_ERR_TYPE;
    i64 l = 0;
    for (i64 i = 0; i < n; i++)
        l += type == RFO_U8 ? 1 : (type == RFO_I16 ? ((const i16 *)x)[i] != RFO_NULL_I16 : ((const i64 *)x)[i] != RFO_NULL_I64);
    if (l == 0) { *out = null_f64(); return RFO_OK; }
    i64 *perm = (i64 *)malloc((size_t)n * 8);
    rfo_sort(type, x, n, 0, perm);
#define AT(j) (type == RFO_U8 ? (i64)((const u8 *)x)[perm[j]] : type == RFO_I16 ? (i64)((const i16 *)x)[perm[j]] : ((const i64 *)x)[perm[j]])
    *out = (l % 2 == 0) ? (f64)wadd64(AT(l / 2 - 1), AT(l / 2)) / 2.0 : (f64)AT(l / 2);
#undef AT
    free(perm);
    return RFO_OK;
}

/* ray_dev (core/math.c:2628-2700): mean = ray_sum / ray_cnt (I32/TIME sums wrap in 32 bits), then
 * sqrt(sum((x - mean)^2 over non-null) / cnt) (ray_sq_sub_partial :2119-2174).  Types whose ray_sum is a type error
 * (DATE, TIMESTAMP) are a type error here as well (the reference dereferences the error object there). */
int rfo_dev(int type, const void *x, int64_t n, double *out) {
    int k = kind_of(type);
    i64 sum[1]; int st;
    if (!k || type == RFO_B8 || type == RFO_SYMBOL) return RFO_ERR_TYPE;
    if (rfo_fold(RFO_SUM, type, x, n, sum, &st) < 0) return RFO_ERR_TYPE;
    i64 l = 0;
    for (i64 i = 0; i < n; i++) {
        switch (k) {
            case K_U8: l++; break;
            case K_I16: l += ((const i16 *)x)[i] != RFO_NULL_I16; break;
            case K_I32: l += ((const i32 *)x)[i] != RFO_NULL_I32; break;
            case K_I64: l += ((const i64 *)x)[i] != RFO_NULL_I64; break;
            default: l += !isnan64(((const f64 *)x)[i]); break;
        }
    }
    if (l == 0) { *out = null_f64(); return RFO_OK; }
    if (l == 1) { *out = 0.0; return RFO_OK; }
    f64 mean;
    if (k == K_F64) { f64 fs; memcpy(&fs, sum, 8); mean = fs / (f64)l; }
    else if (k == K_I32) { i32 s32; memcpy(&s32, sum, 4); mean = (f64)s32 / (f64)l; }
    else mean = (f64)sum[0] / (f64)l;
    f64 acc = 0.0;
    for (i64 i = 0; i < n; i++) {
        f64 v; int ok = 1;
        switch (k) {
            case K_U8: v = (f64)((const u8 *)x)[i]; break;
            case K_I16: ok = ((const i16 *)x)[i] != RFO_NULL_I16; v = (f64)((const i16 *)x)[i]; break;
            case K_I32: ok = ((const i32 *)x)[i] != RFO_NULL_I32; v = (f64)((const i32 *)x)[i]; break;
            case K_I64: ok = ((const i64 *)x)[i] != RFO_NULL_I64; v = (f64)((const i64 *)x)[i]; break;
            default: v = ((const f64 *)x)[i]; ok = !isnan64(v); break;
        }
        if (ok) { f64 t = v - mean; acc += t * t; }
    }
    *out = sqrt(acc / (f64)l);
    return RFO_OK;
}

/* ------------------------------------------------------------------ equi-join row matching (core/index.c) */

/* ray_find -> index_find_i64 (core/index.c:1507-1574) and index_left_join_obj (:2886-2928): both keep, per distinct key
 * (tuple), the first build row (sequential insert, `if slot empty: slot = i`) and answer every probe row with it or NULL.
 * Restated as: sort the build rows by (tuple, row), binary-search every probe tuple. */
typedef struct { int ncols; const int64_t *const *cols; } rfo_tuple_ctx;
static int tuple_cmp(const rfo_tuple_ctx *a, i64 ra, const rfo_tuple_ctx *b, i64 rb) {
    for (int c = 0; c < a->ncols; c++) {
        i64 x = a->cols[c][ra], y = b->cols[c][rb];
        if (x != y) return x < y ? -1 : 1;
    }
    return 0;
}
int rfo_find_rows(int ncols, const int64_t *const *build, int64_t build_len, const int64_t *const *probe, int64_t probe_len, int64_t *ids) {
    rfo_tuple_ctx B = {ncols, build}, P = {ncols, probe};
    i64 *ord = (i64 *)malloc((size_t)(build_len > 0 ? build_len : 1) * 8), *tmp = (i64 *)malloc((size_t)(build_len > 0 ? build_len : 1) * 8);
    for (i64 i = 0; i < build_len; i++) ord[i] = i;
    i64 *src = ord, *dst = tmp;                       /* stable merge sort by tuple: equal tuples stay in row order */
    for (i64 w = 1; w < build_len; w *= 2) {
        for (i64 lo = 0; lo < build_len; lo += 2 * w) {
            i64 mid = lo + w < build_len ? lo + w : build_len, hi = lo + 2 * w < build_len ? lo + 2 * w : build_len;
            i64 a = lo, b = mid, o = lo;
            while (a < mid && b < hi) dst[o++] = tuple_cmp(&B, src[b], &B, src[a]) < 0 ? src[b++] : src[a++];
            while (a < mid) dst[o++] = src[a++];
            while (b < hi) dst[o++] = src[b++];
        }
        i64 *t = src; src = dst; dst = t;
    }
    for (i64 i = 0; i < probe_len; i++) {
        i64 lo = 0, hi = build_len;
        while (lo < hi) { i64 mid = (lo + hi) / 2; if (tuple_cmp(&B, src[mid], &P, i) < 0) lo = mid + 1; else hi = mid; }
        ids[i] = (lo < build_len && tuple_cmp(&B, src[lo], &P, i) == 0) ? src[lo] : RFO_NULL_I64;
    }
    free(ord); free(tmp);
    return RFO_OK;
}

/* index_inner_join_obj (core/index.c:2930-3000): the matched probe rows in ascending order with their first build row */
int64_t rfo_inner_join(int ncols, const int64_t *const *build, int64_t build_len, const int64_t *const *probe, int64_t probe_len,
                       int64_t *probe_ids, int64_t *build_ids) {
    i64 *ids = (i64 *)malloc((size_t)(probe_len > 0 ? probe_len : 1) * 8), j = 0;
    rfo_find_rows(ncols, build, build_len, probe, probe_len, ids);
    for (i64 i = 0; i < probe_len; i++)
        if (ids[i] != RFO_NULL_I64) { probe_ids[j] = i; build_ids[j] = ids[i]; j++; }
    free(ids);
    return j;
}

/* index_asof_join_obj (core/index.c:3194-3268): per key tuple the build rows in row order (push_raw :3219-3223); per probe row
 * index_bin_i64 / index_bin_i32 (:3103-3138) over its key's list: the last entry whose time is <= the probe time */
int rfo_asof_join(int ncols, const int64_t *const *build, int time_type, const void *build_time, int64_t build_len,
                  const int64_t *const *probe, const void *probe_time, int64_t probe_len, int64_t *ids) {
    int tk = kind_of(time_type);
    if (!(tk == K_I64 || tk == K_I32)) return RFO_ERR_TYPE;
    i64 nb = build_len > 0 ? build_len : 1;
    i64 *gid = (i64 *)malloc((size_t)nb * 8), *firsts = (i64 *)malloc((size_t)nb * 8), *rows = (i64 *)malloc((size_t)nb * 8);
    i64 *offs = (i64 *)malloc((size_t)(nb + 1) * 8), *first = (i64 *)malloc((size_t)(probe_len > 0 ? probe_len : 1) * 8), groups = 0;
    rfo_group_multi(ncols, build, NULL, build_len, gid, firsts, &groups);
    rfo_group_rows(gid, NULL, build_len, groups, rows, offs);
    rfo_find_rows(ncols, build, build_len, probe, probe_len, first);
    for (i64 i = 0; i < probe_len; i++) {
        ids[i] = RFO_NULL_I64;
        if (first[i] == RFO_NULL_I64) continue;
        i64 g = gid[first[i]], base = offs[g], len = offs[g + 1] - base, left = 0, right = len - 1, idx = -1;
        while (left <= right) {
            i64 mid = left + (right - left) / 2;
            int le = tk == K_I64 ? ((const i64 *)build_time)[rows[base + mid]] <= ((const i64 *)probe_time)[i]
                                 : ((const i32 *)build_time)[rows[base + mid]] <= ((const i32 *)probe_time)[i];
            if (le) { idx = mid; left = mid + 1; } else right = mid - 1;
        }
        if (idx >= 0) ids[i] = rows[base + idx];
    }
    free(gid); free(firsts); free(rows); free(offs); free(first);
    return RFO_OK;
}

/* ray_distinct -> index_distinct_i64 (core/index.c:551-607), direct-addressing branch only (range <= len or <= MAX_RANGE = 2^20):
 * mark the slots that occur, emit slot + min in ascending order.  Returns the count, or -1 when the range is not dense (the
 * hash branch's result order is the slot order of the reference's own table: not restated). */
int64_t rfo_distinct_i64(const int64_t *keys, int64_t n, int64_t *out) {
    if (n == 0) return 0;
    i64 mn = keys[0], mx = keys[0];
    for (i64 i = 1; i < n; i++) { if (keys[i] < mn) mn = keys[i]; if (keys[i] > mx) mx = keys[i]; }
    i64 range = (i64)((u64)mx - (u64)mn + 1);
    if (range <= 0 || !(range <= n || range <= (1 << 20))) {
        /* hash branch (core/index.c:579-603): an open-addressing table of next_prime(ceil(n / 0.75)) slots (core/hash.c:35-55,
         * ops_next_prime core/ops.c:66-88), slot = key % size with linear probing (ht_oa_tab_next, core/hash.c:129-148), rows
         * inserted in row order, nulls skipped; the result is the table read in SLOT order.  The start slot is a signed C
         * remainder: negative keys index before the table in the reference (Q17) — not restated: -2. */
        i64 size = (i64)ceil((f64)n / 0.75);
        for (;; size++) {
            int prime = size > 1;
            if (size > 3 && (size % 2 == 0 || size % 3 == 0)) prime = 0;
            for (i64 d = 5; prime && d * d <= size; d += 6) if (size % d == 0 || size % (d + 2) == 0) prime = 0;
            if (prime) break;
        }
        i64 *tab = (i64 *)malloc((size_t)size * 8);
        for (i64 s = 0; s < size; s++) tab[s] = RFO_NULL_I64;
        for (i64 i = 0; i < n; i++) {
            i64 k = keys[i];
            if (k == RFO_NULL_I64) continue;
            if (k < 0) { free(tab); return -2; }
            i64 s = k % size;
            while (tab[s] != RFO_NULL_I64 && tab[s] != k) s = (s + 1) % size;
            tab[s] = k;
        }
        i64 j = 0;
        for (i64 s = 0; s < size; s++) if (tab[s] != RFO_NULL_I64) out[j++] = tab[s];
        free(tab);
        return j;
    }
    u8 *mark = (u8 *)calloc((size_t)range, 1);
    for (i64 i = 0; i < n; i++) mark[keys[i] - mn] = 1;
    i64 j = 0;
    for (i64 s = 0; s < range; s++) if (mark[s]) out[j++] = s + mn;
    free(mark);
    return j;
}

/* Window join (reference core/join.c:358-485 -> index_window_join_obj core/index.c:3287-3346 -> the INDEX_TYPE_WINDOW branch of
 * AGGR_ITER core/aggr.c:131-160).  The right table is ordered by (key tuple, time) — ray_window_join sorts it with xasc first —
 * so the rows of one key form one block [first, last]; the index keeps exactly that pair per key (first = the row that created
 * the slot, last = the latest row seen).  For left row i with window [wlo[i], whi[i]] on a 4-byte time column:
 *   li = jtype 0 (window-join):  indexr_bin: the last row of the block with time <= wlo, the block's first row if there is none
 *        jtype 1 (window-join1): indexl_bin: the first row of the block with time >= wlo (the block's first row if none)
 *   ri = indexr_bin: the last row of the block with time <= whi (the block's first row if none)
 *   no block, time[li] > whi, or (jtype 1 and time[ri] < wlo)  ->  the aggregate's Null value; else fold val[li..ri] with the
 *   GROUPED partial of the aggregate (sticky-null sum, +INF-initialised min, null-initialised max, row count).
 * rfo_window_bounds computes (first, last) per left row; rfo_window_aggr folds.  NOT YET ON THE DEVICE. */
int rfo_window_bounds(int ncols, const int64_t *const *right, int64_t rl, const int64_t *const *left, int64_t ll, int64_t *first, int64_t *last) {
    i64 *f = (i64 *)malloc((size_t)(ll > 0 ? ll : 1) * 8);
    rfo_find_rows(ncols, right, rl, left, ll, f);
    /* last row of the key = the last right row whose tuple equals the first row's tuple: scan backwards once per distinct first */
    i64 *lastof = (i64 *)malloc((size_t)(rl > 0 ? rl : 1) * 8);
    for (i64 r = 0; r < rl; r++) lastof[r] = -1;
    rfo_tuple_ctx R = {ncols, right};
    i64 *fr = (i64 *)malloc((size_t)(rl > 0 ? rl : 1) * 8);
    rfo_find_rows(ncols, right, rl, right, rl, fr);            /* every right row -> first row of its key */
    for (i64 r = 0; r < rl; r++) lastof[fr[r]] = r;            /* ascending r: the last assignment wins */
    (void)R;
    for (i64 i = 0; i < ll; i++) { first[i] = f[i]; last[i] = f[i] == RFO_NULL_I64 ? RFO_NULL_I64 : lastof[f[i]]; }
    free(f); free(lastof); free(fr);
    return RFO_OK;
}
static i64 bin_r_i32(i32 val, const i32 *vals, i64 offset, i64 len) {   /* core/aggr.c:39-54 */
    i64 left = 0, right = len - 1, idx = 0;
    vals += offset;
    while (left <= right) { i64 mid = left + (right - left) / 2; if (vals[mid] <= val) { idx = mid; left = mid + 1; } else right = mid - 1; }
    return idx + offset;
}
static i64 bin_l_i32(i32 val, const i32 *vals, i64 offset, i64 len) {   /* core/aggr.c:56-72 */
    i64 left = 0, right = len - 1, idx = 0;
    vals += offset;
    while (left <= right) { i64 mid = left + (right - left) / 2; if (vals[mid] < val) left = mid + 1; else { idx = mid; right = mid - 1; } }
    return idx + offset;
}
int rfo_window_aggr(int op, int val_type, const void *val, const int32_t *rtime, int64_t ll, const int64_t *first, const int64_t *last,
                    const int32_t *wlo, const int32_t *whi, int jtype, void *out, int *out_type) {
    int k = kind_of(val_type);
    if (!(k == K_I64 || k == K_F64)) return RFO_ERR_TYPE;
    if (!(op == RFO_SUM || op == RFO_MIN || op == RFO_MAX || op == RFO_COUNT || op == RFO_AVG)) return RFO_ERR_TYPE;
    *out_type = op == RFO_COUNT ? RFO_I64 : (op == RFO_AVG ? RFO_F64 : val_type);
    for (i64 i = 0; i < ll; i++) {
        i64 li = 0, ri = 0;
        int none = first[i] == RFO_NULL_I64;
        if (!none) {
            i64 fi = first[i], n = last[i] - first[i] + 1;
            li = jtype == 0 ? bin_r_i32(wlo[i], rtime, fi, n) : bin_l_i32(wlo[i], rtime, fi, n);
            ri = bin_r_i32(whi[i], rtime, fi, n);
            if (rtime[li] > whi[i] || (jtype == 1 && rtime[ri] < wlo[i])) none = 1;
        }
        if (op == RFO_COUNT) { ((i64 *)out)[i] = none ? 0 : ri - li + 1 > 0 ? ri - li + 1 : 0; continue; }
        if (op == RFO_AVG) {   /* aggr_avg_partial's WINDOW branch (core/aggr.c:1545-1575) + the final division (:2060) */
            f64 so = 0.0; i64 co = 0;
            if (!none)
                for (i64 x = li; x <= ri; x++) {
                    if (k == K_I64) { i64 v = ((const i64 *)val)[x]; if (v != RFO_NULL_I64) { so += (f64)v; co++; } }
                    else { f64 v = ((const f64 *)val)[x]; if (!isnan64(v)) { so += v; co++; } }
                }
            ((f64 *)out)[i] = co == 0 ? null_f64() : so / (f64)co;
            continue;
        }
        if (k == K_I64) {
            const i64 *v = (const i64 *)val; i64 a;
            if (none) a = RFO_NULL_I64;
            else {
                a = op == RFO_SUM ? 0 : (op == RFO_MIN ? RFO_INF_I64 : RFO_NULL_I64);
                for (i64 x = li; x <= ri; x++)
                    a = op == RFO_SUM ? ((a == RFO_NULL_I64 || v[x] == RFO_NULL_I64) ? RFO_NULL_I64 : wadd64(a, v[x])) : (op == RFO_MIN ? min_i64(a, v[x]) : max_i64(a, v[x]));
            }
            ((i64 *)out)[i] = a;
        } else {
            const f64 *v = (const f64 *)val; f64 a;
            if (none) a = null_f64();
            else {
                a = op == RFO_SUM ? 0.0 : (op == RFO_MIN ? (f64)INFINITY : null_f64());
                for (i64 x = li; x <= ri; x++)
                    a = op == RFO_SUM ? ((isnan64(a) || isnan64(v[x])) ? null_f64() : a + v[x]) : (op == RFO_MIN ? min_f64(a, v[x]) : max_f64(a, v[x]));
            }
            ((f64 *)out)[i] = a;
        }
    }
    return RFO_OK;
}

/* ------------------------------------------------------------------ key sort (core/sort.c) */

/* order-preserving map to u64: integers flip the sign bit (core/sort.c:313), doubles core/sort.c:266-285 */
static inline u64 sortable(int kind, const void *x, i64 i) {
    switch (kind) {
        case K_U8: return ((const u8 *)x)[i];
        case K_I16: return (u64)(uint16_t)(((const i16 *)x)[i] ^ (i16)0x8000);
        case K_I32: return (u64)((uint32_t)((const i32 *)x)[i] ^ 0x80000000u);
        case K_I64: return (u64)((const i64 *)x)[i] ^ 0x8000000000000000ULL;
        default: {
            f64 v = ((const f64 *)x)[i];
            u64 u;
            if (isnan64(v)) return 0;
            memcpy(&u, &v, 8);
            return (u & 0x8000000000000000ULL) ? ~u : (u | 0x8000000000000000ULL);
        }
    }
}

/* The reference runs 16-bit-digit LSD counting passes; any stable sort by the same key yields the identical
 * permutation, so the restatement uses a stable bottom-up merge sort on (key, original index). */
int rfo_sort(int type, const void *x, int64_t n, int descending, int64_t *perm) {
    int k = kind_of(type);
    if (!k || type == RFO_SYMBOL) return RFO_ERR_TYPE;
    if (n == 0) return RFO_OK;
    u64 *key = (u64 *)malloc((size_t)n * 8);
    i64 *tmp = (i64 *)malloc((size_t)n * 8);
    for (i64 i = 0; i < n; i++) { key[i] = sortable(k, x, i); if (descending) key[i] = ~key[i]; perm[i] = i; }
    i64 *src = perm, *dst = tmp;
    for (i64 w = 1; w < n; w *= 2) {
        for (i64 lo = 0; lo < n; lo += 2 * w) {
            i64 mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            i64 a = lo, b = mid, o = lo;
            while (a < mid && b < hi) dst[o++] = (key[src[b]] < key[src[a]]) ? src[b++] : src[a++];
            while (a < mid) dst[o++] = src[a++];
            while (b < hi) dst[o++] = src[b++];
        }
        i64 *t = src; src = dst; dst = t;
    }
    if (src != perm) memcpy(perm, src, (size_t)n * 8);
    free(key); free(tmp);
    return RFO_OK;
}

/* ------------------------------------------------------------------ synthetic data */
uint64_t rfo_splitmix64(uint64_t seed, uint64_t i) {
    u64 z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
