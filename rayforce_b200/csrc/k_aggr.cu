// k_aggr.cu — grouped aggregates (sm_100a): rfb_aggr_dev = aggr_sum/min/max/count/avg (reference core/aggr.c AGGR_ITER :73-161,
// :1078-1453, :1455-2133) over group ids, plus the dispatch to the median / deviation pipelines of k_stats.cu.  Per-group
// accumulators: CTA-private shared memory at low cardinality (32-bit words with carry for the 64-bit integer sums,
// rfb_group.cuh; per-warp fp64 moments, rfb_moments.cuh), device-wide L2 atomics otherwise; finalised by a per-group kernel
// (sticky-null sums, +INF/NULL-initialised min/max, f64 averages).
#include "rfb_group.cuh"


namespace {

struct ValRow {
    const i64 *filter;
    __device__ __forceinline__ i64 operator()(i64 i) const { return filter ? ld_stream(filter + i) : i; }
};


__device__ __forceinline__ f64 key_to_f64(u64 k) {  // inverse of f64_sort_key; key 0 = null
    if (k == 0) return null_f64();
    return bits_f64((k & 0x8000000000000000ULL) ? (k & 0x7FFFFFFFFFFFFFFFULL) : ~k);
}

enum { AK_SUM_I64, AK_SUM_I16, AK_SUM_F64, AK_MIN_I64, AK_MAX_I64, AK_MIN_I32, AK_MAX_I32, AK_MIN_I16, AK_MAX_I16, AK_MIN_F64,
       AK_MAX_F64, AK_COUNT, AK_AVG_I64, AK_AVG_I32, AK_AVG_I16, AK_AVG_F64 };
// 16-bit values have no native atomics: I16 sums accumulate mod 2^32 and min/max in 32-bit slots of the workspace, and the
// finalise kernel narrows them (a sum mod 2^32 truncated to 16 bits is the reference's 16-bit wrapping sum)

// acc: main accumulator array (typed per kind), aux: null flags (sum) or non-null counts (avg)
template <int KIND, typename V> __device__ __forceinline__ void aggr_one(void *acc, void *aux, i64 g, V v) {
    if constexpr (KIND == AK_SUM_I64) {
        if (v == NULL_I64) ((u32 *)aux)[g] = 1u; else atomicAdd((unsigned long long *)acc + g, (unsigned long long)v);
    } else if constexpr (KIND == AK_SUM_I16) {
        if (v == NULL_I16) ((u32 *)aux)[g] = 1u; else atomicAdd((u32 *)acc + g, (u32)(i32)v);
    } else if constexpr (KIND == AK_SUM_F64) {
        if (isnan64(v)) ((u32 *)aux)[g] = 1u; else atomicAdd((f64 *)acc + g, v);
    } else if constexpr (KIND == AK_MIN_I64) {
        if (v != NULL_I64) atomicMin((long long *)acc + g, (long long)v);
    } else if constexpr (KIND == AK_MAX_I64) {
        atomicMax((long long *)acc + g, (long long)v);    // NULL is the minimum: never wins
    } else if constexpr (KIND == AK_MIN_I32) {
        if (v != NULL_I32) atomicMin((int *)acc + g, (int)v);
    } else if constexpr (KIND == AK_MAX_I32) {
        atomicMax((int *)acc + g, (int)v);
    } else if constexpr (KIND == AK_MIN_I16) {
        if (v != NULL_I16) atomicMin((int *)acc + g, (int)v);
    } else if constexpr (KIND == AK_MAX_I16) {
        atomicMax((int *)acc + g, (int)v);
    } else if constexpr (KIND == AK_MIN_F64) {
        if (!isnan64(v)) atomicMin((unsigned long long *)acc + g, (unsigned long long)f64_sort_key(v));
    } else if constexpr (KIND == AK_MAX_F64) {
        atomicMax((unsigned long long *)acc + g, (unsigned long long)f64_sort_key(v));   // NaN -> key 0: never wins
    } else if constexpr (KIND == AK_COUNT) {
        atomicAdd((unsigned long long *)acc + g, 1ULL);
    } else if constexpr (KIND == AK_AVG_I64) {
        if (v != NULL_I64) { atomicAdd((unsigned long long *)acc + g, (unsigned long long)v); atomicAdd((unsigned long long *)aux + g, 1ULL); }
    } else if constexpr (KIND == AK_AVG_I32) {
        if (v != NULL_I32) { atomicAdd((unsigned long long *)acc + g, (unsigned long long)(i64)v); atomicAdd((unsigned long long *)aux + g, 1ULL); }
    } else if constexpr (KIND == AK_AVG_I16) {
        if (v != NULL_I16) { atomicAdd((unsigned long long *)acc + g, (unsigned long long)(i64)v); atomicAdd((unsigned long long *)aux + g, 1ULL); }
    } else if constexpr (KIND == AK_AVG_F64) {
        if (!isnan64(v)) { atomicAdd((f64 *)acc + g, v); atomicAdd((unsigned long long *)aux + g, 1ULL); }
    }
}

// fold one CTA-private slot (shared memory) into the device-wide accumulators
template <int KIND> __device__ __forceinline__ void aggr_merge(void *acc, void *aux, const void *sacc, const void *saux, i64 g) {
    if constexpr (KIND == AK_SUM_I64 || KIND == AK_COUNT) {
        const u64 v = ((const u64 *)sacc)[g];
        if (v) atomicAdd((unsigned long long *)acc + g, (unsigned long long)v);
        if (KIND == AK_SUM_I64 && ((const u32 *)saux)[g]) ((u32 *)aux)[g] = 1u;
    } else if constexpr (KIND == AK_SUM_I16) {
        const u32 v = ((const u32 *)sacc)[g];
        if (v) atomicAdd((u32 *)acc + g, v);
        if (((const u32 *)saux)[g]) ((u32 *)aux)[g] = 1u;
    } else if constexpr (KIND == AK_MIN_I64) atomicMin((long long *)acc + g, ((const long long *)sacc)[g]);
    else if constexpr (KIND == AK_MAX_I64) atomicMax((long long *)acc + g, ((const long long *)sacc)[g]);
    else if constexpr (KIND == AK_MIN_I32 || KIND == AK_MIN_I16) atomicMin((int *)acc + g, ((const int *)sacc)[g]);
    else if constexpr (KIND == AK_MAX_I32 || KIND == AK_MAX_I16) atomicMax((int *)acc + g, ((const int *)sacc)[g]);
    else if constexpr (KIND == AK_MIN_F64) atomicMin((unsigned long long *)acc + g, ((const unsigned long long *)sacc)[g]);
    else if constexpr (KIND == AK_MAX_F64) atomicMax((unsigned long long *)acc + g, ((const unsigned long long *)sacc)[g]);
    else if constexpr (KIND == AK_AVG_I64 || KIND == AK_AVG_I32 || KIND == AK_AVG_I16) {
        const u64 c = ((const u64 *)saux)[g];
        if (c) { atomicAdd((unsigned long long *)acc + g, (unsigned long long)((const u64 *)sacc)[g]); atomicAdd((unsigned long long *)aux + g, (unsigned long long)c); }
    } else if constexpr (KIND == AK_AVG_F64) {
        const u64 c = ((const u64 *)saux)[g];
        if (c) { atomicAdd((f64 *)acc + g, ((const f64 *)sacc)[g]); atomicAdd((unsigned long long *)aux + g, (unsigned long long)c); }
    }
}

constexpr int PRIV_GROUPS = 3072;   // CTA-private accumulators in shared memory: 2 x 8 B x 3072 = 48 KB
constexpr int PRIV32_GROUPS = 4608; // 32-bit-word accumulators (sums, counts, averages of integers): 12 B x 4608 = 54 KB, 4 CTAs per SM

// PRIV: groups <= PRIV_GROUPS.  Every CTA folds its rows into shared-memory accumulators (initialised from the device-wide
// ones' initial values, which are each operator's identity) and merges them once at the end: the hot atomics never leave
// the SM, which is what low-cardinality keys (H2O id1-like, 100 groups) need — device-wide atomics serialise per address.
template <int KIND, typename V, bool PRIV>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_aggr(const V *__restrict__ val, ValRow row, const i64 *__restrict__ gid, i64 n, i64 groups, void *acc, void *aux) {
    constexpr int U = 4;
    extern __shared__ u64 s_priv[];
    void *a = acc, *x = aux;
    if constexpr (PRIV) {
        // private slots start at the operator's identity (NOT a copy of the device-wide slots: an early CTA may already
        // have merged into those)
        u64 *sa = s_priv, *sx = s_priv + groups;
        for (i64 g = threadIdx.x; g < groups; g += THREADS) {
            if constexpr (KIND == AK_MIN_I64) ((i64 *)sa)[g] = RFB_INF_I64;
            else if constexpr (KIND == AK_MAX_I64) ((i64 *)sa)[g] = NULL_I64;
            else if constexpr (KIND == AK_MIN_I32) ((i32 *)sa)[g] = (i32)0x7FFFFFFF;
            else if constexpr (KIND == AK_MAX_I32) ((i32 *)sa)[g] = NULL_I32;
            else if constexpr (KIND == AK_MIN_I16) ((i32 *)sa)[g] = (i32)0x7FFF;
            else if constexpr (KIND == AK_MAX_I16) ((i32 *)sa)[g] = (i32)NULL_I16;
            else if constexpr (KIND == AK_MIN_F64) sa[g] = f64_sort_key(bits_f64(0x7FF0000000000000ULL));
            else sa[g] = 0;   // sums, counts, averages, MAX_F64 (key 0 = null)
            sx[g] = 0;
        }
        __syncthreads();
        a = sa;
        x = sx;
    }
    const i64 stride = (i64)gridDim.x * THREADS;
    i64 i = (i64)blockIdx.x * THREADS + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        i64 g[U];
        V v[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            g[j] = ld_stream(gid + i + j * stride);
            if constexpr (KIND == AK_COUNT) v[j] = V();
            else v[j] = row.filter ? __ldg(val + row(i + j * stride)) : ld_stream(val + i + j * stride);
        }
#pragma unroll
        for (int j = 0; j < U; j++) aggr_one<KIND, V>(a, x, g[j], v[j]);
    }
    for (; i < n; i += stride) {
        V v;
        if constexpr (KIND == AK_COUNT) v = V();
        else v = row.filter ? __ldg(val + row(i)) : ld_stream(val + i);
        aggr_one<KIND, V>(a, x, ld_stream(gid + i), v);
    }
    if constexpr (PRIV) {
        __syncthreads();
        for (i64 g = threadIdx.x; g < groups; g += THREADS) aggr_merge<KIND>(acc, aux, a, x, g);
    }
}

// PRIV for the kinds whose accumulators are 64-bit integer sums and counts: 32-bit shared atomics with carry (SAcc) instead
// of 64-bit shared atomicAdd, which sm_100a runs as a compare-and-swap loop.  cnt word: SUM -> sticky-null flag only,
// COUNT -> rows, AVG -> non-null rows.
template <int KIND, typename V>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_aggr_priv32(const V *__restrict__ val, ValRow row, const i64 *__restrict__ gid, i64 n, int groups, void *acc, void *aux) {
    constexpr int U = 4;
    extern __shared__ u32 s_acc32[];
    const SAcc a{s_acc32, s_acc32 + groups, s_acc32 + 2 * groups};
    sacc_zero(a, groups);
    __syncthreads();
    auto one = [&](u32 g, V v) {
        if constexpr (KIND == AK_COUNT) atomicAdd(&a.cnt[g], 1u);
        else if constexpr (KIND == AK_SUM_I64) {
            if (v == NULL_I64) atomicOr(&a.cnt[g], NULL_FLAG);
            else {
                const u32 lo = (u32)(u64)v;
                u32 hi = (u32)((u64)v >> 32);
                const u32 old = atomicAdd(&a.lo[g], lo);
                hi += (u32)((u32)(old + lo) < lo);
                if (hi) atomicAdd(&a.hi[g], hi);
            }
        } else {   // averages: i64 sum of the non-null values + their count
            if (!Elem<V>::is_null(v)) sacc_add(a, g, (i64)v);
        }
    };
    const i64 stride = (i64)gridDim.x * THREADS;
    i64 i = (i64)blockIdx.x * THREADS + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        i64 g[U];
        V v[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            g[j] = ld_stream(gid + i + j * stride);
            if constexpr (KIND == AK_COUNT) v[j] = V();
            else v[j] = row.filter ? __ldg(val + row(i + j * stride)) : ld_stream(val + i + j * stride);
        }
#pragma unroll
        for (int j = 0; j < U; j++) one((u32)g[j], v[j]);
    }
    for (; i < n; i += stride) {
        V v;
        if constexpr (KIND == AK_COUNT) v = V();
        else v = row.filter ? __ldg(val + row(i)) : ld_stream(val + i);
        one((u32)ld_stream(gid + i), v);
    }
    __syncthreads();
    for (int g = threadIdx.x; g < groups; g += THREADS) {
        const u32 c = a.cnt[g];
        const u64 sum = ((u64)a.hi[g] << 32) | a.lo[g];
        if constexpr (KIND == AK_COUNT) { if (c) atomicAdd((unsigned long long *)acc + g, (unsigned long long)c); }
        else if constexpr (KIND == AK_SUM_I64) {
            if (sum) atomicAdd((unsigned long long *)acc + g, (unsigned long long)sum);
            if (c & NULL_FLAG) ((u32 *)aux)[g] = 1u;
        } else if (c) {
            atomicAdd((unsigned long long *)acc + g, (unsigned long long)sum);
            atomicAdd((unsigned long long *)aux + g, (unsigned long long)(c & ~NULL_FLAG));
        }
    }
}

template <int KIND> __global__ void k_aggr_final(void *out, const void *acc, const void *aux, i64 groups) {
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (i64)gridDim.x * blockDim.x) {
        if constexpr (KIND == AK_SUM_I64) { if (((const u32 *)aux)[g]) ((i64 *)out)[g] = NULL_I64; }
        else if constexpr (KIND == AK_SUM_I16) ((i16 *)out)[g] = ((const u32 *)aux)[g] ? NULL_I16 : (i16)(unsigned short)((const u32 *)acc)[g];
        else if constexpr (KIND == AK_MIN_I16 || KIND == AK_MAX_I16) ((i16 *)out)[g] = (i16)((const i32 *)acc)[g];
        else if constexpr (KIND == AK_SUM_F64) { if (((const u32 *)aux)[g]) ((f64 *)out)[g] = null_f64(); }
        else if constexpr (KIND == AK_MIN_F64 || KIND == AK_MAX_F64) ((f64 *)out)[g] = key_to_f64(((const u64 *)acc)[g]);
        else if constexpr (KIND == AK_AVG_I64 || KIND == AK_AVG_I32 || KIND == AK_AVG_I16) {
            const i64 c = ((const i64 *)aux)[g];
            ((f64 *)out)[g] = c == 0 ? null_f64() : __ddiv_rn((f64)((const i64 *)acc)[g], (f64)c);
        } else if constexpr (KIND == AK_AVG_F64) {
            const i64 c = ((const i64 *)aux)[g];
            ((f64 *)out)[g] = c == 0 ? null_f64() : __ddiv_rn(((const f64 *)acc)[g], (f64)c);
        }
    }
}

template <int KIND, typename V>
int run_aggr(rfb_ctx_t *ctx, const void *val, const i64 *filter, const i64 *gid, i64 len, i64 groups, void *acc, void *aux, void *out) {
    if (len > 0) {
        const int grid = rfb_grid_for(ctx, len, THREADS * 4, BLOCKS_PER_SM);
        // F64 sums stay on the device-wide path (one accumulator per group, like the reference's single running sum)
        if constexpr (KIND == AK_SUM_I64 || KIND == AK_COUNT || KIND == AK_AVG_I64 || KIND == AK_AVG_I32 || KIND == AK_AVG_I16) {
            if (groups <= PRIV32_GROUPS && len >= 65536 && len / grid < (1ll << 31)) {
                RFB_CUDA(cudaFuncSetAttribute(k_aggr_priv32<KIND, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, PRIV32_GROUPS * 12));
                k_aggr_priv32<KIND, V><<<grid, THREADS, (size_t)groups * 12, ctx->stream>>>((const V *)val, ValRow{filter}, gid, len, (int)groups, acc, aux);
                RFB_CHECK_LAUNCH(ctx);
                goto finalise;
            }
        }
        if constexpr (KIND != AK_SUM_F64) {
            if (groups <= PRIV_GROUPS && len >= 65536) {
                k_aggr<KIND, V, true><<<grid, THREADS, (size_t)groups * 16, ctx->stream>>>((const V *)val, ValRow{filter}, gid, len, groups, acc, aux);
                RFB_CHECK_LAUNCH(ctx);
                goto finalise;
            }
        }
        k_aggr<KIND, V, false><<<grid, THREADS, 0, ctx->stream>>>((const V *)val, ValRow{filter}, gid, len, groups, acc, aux);
        RFB_CHECK_LAUNCH(ctx);
    }
finalise:
    k_aggr_final<KIND><<<rfb_grid_for(ctx, groups, 256, 8), 256, 0, ctx->stream>>>(out, acc, aux, groups);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

// ---- aggr_first / aggr_last (core/aggr.c:441-577, 851-1075): one POSITION per group, then a gather.
// first: the group's first row, whatever its value (the reference's fast path reads in[first_ids[g]]; every index this
//        library builds carries first_ids).  last: the group's last NON-NULL value, null when it has none — what the reference
//        computes for one chunk of rows (aggr_last_partial); with several worker chunks its merge keeps the FIRST chunk's
//        answer (AGGR_COLLECT `if (out == NULL) out = in`), so above its 16384-row parallel threshold its own result depends
//        on the thread count.
// LAST over worker chunks: position key = chunk << 40 | (2^40 - 1 - row), minimised: the first chunk with a non-null value wins,
// inside it the last row.  chunk = row / chunk_rows capped at nchunks - 1 (the last chunk takes the remainder rows).
constexpr u64 POS_MASK = (1ull << 40) - 1;
template <typename V, bool LAST>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_aggr_pos(const V *__restrict__ val, ValRow row, const i64 *__restrict__ gid, i64 len, i64 chunk_rows, i64 nchunks, u64 *pos) {
    for (i64 i = (i64)blockIdx.x * THREADS + threadIdx.x; i < len; i += (i64)gridDim.x * THREADS) {
        const i64 g = ld_stream(gid + i);
        if constexpr (LAST) {
            if (Elem<V>::is_null(val[row(i)])) continue;
            u64 c = 0;
            if (nchunks > 1) {
                c = len < (1ll << 32) ? (u64)((u32)i / (u32)chunk_rows) : (u64)i / (u64)chunk_rows;
                if (c > (u64)(nchunks - 1)) c = (u64)(nchunks - 1);
            }
            const u64 key = (c << 40) | (POS_MASK - (u64)i);
            if (__ldcg(&pos[g]) > key) atomicMin((unsigned long long *)&pos[g], (unsigned long long)key);
        } else {
            if (__ldcg(&pos[g]) > (u64)i) atomicMin((unsigned long long *)&pos[g], (unsigned long long)i);
        }
    }
}
template <typename V, bool LAST>
__global__ void __launch_bounds__(256) k_aggr_pos_final(const V *__restrict__ val, ValRow row, const u64 *__restrict__ pos, i64 groups, V *out) {
    for (i64 g = (i64)blockIdx.x * 256 + threadIdx.x; g < groups; g += (i64)gridDim.x * 256) {
        const u64 p = pos[g];
        if (p == NO_ROW) out[g] = Elem<V>::null();
        else out[g] = val[row(LAST ? (i64)(POS_MASK - (p & POS_MASK)) : (i64)p)];
    }
}
template <typename V, bool LAST>
int run_pos(rfb_ctx_t *ctx, const void *val, const i64 *filter, const i64 *gid, i64 len, i64 groups, i64 nchunks, u64 *pos, void *out) {
    if (len >= (1ll << 40)) { rfb_set_error("aggr_first / aggr_last: more than 2^40 rows"); return RFB_ERR_ARG; }
    if (nchunks < 1 || len / nchunks == 0) nchunks = 1;
    RFB_CUDA(cudaMemsetAsync(pos, 0xFF, (size_t)groups * 8, ctx->stream));
    if (len > 0) {
        k_aggr_pos<V, LAST><<<rfb_grid_for(ctx, len, THREADS * 4, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>((const V *)val, ValRow{filter}, gid, len, len / nchunks, nchunks, pos);
        RFB_CHECK_LAUNCH(ctx);
    }
    k_aggr_pos_final<V, LAST><<<rfb_grid_for(ctx, groups, 256, 8), 256, 0, ctx->stream>>>((const V *)val, ValRow{filter}, pos, groups, (V *)out);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

int pos_dispatch(rfb_ctx_t *ctx, bool last, int k, const void *val, const i64 *filter, const i64 *gid, i64 len, i64 groups, i64 nchunks, u64 *pos, void *out) {
    switch (k) {
        case K_U8: return run_pos<u8, false>(ctx, val, filter, gid, len, groups, 1, pos, out);
        case K_I16: return last ? run_pos<i16, true>(ctx, val, filter, gid, len, groups, nchunks, pos, out) : run_pos<i16, false>(ctx, val, filter, gid, len, groups, 1, pos, out);
        case K_I32: return last ? run_pos<i32, true>(ctx, val, filter, gid, len, groups, nchunks, pos, out) : run_pos<i32, false>(ctx, val, filter, gid, len, groups, 1, pos, out);
        case K_I64: return last ? run_pos<i64, true>(ctx, val, filter, gid, len, groups, nchunks, pos, out) : run_pos<i64, false>(ctx, val, filter, gid, len, groups, 1, pos, out);
        default: return last ? run_pos<f64, true>(ctx, val, filter, gid, len, groups, nchunks, pos, out) : run_pos<f64, false>(ctx, val, filter, gid, len, groups, 1, pos, out);
    }
}

}  // namespace

extern "C" int rfb_aggr_type(int op, int val_type) {
    const int k = rfb_kind_of(val_type);
    if (!k) return RFB_ERR_TYPE;
    switch (op) {
        case RFB_A_COUNT: return (k == K_U8 || k == K_I16) ? RFB_ERR_TYPE : RFB_I64;
        // the non-parted drivers' switch tables: aggr_sum core/aggr.c:1107-1150, aggr_max/min :1152-1315, aggr_avg :2013-2133
        case RFB_A_SUM: return (val_type == RFB_I16 || val_type == RFB_I64 || k == K_F64) ? val_type : RFB_ERR_TYPE;
        case RFB_A_MIN: case RFB_A_MAX:
            return (val_type == RFB_I16 || val_type == RFB_I64 || val_type == RFB_TIMESTAMP || val_type == RFB_DATE || val_type == RFB_TIME || val_type == RFB_F64) ? val_type : RFB_ERR_TYPE;
        case RFB_A_AVG: return (val_type == RFB_I16 || k == K_I32 || val_type == RFB_I64 || k == K_F64) ? RFB_F64 : RFB_ERR_TYPE;
        case RFB_A_MED: return RFB_F64;   // aggr_collect takes every column type; types without a median give nulls (core/aggr.c:2182-2184)
        case RFB_A_FIRST: return val_type;   // every fixed-width type (core/aggr.c:455-573)
        case RFB_A_LAST: return k == K_U8 ? RFB_ERR_TYPE : val_type;   // aggr_last has no U8/B8 case (core/aggr.c:904-931)
        case RFB_A_DEV: return (val_type == RFB_I16 || k == K_I32 || val_type == RFB_I64 || val_type == RFB_TIMESTAMP || k == K_F64) ? RFB_F64 : RFB_ERR_TYPE;   // core/aggr.c:2873-2880
        default: return RFB_ERR_TYPE;
    }
}

extern "C" int rfb_aggr_dev(rfb_ctx_t *ctx, int op, int val_type, const void *val, const int64_t *filter,
                            const int64_t *group_ids, int64_t len, int64_t groups, void *out) {
    RFB_ARG(ctx && len >= 0 && groups >= 0 && ((val && group_ids) || len == 0) && (out || groups == 0), "rfb_aggr_dev");
    const int ot = rfb_aggr_type(op, val_type);
    if (ot < 0) { rfb_set_error("aggr %d: unsupported value type %d", op, val_type); return RFB_ERR_TYPE; }
    if (groups == 0) return RFB_OK;
    if (op == RFB_A_MED) return rfb_aggr_med_launch(ctx, val_type, val, filter, group_ids, len, groups, (f64 *)out);
    if (op == RFB_A_DEV) return rfb_aggr_stddev_launch(ctx, val_type, val, filter, group_ids, len, groups, (f64 *)out);
    const int k = rfb_kind_of(val_type);
    void *w;
    int rc = rfb_ensure_work(ctx, 2 * align256((size_t)groups * 8), &w);
    if (rc) return rc;
    void *acc = w, *aux = (char *)w + align256((size_t)groups * 8);
    // 1e4 .. 2.6e5 groups of i64 values: sum / avg through the key-range partition passes of the fused group-by (group id = key)
    if ((op == RFB_A_SUM || op == RFB_A_AVG) && k == K_I64 && !filter && groups > PRIV32_GROUPS && len >= 65536) {
        const size_t g8 = align256((size_t)groups * 8);
        rc = rfb_ensure_work(ctx, 3 * g8 + rfb_narrow_sums_bytes(len), &w);
        if (rc) return rc;
        acc = w; aux = (char *)w + g8;
        void *third = (char *)w + 2 * g8, *store = (char *)w + 3 * g8;
        bool done = false;
        if (op == RFB_A_SUM) {       // sums straight into `out`, sticky-null flags in aux, row counts (unused) in the third array
            RFB_CUDA(cudaMemsetAsync(out, 0, (size_t)groups * 8, ctx->stream));
            RFB_CUDA(cudaMemsetAsync(aux, 0, 2 * g8, ctx->stream));
            rc = rfb_narrow_sums(ctx, group_ids, (const i64 *)val, len, groups, store, (u64 *)out, (u64 *)third, (u32 *)aux, true, &done);
            if (rc) return rc;
            if (done) {
                k_aggr_final<AK_SUM_I64><<<rfb_grid_for(ctx, groups, 256, 8), 256, 0, ctx->stream>>>(out, out, aux, groups);
                RFB_CHECK_LAUNCH(ctx);
                return RFB_OK;
            }
        } else {                     // sums in acc, non-null counts in aux, null flags (unused) in the third array
            RFB_CUDA(cudaMemsetAsync(acc, 0, 3 * g8, ctx->stream));
            rc = rfb_narrow_sums(ctx, group_ids, (const i64 *)val, len, groups, store, (u64 *)acc, (u64 *)aux, (u32 *)third, false, &done);
            if (rc) return rc;
            if (done) {
                k_aggr_final<AK_AVG_I64><<<rfb_grid_for(ctx, groups, 256, 8), 256, 0, ctx->stream>>>(out, acc, aux, groups);
                RFB_CHECK_LAUNCH(ctx);
                return RFB_OK;
            }
        }
    }
    if (op == RFB_A_FIRST || op == RFB_A_LAST) return pos_dispatch(ctx, op == RFB_A_LAST, k, val, filter, group_ids, len, groups, 1, (u64 *)acc, out);
    switch (op) {
        case RFB_A_COUNT:
            RFB_CUDA(cudaMemsetAsync(out, 0, (size_t)groups * 8, ctx->stream));
            return run_aggr<AK_COUNT, i64>(ctx, nullptr, nullptr, group_ids, len, groups, out, aux, out);
        case RFB_A_SUM:
            RFB_CUDA(cudaMemsetAsync(out, 0, (size_t)groups * rfb_type_size(val_type), ctx->stream));
            RFB_CUDA(cudaMemsetAsync(aux, 0, (size_t)groups * 4, ctx->stream));
            if (k == K_I64) return run_aggr<AK_SUM_I64, i64>(ctx, val, filter, group_ids, len, groups, out, aux, out);
            if (k == K_I16) {
                RFB_CUDA(cudaMemsetAsync(acc, 0, (size_t)groups * 4, ctx->stream));
                return run_aggr<AK_SUM_I16, i16>(ctx, val, filter, group_ids, len, groups, acc, aux, out);
            }
            // f64 sums: per-group moments with shared-memory privatisation at low cardinality (rfb_moments.cuh); the count it
            // also produces lands in the workspace and is not used
            RFB_CUDA(cudaMemsetAsync(acc, 0, (size_t)groups * 8, ctx->stream));
            rc = moments::launch<f64, false>(ctx, val, filter, group_ids, len, groups, (f64 *)out, nullptr, (unsigned long long *)acc, (u32 *)aux);
            if (rc) return rc;
            k_aggr_final<AK_SUM_F64><<<rfb_grid_for(ctx, groups, 256, 8), 256, 0, ctx->stream>>>(out, out, aux, groups);
            RFB_CHECK_LAUNCH(ctx);
            return RFB_OK;
        case RFB_A_MIN:
            if (k == K_I64) { rc = fill<i64>(ctx, (i64 *)out, groups, RFB_INF_I64); if (rc) return rc; return run_aggr<AK_MIN_I64, i64>(ctx, val, filter, group_ids, len, groups, out, aux, out); }
            if (k == K_I32) { rc = fill<i32>(ctx, (i32 *)out, groups, (i32)0x7FFFFFFF); if (rc) return rc; return run_aggr<AK_MIN_I32, i32>(ctx, val, filter, group_ids, len, groups, out, aux, out); }
            if (k == K_I16) { rc = fill<i32>(ctx, (i32 *)acc, groups, (i32)0x7FFF); if (rc) return rc; return run_aggr<AK_MIN_I16, i16>(ctx, val, filter, group_ids, len, groups, acc, aux, out); }
            rc = fill<u64>(ctx, (u64 *)acc, groups, f64_sort_key(bits_f64(0x7FF0000000000000ULL)));
            if (rc) return rc;
            return run_aggr<AK_MIN_F64, f64>(ctx, val, filter, group_ids, len, groups, acc, aux, out);
        case RFB_A_MAX:
            if (k == K_I64) { rc = fill<i64>(ctx, (i64 *)out, groups, NULL_I64); if (rc) return rc; return run_aggr<AK_MAX_I64, i64>(ctx, val, filter, group_ids, len, groups, out, aux, out); }
            if (k == K_I32) { rc = fill<i32>(ctx, (i32 *)out, groups, NULL_I32); if (rc) return rc; return run_aggr<AK_MAX_I32, i32>(ctx, val, filter, group_ids, len, groups, out, aux, out); }
            if (k == K_I16) { rc = fill<i32>(ctx, (i32 *)acc, groups, (i32)NULL_I16); if (rc) return rc; return run_aggr<AK_MAX_I16, i16>(ctx, val, filter, group_ids, len, groups, acc, aux, out); }
            RFB_CUDA(cudaMemsetAsync(acc, 0, (size_t)groups * 8, ctx->stream));
            return run_aggr<AK_MAX_F64, f64>(ctx, val, filter, group_ids, len, groups, acc, aux, out);
        default:  // AVG
            RFB_CUDA(cudaMemsetAsync(acc, 0, (size_t)groups * 8, ctx->stream));
            RFB_CUDA(cudaMemsetAsync(aux, 0, (size_t)groups * 8, ctx->stream));
            if (k == K_I64) return run_aggr<AK_AVG_I64, i64>(ctx, val, filter, group_ids, len, groups, acc, aux, out);
            if (k == K_I32) return run_aggr<AK_AVG_I32, i32>(ctx, val, filter, group_ids, len, groups, acc, aux, out);
            if (k == K_I16) return run_aggr<AK_AVG_I16, i16>(ctx, val, filter, group_ids, len, groups, acc, aux, out);
            rc = moments::launch<f64, false>(ctx, val, filter, group_ids, len, groups, (f64 *)acc, nullptr, (unsigned long long *)aux, nullptr);
            if (rc) return rc;
            k_aggr_final<AK_AVG_F64><<<rfb_grid_for(ctx, groups, 256, 8), 256, 0, ctx->stream>>>(out, acc, aux, groups);
            RFB_CHECK_LAUNCH(ctx);
            return RFB_OK;
    }
}

extern "C" int rfb_aggr_last_dev(rfb_ctx_t *ctx, int val_type, const void *val, const int64_t *filter, const int64_t *group_ids,
                                 int64_t len, int64_t groups, int64_t nchunks, void *out) {
    RFB_ARG(ctx && len >= 0 && groups >= 0 && ((val && group_ids) || len == 0) && (out || groups == 0), "rfb_aggr_last_dev");
    if (rfb_aggr_type(RFB_A_LAST, val_type) < 0) { rfb_set_error("aggr_last: unsupported value type %d", val_type); return RFB_ERR_TYPE; }
    if (groups == 0) return RFB_OK;
    void *w;
    int rc = rfb_ensure_work(ctx, align256((size_t)groups * 8), &w);
    if (rc) return rc;
    return pos_dispatch(ctx, true, rfb_kind_of(val_type), val, filter, group_ids, len, groups, nchunks, (u64 *)w, out);
}
