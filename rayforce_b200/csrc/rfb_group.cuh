// rfb_group.cuh — device code shared by the grouping kernels: k_group.cu (group index, distinct, multi-key grouping),
// k_aggr.cu (grouped aggregates) and k_fused_group.cu (fused group-by): row sources and slot functions, the open-addressing
// table, first-row claims, the scope (min/max) kernels, and the 32-bit-word shared-memory accumulators.  Everything lives in an
// anonymous namespace: each translation unit gets its own copy.
#pragma once

#include "rfb_scan.cuh"
#include "rfb_moments.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int BLOCKS_PER_SM = 4;
constexpr u64 NO_ROW = ~0ULL;
constexpr int NUM_J = 8;  // rows per lane in the numbering pass

template <typename T> __global__ void k_fill(T *p, i64 n, T v) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) p[i] = v;
}
template <typename T> int fill(rfb_ctx_t *ctx, T *p, i64 n, T v) {
    if (n <= 0) return RFB_OK;
    k_fill<T><<<rfb_grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(p, n, v);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

// ------------------------------------------------------------------ row sources and slot functions

// position i of the (filtered) row sequence -> key
struct KeySrc {
    const i64 *keys;
    const i64 *filter;  // or nullptr
    __device__ __forceinline__ i64 operator()(i64 i) const { return filter ? __ldg(keys + ld_stream(filter + i)) : ld_stream(keys + i); }
};

struct DenseSlot {
    i64 min;
    __device__ __forceinline__ i64 operator()(i64 key) const { return (i64)((u64)key - (u64)min); }
};

__host__ __device__ __forceinline__ u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// open-addressing table: tk[cap] keys (EMPTY = NULL_I64, like the reference's ht_oa tables, core/hash.c:35-56) plus one
// dedicated slot `cap` for the NULL_I64 key itself (which the reference's table cannot represent).
struct HashSlot {
    i64 *tk;
    u64 mask;
    i64 cap;
    __device__ __forceinline__ i64 insert(i64 key) const {
        if (key == NULL_I64) return cap;
        u64 s = mix64((u64)key) & mask;
        while (true) {
            i64 cur = (i64)scan::ld_relaxed((const u64 *)&tk[s]);
            if (cur == key) return (i64)s;
            if (cur == NULL_I64) {
                const i64 old = (i64)atomicCAS((unsigned long long *)&tk[s], (unsigned long long)NULL_I64, (unsigned long long)key);
                if (old == NULL_I64 || old == key) return (i64)s;
            }
            s = (s + 1) & mask;
        }
    }
    __device__ __forceinline__ i64 operator()(i64 key) const {  // lookup of a key known to be present
        if (key == NULL_I64) return cap;
        u64 s = mix64((u64)key) & mask;
        while (__ldg(&tk[s]) != key) s = (s + 1) & mask;
        return (i64)s;
    }
};

__device__ __forceinline__ void claim_first(u64 *first_row, i64 slot, i64 row) {
    if (__ldcg(&first_row[slot]) > (u64)row) atomicMin((unsigned long long *)&first_row[slot], (unsigned long long)row);
}

// ------------------------------------------------------------------ 1. scope

__global__ void k_scope_init(i64 *mm) { mm[0] = RFB_INF_I64; mm[1] = NULL_I64; mm[2] = 0; }

template <typename Src>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM) k_scope(Src src, i64 n, i64 *mm) {
    __shared__ i64 red[32];
    i64 lo = RFB_INF_I64, hi = NULL_I64;
    constexpr int U = 4;
    const i64 stride = (i64)gridDim.x * THREADS;
    i64 i = (i64)blockIdx.x * THREADS + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        i64 k[U];
#pragma unroll
        for (int j = 0; j < U; j++) k[j] = src(i + j * stride);
#pragma unroll
        for (int j = 0; j < U; j++) { lo = k[j] < lo ? k[j] : lo; hi = k[j] > hi ? k[j] : hi; }
    }
    for (; i < n; i += stride) { const i64 k = src(i); lo = k < lo ? k : lo; hi = k > hi ? k : hi; }
    struct Mn { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b < a ? b : a; } };
    struct Mx { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b > a ? b : a; } };
    lo = block_reduce<i64>(lo, Mn(), RFB_INF_I64, red);
    hi = block_reduce<i64>(hi, Mx(), NULL_I64, red);
    if (threadIdx.x == 0) {
        atomicMin((long long *)&mm[0], (long long)lo);
        atomicMax((long long *)&mm[1], (long long)hi);
    }
}

// unfiltered, 16-byte aligned key column: two keys per 128-bit load, four loads in flight per thread
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM) k_scope_vec(const i64 *__restrict__ keys, i64 n, i64 *mm) {
    __shared__ i64 red[32];
    i64 lo = RFB_INF_I64, hi = NULL_I64;
    constexpr int U = 4;
    const i64 pairs = n >> 1, stride = (i64)gridDim.x * THREADS;
    i64 i = (i64)blockIdx.x * THREADS + threadIdx.x;
    for (; i + (U - 1) * stride < pairs; i += U * stride) {
        vec16 v[U];
#pragma unroll
        for (int j = 0; j < U; j++) v[j] = ld_stream16(keys + 2 * (i + j * stride));
#pragma unroll
        for (int j = 0; j < U; j++) {
            const i64 a = (i64)v[j].lo, b = (i64)v[j].hi;
            lo = a < lo ? a : lo; hi = a > hi ? a : hi;
            lo = b < lo ? b : lo; hi = b > hi ? b : hi;
        }
    }
    for (; i < pairs; i += stride) {
        const vec16 v = ld_stream16(keys + 2 * i);
        const i64 a = (i64)v.lo, b = (i64)v.hi;
        lo = a < lo ? a : lo; hi = a > hi ? a : hi;
        lo = b < lo ? b : lo; hi = b > hi ? b : hi;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) { const i64 a = keys[n - 1]; lo = a < lo ? a : lo; hi = a > hi ? a : hi; }
    struct Mn { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b < a ? b : a; } };
    struct Mx { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b > a ? b : a; } };
    lo = block_reduce<i64>(lo, Mn(), RFB_INF_I64, red);
    hi = block_reduce<i64>(hi, Mx(), NULL_I64, red);
    if (threadIdx.x == 0) {
        atomicMin((long long *)&mm[0], (long long)lo);
        atomicMax((long long *)&mm[1], (long long)hi);
    }
}

// group_ids[row] = gid_of_slot[key - min] for an unfiltered dense key column, two rows per 128-bit load / store
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_assign_dense_vec(const i64 *__restrict__ keys, i64 kmin, i64 n, const i64 *__restrict__ gid_of_slot, i64 *__restrict__ group_ids) {
    constexpr int U = 4;
    const i64 pairs = n >> 1, stride = (i64)gridDim.x * THREADS;
    i64 i = (i64)blockIdx.x * THREADS + threadIdx.x;
    for (; i + (U - 1) * stride < pairs; i += U * stride) {
        vec16 v[U], g[U];
#pragma unroll
        for (int j = 0; j < U; j++) v[j] = ld_stream16(keys + 2 * (i + j * stride));
#pragma unroll
        for (int j = 0; j < U; j++) { g[j].lo = (u64)__ldg(&gid_of_slot[(i64)(v[j].lo - (u64)kmin)]); g[j].hi = (u64)__ldg(&gid_of_slot[(i64)(v[j].hi - (u64)kmin)]); }
#pragma unroll
        for (int j = 0; j < U; j++) st_stream16(group_ids + 2 * (i + j * stride), g[j]);
    }
    for (; i < pairs; i += stride) {
        const vec16 v = ld_stream16(keys + 2 * i);
        vec16 g;
        g.lo = (u64)__ldg(&gid_of_slot[(i64)(v.lo - (u64)kmin)]);
        g.hi = (u64)__ldg(&gid_of_slot[(i64)(v.hi - (u64)kmin)]);
        st_stream16(group_ids + 2 * i, g);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) group_ids[n - 1] = gid_of_slot[(i64)((u64)keys[n - 1] - (u64)kmin)];
}

// ------------------------------------------------------------------ 2. claim

template <typename Src, typename Slot, bool INSERT>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM) k_claim(Src src, Slot slot, i64 r0, i64 n, u64 *first_row) {   // rows [r0, n)
    constexpr int U = 4;
    const i64 stride = (i64)gridDim.x * THREADS;
    i64 i = r0 + (i64)blockIdx.x * THREADS + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        i64 k[U];
#pragma unroll
        for (int j = 0; j < U; j++) k[j] = src(i + j * stride);
#pragma unroll
        for (int j = 0; j < U; j++) {
            i64 s;
            if constexpr (INSERT) s = slot.insert(k[j]); else s = slot(k[j]);
            claim_first(first_row, s, i + j * stride);
        }
    }
    for (; i < n; i += stride) {
        const i64 k = src(i);
        i64 s;
        if constexpr (INSERT) s = slot.insert(k); else s = slot(k);
        claim_first(first_row, s, i);
    }
}

// mm[4] = slots whose first row is known
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM) k_claimed_count(const u64 *first_row, i64 slots, i64 *mm) {
    __shared__ i64 red[32];
    i64 claimed = 0;
    for (i64 s = (i64)blockIdx.x * THREADS + threadIdx.x; s < slots; s += (i64)gridDim.x * THREADS) claimed += first_row[s] != NO_ROW;
    claimed = block_reduce<i64>(claimed, OpAddWrap(), 0, red);
    if (threadIdx.x == 0) atomicAdd((unsigned long long *)&mm[4], (unsigned long long)claimed);
}

// max over the claimed first rows (bounds the numbering pass)
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM) k_max_first(const u64 *first_row, i64 slots, i64 *mm) {
    __shared__ i64 red[32];
    i64 hi = -1;
    for (i64 i = (i64)blockIdx.x * THREADS + threadIdx.x; i < slots; i += (i64)gridDim.x * THREADS) {
        const u64 f = first_row[i];
        if (f != NO_ROW && (i64)f > hi) hi = (i64)f;
    }
    struct Mx { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b > a ? b : a; } };
    hi = block_reduce<i64>(hi, Mx(), (i64)-1, red);
    if (threadIdx.x == 0) atomicMax((long long *)&mm[2], (long long)(hi + 1));
}

// ------------------------------------------------------------------ 3. number

template <typename Src, typename Slot>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_number(Src src, Slot slot, i64 limit, const u64 *first_row, i64 *gid_of_slot, i64 *first_ids, scan::TileCtl ctl) {
    __shared__ scan::TileSmem sm;
    scan::compact_rows<NUM_J>(
        limit, ctl, sm, [&](i64 r) { return __ldcg(&first_row[slot(src(r))]) == (u64)r; },
        [&](i64 r, i64 g) {
            first_ids[g] = r;
            gid_of_slot[slot(src(r))] = g;
        });
}

// ------------------------------------------------------------------ 4. assign

template <typename Src, typename Slot>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_assign(Src src, Slot slot, i64 n, const i64 *__restrict__ gid_of_slot, i64 *__restrict__ group_ids) {
    constexpr int U = 4;
    const i64 stride = (i64)gridDim.x * THREADS;
    i64 i = (i64)blockIdx.x * THREADS + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        i64 k[U], g[U];
#pragma unroll
        for (int j = 0; j < U; j++) k[j] = src(i + j * stride);
#pragma unroll
        for (int j = 0; j < U; j++) g[j] = __ldg(&gid_of_slot[slot(k[j])]);
#pragma unroll
        for (int j = 0; j < U; j++) __stcs(group_ids + i + j * stride, g[j]);
    }
    for (; i < n; i += stride) group_ids[i] = __ldg(&gid_of_slot[slot(src(i))]);
}

int d2h_sync(rfb_ctx_t *ctx, void *dst, const void *src, size_t bytes) {
    RFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return RFB_OK;
}

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

// ---- shared-memory accumulators (32-bit words)
constexpr u32 NULL_FLAG = 0x80000000u;
struct SAcc { u32 *lo, *hi, *cnt; };

__device__ __forceinline__ void sacc_add(const SAcc &a, u32 s, i64 v) {
    if (v == NULL_I64) atomicOr(&a.cnt[s], NULL_FLAG);
    else {
        const u32 lo = (u32)(u64)v;
        u32 hi = (u32)((u64)v >> 32);
        const u32 old = atomicAdd(&a.lo[s], lo);
        hi += (u32)((u32)(old + lo) < lo);   // this row wrapped the low word: carry
        if (hi) atomicAdd(&a.hi[s], hi);
    }
    atomicAdd(&a.cnt[s], 1u);
}
// the same for one row per lane of a full warp: when all 32 lanes are selected, non-null and hit the SAME slot (a single-key or
// heavily skewed column) the warp adds its values with shuffles and issues one atomic set instead of 32 conflicting ones
__device__ __forceinline__ void sacc_add_warp(const SAcc &a, bool sel, u32 s, i64 v) {
    const u32 s0 = __shfl_sync(0xffffffffu, s, 0);
    if (__all_sync(0xffffffffu, sel && s == s0 && v != NULL_I64)) {
        u64 t = (u64)v;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if ((threadIdx.x & 31) == 0) {
            const u32 lo = (u32)t;
            u32 hi = (u32)(t >> 32);
            const u32 old = atomicAdd(&a.lo[s0], lo);
            hi += (u32)((u32)(old + lo) < lo);
            if (hi) atomicAdd(&a.hi[s0], hi);
            atomicAdd(&a.cnt[s0], 32u);
        }
    } else if (sel) sacc_add(a, s, v);
}
__device__ __forceinline__ void sacc_zero(const SAcc &a, int slots) {
    for (int s = threadIdx.x; s < slots; s += blockDim.x) { a.lo[s] = 0; a.hi[s] = 0; a.cnt[s] = 0; }
}

}  // namespace
