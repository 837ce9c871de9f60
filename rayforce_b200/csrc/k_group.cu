// k_group.cu — hash / perfect-hash group-by and grouped aggregates (sm_100a).
//
//   rfb_group_i64_dev         index_group -> index_group_i64: scope + group numbering (reference core/index.c:402-435,
//                             2002-2092 perfect hash, 1777-1911 + core/hash.c open addressing)
//   rfb_aggr_dev              aggr_sum/min/max/count/avg (core/aggr.c AGGR_ITER :73-161, :1078-1453, :1455-2133)
//   rfb_group_sum_count_dev   select {s: (sum v) c: (count v) from t by k [where ...]} fused: never materialises group ids
//
// Group numbering must be the reference's: groups are numbered in order of first occurrence in (filtered) row order
// (core/index.c:2037-2055, sequential there).  The device does it without a sequential scan:
//   1. scope     min/max of the keys (one streaming reduction)                                   -> dense or sparse?
//   2. claim     every row does first_row[slot(key)] = min(first_row[slot], row): a plain load filters out almost every
//                atomic once a slot has been claimed by an earlier row (values only decrease, so a stale read is safe)
//                dense  : slot = key - min (direct addressing, "perfect hash")
//                sparse : slot = open-addressing table in HBM/L2, 64-bit CAS insert, linear probing, load factor <= 0.5
//   3. number    rows [0, max(first_row)] are compacted in row order by the predicate first_row[slot(key_row)] == row
//                (the chained-scan compaction of rfb_scan.cuh): the g-th such row IS the first row of group g
//                -> first_ids[g] = row, gid_of_slot[slot] = g
//   4. assign    group_ids[row] = gid_of_slot[slot(key_row)]
// Aggregates are per-group accumulators updated with L2 atomics (red.global), finalised by a small per-group kernel
// (sticky-null sums, +INF/NULL-initialised min/max, f64 averages).
#include <type_traits>

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

template <typename Src, typename Slot, bool INSERT>
int number_groups(rfb_ctx_t *ctx, Src src, Slot slot, i64 len, i64 slots, u64 *first_row, i64 *gid_of_slot, void *tile_work,
                  i64 *mm, i64 *group_ids, i64 *first_ids, i64 *groups) {
    const int grid = rfb_grid_for(ctx, len, THREADS * 4, BLOCKS_PER_SM);
    if constexpr (INSERT) {
        k_claim<Src, Slot, INSERT><<<grid, THREADS, 0, ctx->stream>>>(src, slot, 0, len, first_row);
        RFB_CHECK_LAUNCH(ctx);
    } else {
        // direct addressing: claims run over a growing row prefix and stop as soon as EVERY slot of the key range has its first
        // row (low-cardinality keys: all of them show up within the first few thousand rows; claiming over the whole column
        // would only hammer the same few L2 lines — 100 keys cost 1.1 ms per 1e7 rows that way).  A range with unused
        // slots never completes and ends up claiming over all rows, as before.
        i64 r0 = 0, r1 = 32 * slots > 65536 ? 32 * slots : 65536;
        while (true) {
            if (r1 > len) r1 = len;
            k_claim<Src, Slot, INSERT><<<rfb_grid_for(ctx, r1 - r0, THREADS * 4, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(src, slot, r0, r1, first_row);
            RFB_CHECK_LAUNCH(ctx);
            if (r1 == len) break;
            RFB_CUDA(cudaMemsetAsync(mm + 4, 0, 8, ctx->stream));
            k_claimed_count<<<rfb_grid_for(ctx, slots, THREADS, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(first_row, slots, mm);
            RFB_CHECK_LAUNCH(ctx);
            i64 claimed = 0;
            int rc2 = d2h_sync(ctx, &claimed, mm + 4, 8);
            if (rc2) return rc2;
            if (claimed == slots) break;
            r0 = r1;
            r1 = r1 * 4;
        }
    }
    k_max_first<<<rfb_grid_for(ctx, slots, THREADS, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(first_row, slots, mm);
    RFB_CHECK_LAUNCH(ctx);
    i64 limit = 0;
    int rc = d2h_sync(ctx, &limit, mm + 2, 8);
    if (rc) return rc;
    const i64 tiles = (limit + scan::RowTile<NUM_J>::TILE - 1) / scan::RowTile<NUM_J>::TILE;
    scan::TileCtl ctl;
    rc = scan::prepare_tiles(ctx, tile_work, tiles, ctx->h_count, &ctl);
    if (rc) return rc;
    k_number<Src, Slot><<<(unsigned)tiles, THREADS, 0, ctx->stream>>>(src, slot, limit, first_row, gid_of_slot, first_ids, ctl);
    RFB_CHECK_LAUNCH(ctx);
    if (group_ids) {
        bool done = false;
        if constexpr (!INSERT && std::is_same<Src, KeySrc>::value && std::is_same<Slot, DenseSlot>::value) {
            if (!src.filter && aligned16(src.keys) && aligned16(group_ids)) {
                k_assign_dense_vec<<<rfb_grid_for(ctx, len, THREADS * 8, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(src.keys, slot.min, len, gid_of_slot, group_ids);
                done = true;
            }
        }
        if (!done) k_assign<Src, Slot><<<grid, THREADS, 0, ctx->stream>>>(src, slot, len, gid_of_slot, group_ids);
        RFB_CHECK_LAUNCH(ctx);
    }
    RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    *groups = *(volatile i64 *)ctx->h_count;
    return RFB_OK;
}

}  // namespace

extern "C" int rfb_group_i64_dev(rfb_ctx_t *ctx, const int64_t *keys, const int64_t *filter, int64_t len,
                                 int64_t *group_ids, int64_t *first_ids, rfb_group_info_t *info) {
    RFB_ARG(ctx && info && len >= 0 && ((keys && first_ids) || len == 0), "rfb_group_i64_dev");
    memset(info, 0, sizeof(*info));
    if (len == 0) {  // core/index.c:408-409: empty scope -> dense path with zero groups
        info->min = info->max = NULL_I64;
        info->dense = 1;
        info->index_type = RFB_INDEX_SHIFT;
        return RFB_OK;
    }
    KeySrc src{keys, filter};
    i64 *mm = (i64 *)((char *)ctx->d_scratch + 32768);  // {min, max, limit}
    k_scope_init<<<1, 1, 0, ctx->stream>>>(mm);
    RFB_CHECK_LAUNCH(ctx);
    if (!filter && aligned16(keys)) k_scope_vec<<<rfb_grid_for(ctx, len, THREADS * 8, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(keys, len, mm);
    else k_scope<KeySrc><<<rfb_grid_for(ctx, len, THREADS * 4, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(src, len, mm);
    RFB_CHECK_LAUNCH(ctx);
    i64 h[2];
    int rc = d2h_sync(ctx, h, mm, 16);
    if (rc) return rc;
    info->min = h[0];
    info->max = h[1];
    info->range = (i64)((u64)h[1] - (u64)h[0] + 1);
    const i64 tiles_max = (len + scan::RowTile<NUM_J>::TILE - 1) / scan::RowTile<NUM_J>::TILE;
    i64 groups = 0;
    if (info->range > 0 && info->range <= len) {
        // dense: first_row[range] | gid_of_slot[range] | tile states
        const i64 range = info->range;
        const size_t b1 = align256((size_t)range * 8);
        void *w;
        rc = rfb_ensure_work(ctx, 2 * b1 + scan::tiles_bytes(tiles_max), &w);
        if (rc) return rc;
        u64 *first_row = (u64 *)w;
        i64 *gid_of_slot = (i64 *)((char *)w + b1);
        RFB_CUDA(cudaMemsetAsync(first_row, 0xFF, (size_t)range * 8, ctx->stream));
        rc = number_groups<KeySrc, DenseSlot, false>(ctx, src, DenseSlot{info->min}, len, range, first_row, gid_of_slot, (char *)w + 2 * b1,
                                             mm, group_ids, first_ids, &groups);
        if (rc) return rc;
        info->dense = 1;
        info->index_type = range <= RFB_INDEX_SCOPE_LIMIT ? RFB_INDEX_SHIFT : RFB_INDEX_IDS;  // core/index.c:2063
    } else {
        // sparse: tk[cap+1] | first_row[cap+1] | gid_of_slot[cap+1] | tile states.  cap = power of two >= 2*len
        i64 cap = 1024;
        while (cap < 2 * len) cap <<= 1;
        const size_t b1 = align256((size_t)(cap + 1) * 8);
        void *w;
        rc = rfb_ensure_work(ctx, 3 * b1 + scan::tiles_bytes(tiles_max), &w);
        if (rc) return rc;
        i64 *tk = (i64 *)w;
        u64 *first_row = (u64 *)((char *)w + b1);
        i64 *gid_of_slot = (i64 *)((char *)w + 2 * b1);
        rc = fill<i64>(ctx, tk, cap + 1, NULL_I64);   // EMPTY marker
        if (rc) return rc;
        RFB_CUDA(cudaMemsetAsync(first_row, 0xFF, (size_t)(cap + 1) * 8, ctx->stream));
        HashSlot hs{tk, (u64)(cap - 1), cap};
        rc = number_groups<KeySrc, HashSlot, true>(ctx, src, hs, len, cap + 1, first_row, gid_of_slot, (char *)w + 3 * b1, mm, group_ids,
                                           first_ids, &groups);
        if (rc) return rc;
        info->dense = 0;
        info->index_type = RFB_INDEX_IDS;
    }
    info->groups = groups;
    return RFB_OK;
}

// ------------------------------------------------------------------ distinct keys (dense domain)

namespace {
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM) k_mark_keys(const i64 *__restrict__ keys, i64 n, i64 kmin, u8 *mark) {
    for (i64 i = (i64)blockIdx.x * THREADS + threadIdx.x; i < n; i += (i64)gridDim.x * THREADS) {
        const i64 s = (i64)((u64)ld_stream(keys + i) - (u64)kmin);
        if (!__ldg(mark + s)) mark[s] = 1;            // plain store: every writer writes the same byte
    }
}
__global__ void __launch_bounds__(THREADS) k_add_const(i64 *p, i64 n, i64 c) {
    for (i64 i = (i64)blockIdx.x * THREADS + threadIdx.x; i < n; i += (i64)gridDim.x * THREADS) p[i] = (i64)((u64)p[i] + (u64)c);
}
}  // namespace

extern "C" int rfb_distinct_i64_dev(rfb_ctx_t *ctx, const int64_t *keys, int64_t n, int64_t *out, int64_t *count) {
    RFB_ARG(ctx && count && n >= 0 && ((keys && out) || n == 0), "rfb_distinct_i64_dev");
    *count = 0;
    if (n == 0) return RFB_OK;
    i64 *mm = (i64 *)((char *)ctx->d_scratch + 32768);
    k_scope_init<<<1, 1, 0, ctx->stream>>>(mm);
    RFB_CHECK_LAUNCH(ctx);
    if (aligned16(keys)) k_scope_vec<<<rfb_grid_for(ctx, n, THREADS * 8, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(keys, n, mm);
    else k_scope<KeySrc><<<rfb_grid_for(ctx, n, THREADS * 4, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(KeySrc{keys, nullptr}, n, mm);
    RFB_CHECK_LAUNCH(ctx);
    i64 h[2];
    int rc = d2h_sync(ctx, h, mm, 16);
    if (rc) return rc;
    const i64 range = (i64)((u64)h[1] - (u64)h[0] + 1);
    // index_distinct_i64's direct-addressing branch (core/index.c:558): range <= len or range <= MAX_RANGE (2^20)
    if (range <= 0 || !(range <= n || range <= (1ll << 20))) {
        rfb_set_error("distinct: key range %lld is not dense (the reference's hash branch emits its table's slot order)", (long long)range);
        return RFB_ERR_ARG;
    }
    void *aux;
    rc = rfb_ensure_aux(ctx, (size_t)range, &aux);
    if (rc) return rc;
    u8 *mark = (u8 *)aux;
    RFB_CUDA(cudaMemsetAsync(mark, 0, (size_t)range, ctx->stream));
    k_mark_keys<<<rfb_grid_for(ctx, n, THREADS * 4, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(keys, n, h[0], mark);
    RFB_CHECK_LAUNCH(ctx);
    rc = rfb_where_dev(ctx, mark, range, out, count);       // ascending slots of the keys that occur ...
    if (rc) return rc;
    if (*count > 0 && h[0] != 0) {                            // ... + min = the keys
        k_add_const<<<rfb_grid_for(ctx, *count, THREADS * 4, 8), THREADS, 0, ctx->stream>>>(out, *count, h[0]);
        RFB_CHECK_LAUNCH(ctx);
    }
    return RFB_OK;
}

// ------------------------------------------------------------------ multi-key grouping: perfect-hash key fusion

namespace {
constexpr int MAX_KEY_COLS = 8;

// ---- row-hash path (core/index.c:2556-2729 hashes every row with hash_index_u64 and groups equal tuples through
// per-partition open-addressing tables; the device keeps ONE table of representative rows: a slot is claimed with a CAS on
// the row id and a probe compares the full key tuple of the probing row with the slot's representative)
struct RowSrc {                      // position i of the (filtered) row sequence -> row id (the "key" of the numbering passes)
    const i64 *filter;
    __device__ __forceinline__ i64 operator()(i64 i) const { return filter ? ld_stream(filter + i) : i; }
};
struct TupleSlot {
    const i64 *col[MAX_KEY_COLS];
    int ncols;
    i64 *rep;       // [cap] representative row of each slot, NULL_I64 = empty (row ids are >= 0)
    u64 mask;
    __device__ __forceinline__ u64 hash(i64 row) const {
        u64 h = 0x9E3779B97F4A7C15ULL;
        for (int c = 0; c < ncols; c++) h = mix64(h ^ (u64)__ldg(col[c] + row)) + 0x9E3779B97F4A7C15ULL;
        return h;
    }
    __device__ __forceinline__ bool same(i64 a, i64 b) const {
        if (a == b) return true;
        for (int c = 0; c < ncols; c++)
            if (__ldg(col[c] + a) != __ldg(col[c] + b)) return false;
        return true;
    }
    __device__ __forceinline__ i64 insert(i64 row) const {
        u64 s = hash(row) & mask;
        while (true) {
            i64 cur = (i64)scan::ld_relaxed((const u64 *)&rep[s]);
            if (cur == NULL_I64) {
                cur = (i64)atomicCAS((unsigned long long *)&rep[s], (unsigned long long)NULL_I64, (unsigned long long)row);
                if (cur == NULL_I64) return (i64)s;
            }
            if (same(cur, row)) return (i64)s;
            s = (s + 1) & mask;
        }
    }
    __device__ __forceinline__ i64 operator()(i64 row) const {   // lookup of a tuple known to be present
        u64 s = hash(row) & mask;
        while (!same(__ldcg(&rep[s]), row)) s = (s + 1) & mask;
        return (i64)s;
    }
};

struct FuseSpec {
    const i64 *col[MAX_KEY_COLS];
    i64 min[MAX_KEY_COLS];
    i64 stride[MAX_KEY_COLS];
    int ncols;
};
// fused[i] = sum_c (col_c[row_i] - min_c) * stride_c : a bijection from key tuples to [0, prod ranges)
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM) k_fuse_keys(FuseSpec f, const i64 *__restrict__ filter, i64 n, i64 *__restrict__ fused) {
    for (i64 i = (i64)blockIdx.x * THREADS + threadIdx.x; i < n; i += (i64)gridDim.x * THREADS) {
        const i64 row = filter ? ld_stream(filter + i) : i;
        u64 k = 0;
        for (int c = 0; c < f.ncols; c++) {
            const i64 v = filter ? __ldg(f.col[c] + row) : ld_stream(f.col[c] + row);
            k += ((u64)v - (u64)f.min[c]) * (u64)f.stride[c];
        }
        fused[i] = (i64)k;
    }
}
}  // namespace

extern "C" int rfb_group_keys_i64_dev(rfb_ctx_t *ctx, int ncols, const int64_t *const *cols, const int64_t *filter, int64_t len,
                                      int64_t *group_ids, int64_t *first_ids, rfb_group_info_t *info) {
    RFB_ARG(ctx && info && cols && ncols >= 1 && ncols <= MAX_KEY_COLS && len >= 0 && (first_ids || len == 0), "rfb_group_keys_i64_dev");
    for (int c = 0; c < ncols; c++) RFB_ARG(cols[c] || len == 0, "rfb_group_keys_i64_dev: key column");
    if (ncols == 1 || len == 0) return rfb_group_i64_dev(ctx, len ? cols[0] : nullptr, filter, len, group_ids, first_ids, info);
    FuseSpec f;
    f.ncols = ncols;
    i64 *mm = (i64 *)((char *)ctx->d_scratch + 32768);
    unsigned __int128 space = 1;
    bool hashed = false;
    i64 range[MAX_KEY_COLS];
    for (int c = 0; c < ncols; c++) {   // per-column scope (core/index.c:2308-2340 does the same before fusing)
        KeySrc src{cols[c], filter};
        k_scope_init<<<1, 1, 0, ctx->stream>>>(mm);
        RFB_CHECK_LAUNCH(ctx);
        k_scope<KeySrc><<<rfb_grid_for(ctx, len, THREADS * 4, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(src, len, mm);
        RFB_CHECK_LAUNCH(ctx);
        i64 h[2];
        int rc = d2h_sync(ctx, h, mm, 16);
        if (rc) return rc;
        f.col[c] = cols[c];
        f.min[c] = h[0];
        const unsigned __int128 r = (unsigned __int128)((u64)h[1] - (u64)h[0]) + 1;
        space *= r;
        if (space > ((unsigned __int128)1 << 62)) { hashed = true; space = 1; }   // no perfect hash: group by row hash
        range[c] = (i64)r;
    }
    if (hashed) {
        // rep[cap] | first_row[cap] | gid_of_slot[cap] | tile states.  cap = power of two >= 2*len
        i64 cap = 1024;
        while (cap < 2 * len) cap <<= 1;
        const i64 tiles_max = (len + scan::RowTile<NUM_J>::TILE - 1) / scan::RowTile<NUM_J>::TILE;
        const size_t b1 = align256((size_t)cap * 8);
        void *w;
        int rc = rfb_ensure_work(ctx, 3 * b1 + scan::tiles_bytes(tiles_max), &w);
        if (rc) return rc;
        TupleSlot ts;
        ts.ncols = ncols;
        for (int c = 0; c < ncols; c++) ts.col[c] = cols[c];
        ts.rep = (i64 *)w;
        ts.mask = (u64)(cap - 1);
        u64 *first_row = (u64 *)((char *)w + b1);
        i64 *gid_of_slot = (i64 *)((char *)w + 2 * b1);
        rc = fill<i64>(ctx, ts.rep, cap, NULL_I64);
        if (rc) return rc;
        RFB_CUDA(cudaMemsetAsync(first_row, 0xFF, (size_t)cap * 8, ctx->stream));
        i64 groups = 0;
        rc = number_groups<RowSrc, TupleSlot, true>(ctx, RowSrc{filter}, ts, len, cap, first_row, gid_of_slot, (char *)w + 3 * b1, mm,
                                                    group_ids, first_ids, &groups);
        if (rc) return rc;
        info->groups = groups;
        info->dense = 0;
        info->index_type = RFB_INDEX_IDS;
        info->min = info->max = NULL_I64;
        info->range = 0;
        return RFB_OK;
    }
    i64 stride = 1;
    for (int c = ncols - 1; c >= 0; c--) { f.stride[c] = stride; stride *= range[c]; }
    void *aux;
    int rc = rfb_ensure_aux(ctx, (size_t)len * 8, &aux);
    if (rc) return rc;
    i64 *fused = (i64 *)aux;
    k_fuse_keys<<<rfb_grid_for(ctx, len, THREADS * 4, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(f, filter, len, fused);
    RFB_CHECK_LAUNCH(ctx);
    return rfb_group_i64_dev(ctx, fused, nullptr, len, group_ids, first_ids, info);
}

// ------------------------------------------------------------------ grouped aggregates

namespace {

struct ValRow {
    const i64 *filter;
    __device__ __forceinline__ i64 operator()(i64 i) const { return filter ? ld_stream(filter + i) : i; }
};


__device__ __forceinline__ f64 key_to_f64(u64 k) {  // inverse of f64_sort_key; key 0 = null
    if (k == 0) return null_f64();
    return bits_f64((k & 0x8000000000000000ULL) ? (k & 0x7FFFFFFFFFFFFFFFULL) : ~k);
}

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

// ------------------------------------------------------------------ fused dense group-by: sum + count [+ where]
//
// Three accumulate strategies, picked from the key range found by the scope pass:
//   range <= KP (8192)        CTA-private accumulators in shared memory, merged once per CTA
//   range <= MAX_PARTS * KP   two passes over a key-range PARTITIONED copy of the selected rows: the scatter pass splits
//                             the rows into P = range/KP partitions of (16-bit slot, value) pairs, the accumulate pass
//                             gives every CTA one partition at a time, whose KP accumulators fit shared memory.  36 B/row
//                             of streaming traffic instead of two L2 atomics per row (L2 atomics cap at ~175 G/s on B200,
//                             which is what bounded the 1e5-key config at 12.3 ms per 1e9 rows)
//   otherwise                 device-wide accumulators updated with L2 atomics
// Shared-memory accumulators are 32-bit words (sum low / sum high / count): sm_100a has native 32-bit shared atomics
// (ATOMS.ADD) but implements 64-bit shared adds as a compare-and-swap loop (ATOMS.CAST.SPIN.64).  The 64-bit wrapping
// sum is kept exact by carrying: the returning add on the low word tells the one row that wrapped it to add 1 to the high word.

namespace {

struct Accums {
    u64 *first_row;   // [range]
    u64 *sum;         // [range] wrapping i64 sums of the non-null values
    u64 *cnt;         // [range] rows (nulls included: aggr_count counts rows, core/aggr.c:1336-1342)
    u32 *has_null;    // [range] sticky-null marker for the sum (core/aggr.c:1088)
};

constexpr int KP_LOG = 13, KP = 1 << KP_LOG;   // keys per partition = shared-memory accumulator slots per CTA (96 KB)
constexpr int MAX_PARTS = 256;
constexpr int PT = 512;                        // threads per CTA of the accumulate kernels (2 CTAs per SM)
constexpr int PTILE = PT * 8;                  // rows per tile / work unit: 4 pairs per thread
constexpr int ST = 256;                        // threads per CTA of the scope kernel (4 CTAs per SM)
constexpr int STILE = ST * 8;
#ifndef RFB_SC_T
#define RFB_SC_T 256      /* measured on B200 (1e9 rows, 1e5 i32 keys): 256 x 8 x 4 CTAs 7.17 ms, 512 x 4 x 3 CTAs 7.55 ms */
#define RFB_SC_R 8
#define RFB_SC_CTAS 4
#endif
constexpr int SC_T = RFB_SC_T, SC_R = RFB_SC_R, SC_CTAS = RFB_SC_CTAS, SC_TILE = SC_T * SC_R;   // scatter kernel geometry

// two consecutive elements with one vector load (p must be aligned to 2 * sizeof(T))
template <typename T> __device__ __forceinline__ void ld_pair(const T *p, i64 pair, T &a, T &b) {
    if constexpr (sizeof(T) == 8) {
        const vec16 v = ld_stream16(p + 2 * pair);
        if constexpr (Elem<T>::kind == K_F64) { a = bits_f64(v.lo); b = bits_f64(v.hi); }
        else { a = (T)v.lo; b = (T)v.hi; }
    } else {
        static_assert(sizeof(T) == 4, "pair loads: 4- or 8-byte elements");
        const u64 w = __ldcs((const unsigned long long *)p + pair);
        a = (T)(u32)w;
        b = (T)(u32)(w >> 32);
    }
}
template <typename T> static inline bool pair_aligned(const T *p) { return (((uintptr_t)p) & (2 * sizeof(T) - 1)) == 0; }

template <typename K, typename P, bool HAS_PRED>
struct FusedSrc {
    typedef K key_t;
    const K *keys;
    const P *pred;
    PredRange pr;
    __device__ __forceinline__ bool selected(i64 i) const {
        if constexpr (HAS_PRED) return pred_test(pred_key<P>(ld_stream(pred + i)), pr);
        else return true;
    }
    __device__ __forceinline__ i64 key(i64 i) const { return (i64)ld_stream(keys + i); }
    __device__ __forceinline__ void key_pair(i64 pair, i64 &a, i64 &b) const {
        K x, y;
        ld_pair<K>(keys, pair, x, y);
        a = (i64)x;
        b = (i64)y;
    }
    __device__ __forceinline__ void selected_pair(i64 pair, bool &a, bool &b) const {
        if constexpr (HAS_PRED) {
            P x, y;
            ld_pair<P>(pred, pair, x, y);
            a = pred_test(pred_key<P>(x), pr);
            b = pred_test(pred_key<P>(y), pr);
        } else { a = b = true; }
    }
    bool vec_ok(const i64 *val) const { return pair_aligned(keys) && pair_aligned(val) && (!HAS_PRED || pair_aligned(pred)); }
};

// rows [base, base + R * NT) of (key, value, selected): row of (thread, j, h) = base + 2 * (j * NT + thread) + h, so that
// every load instruction of a warp covers one contiguous, fully used run of bytes
template <int NT, bool WITH_VAL, int R, typename FS>
__device__ __forceinline__ void load_tile(const FS &fs, const i64 *__restrict__ val, i64 base, i64 n, bool vec, i64 (&k)[R], i64 (&v)[R], bool (&sel)[R]) {
    static_assert(R % 2 == 0, "rows per thread come in pairs");
    if (vec && base + R * NT <= n) {
        const i64 pbase = base >> 1;
#pragma unroll
        for (int j = 0; j < R / 2; j++) {
            const i64 pair = pbase + j * NT + threadIdx.x;
            fs.key_pair(pair, k[2 * j], k[2 * j + 1]);
            if constexpr (WITH_VAL) ld_pair<i64>(val, pair, v[2 * j], v[2 * j + 1]);
            fs.selected_pair(pair, sel[2 * j], sel[2 * j + 1]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < R; j++) {
            const i64 r = base + 2 * ((j >> 1) * NT + threadIdx.x) + (j & 1);
            sel[j] = r < n && fs.selected(r);
            k[j] = r < n ? fs.key(r) : 0;
            if constexpr (WITH_VAL) v[j] = r < n ? ld_stream(val + r) : 0;
        }
    }
}

// merge into the device-wide accumulators and clear; slot0 = device-wide slot of local slot 0 (a local slot that
// received rows always maps inside [0, range))
__device__ __forceinline__ void sacc_flush(const SAcc &a, int slots, i64 slot0, const Accums &ga) {
    for (int s = threadIdx.x; s < slots; s += blockDim.x) {
        const u32 c = a.cnt[s];
        if (!c) continue;
        const i64 g = slot0 + s;
        const u64 sum = ((u64)a.hi[s] << 32) | a.lo[s];
        if (sum) atomicAdd((unsigned long long *)ga.sum + g, (unsigned long long)sum);
        atomicAdd((unsigned long long *)ga.cnt + g, (unsigned long long)(c & ~NULL_FLAG));
        if (c & NULL_FLAG) ga.has_null[g] = 1u;
        a.lo[s] = 0; a.hi[s] = 0; a.cnt[s] = 0;
    }
}

// ---- scope: min/max of the selected keys
constexpr int MM_WORDS = 8;   // mm[0..7] = {min, max, limit, nonempty, claimed, -, -, -}
__global__ void k_fused_scope_init(i64 *mm) {
    for (int i = threadIdx.x; i < MM_WORDS; i += blockDim.x) mm[i] = i == 0 ? RFB_INF_I64 : (i == 1 ? NULL_I64 : 0);
}

template <typename FS>
__global__ void __launch_bounds__(ST, 4) k_fused_scope(FS fs, i64 n, bool vec, i64 *mm) {
    __shared__ i64 red[32];
    i64 lo = RFB_INF_I64, hi = NULL_I64;
    const i64 tiles = (n + STILE - 1) / STILE;
    for (i64 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        i64 k[8], v[8];
        bool sel[8];
        load_tile<ST, false, 8>(fs, nullptr, tile * STILE, n, vec, k, v, sel);
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (sel[j]) { lo = k[j] < lo ? k[j] : lo; hi = k[j] > hi ? k[j] : hi; }
    }
    struct Mn { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b < a ? b : a; } };
    struct Mx { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b > a ? b : a; } };
    lo = block_reduce<i64>(lo, Mn(), RFB_INF_I64, red);
    hi = block_reduce<i64>(hi, Mx(), NULL_I64, red);
    if (threadIdx.x == 0) {
        atomicMin((long long *)&mm[0], (long long)lo);
        atomicMax((long long *)&mm[1], (long long)hi);
    }
}

// first-row claims over a row prefix [r0, r1) only: in the accumulate pass a claim would cost one L2 read per row although
// it can only change anything while a key has not been seen yet; the host extends the prefix until every non-empty slot
// has been claimed (one short pass for any column whose keys all occur early, e.g. uniform keys).
template <typename FS>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM) k_fused_claim(FS fs, i64 r0, i64 r1, i64 kmin, u64 *first_row) {
    for (i64 i = r0 + (i64)blockIdx.x * THREADS + threadIdx.x; i < r1; i += (i64)gridDim.x * THREADS)
        if (fs.selected(i)) claim_first(first_row, (i64)((u64)fs.key(i) - (u64)kmin), i);
}

// mm[3] = slots with rows, mm[4] = slots whose first row is known
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM) k_slot_census(const u64 *first_row, const u64 *cnt, i64 range, i64 *mm) {
    __shared__ i64 red[32];
    i64 nonempty = 0, claimed = 0;
    for (i64 s = (i64)blockIdx.x * THREADS + threadIdx.x; s < range; s += (i64)gridDim.x * THREADS) {
        nonempty += cnt[s] != 0;
        claimed += first_row[s] != NO_ROW;
    }
    nonempty = block_reduce<i64>(nonempty, OpAddWrap(), 0, red);
    claimed = block_reduce<i64>(claimed, OpAddWrap(), 0, red);
    if (threadIdx.x == 0) { atomicAdd((unsigned long long *)&mm[3], (unsigned long long)nonempty); atomicAdd((unsigned long long *)&mm[4], (unsigned long long)claimed); }
}

// ---- accumulate, strategy 3: device-wide accumulators, two L2 atomics per row
template <typename FS>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_fused_accum_l2(FS fs, const i64 *__restrict__ val, i64 n, i64 kmin, Accums a) {
    constexpr int U = 4;
    const i64 stride = (i64)gridDim.x * THREADS;
    auto one = [&](i64 k, i64 v, bool sel) {
        if (!sel) return;
        const i64 s = (i64)((u64)k - (u64)kmin);
        if (v == NULL_I64) a.has_null[s] = 1u; else atomicAdd((unsigned long long *)a.sum + s, (unsigned long long)v);
        atomicAdd((unsigned long long *)a.cnt + s, 1ULL);
    };
    i64 i = (i64)blockIdx.x * THREADS + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        i64 k[U], v[U];
        bool sel[U];
#pragma unroll
        for (int j = 0; j < U; j++) { k[j] = fs.key(i + j * stride); v[j] = ld_stream(val + i + j * stride); sel[j] = fs.selected(i + j * stride); }
#pragma unroll
        for (int j = 0; j < U; j++) one(k[j], v[j], sel[j]);
    }
    for (; i < n; i += stride) one(fs.key(i), ld_stream(val + i), fs.selected(i));
}

// ---- accumulate, strategy 1: range <= KP, CTA-private shared-memory accumulators (3 x 4 B x range, dynamic)
template <typename FS>
__global__ void __launch_bounds__(PT, 2)
k_fused_accum_smem(FS fs, const i64 *__restrict__ val, i64 n, bool vec, i64 kmin, int range, Accums ga) {
    extern __shared__ u32 s_acc[];
    const SAcc a{s_acc, s_acc + range, s_acc + 2 * range};
    sacc_zero(a, range);
    __syncthreads();
    const i64 tiles = (n + PTILE - 1) / PTILE;
    for (i64 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        i64 k[8], v[8];
        bool sel[8];
        load_tile<PT, true, 8>(fs, val, tile * PTILE, n, vec, k, v, sel);
#pragma unroll
        for (int j = 0; j < 8; j++) sacc_add_warp(a, sel[j], (u32)((u64)k[j] - (u64)kmin), v[j]);
    }
    __syncthreads();
    sacc_flush(a, range, 0, ga);
}

// strategy 1 without a scope pass: slot = key mod KP.  Any KP consecutive integers have distinct residues, so when the keys
// turn out to span fewer than KP values (the kernel finds min/max on the way, a row sample made it likely) the residue IS a
// perfect hash, and k_mod_remap afterwards moves slot (key mod KP) to slot (key - min).  Saves the 4-8 B/row scope pass.
template <typename FS>
__global__ void __launch_bounds__(PT, 2)
k_fused_accum_mod(FS fs, const i64 *__restrict__ val, i64 n, bool vec, Accums gmod, i64 *mm) {
    extern __shared__ u32 s_acc[];
    __shared__ i64 red[32];
    const SAcc a{s_acc, s_acc + KP, s_acc + 2 * KP};
    sacc_zero(a, KP);
    __syncthreads();
    typedef typename FS::key_t KT;
    KT lo = sizeof(KT) == 4 ? (KT)0x7FFFFFFF : (KT)RFB_INF_I64, hi = sizeof(KT) == 4 ? (KT)NULL_I32 : (KT)NULL_I64;
    const i64 tiles = (n + PTILE - 1) / PTILE;
    for (i64 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        i64 k[8], v[8];
        bool sel[8];
        load_tile<PT, true, 8>(fs, val, tile * PTILE, n, vec, k, v, sel);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (sel[j]) {
                lo = (KT)k[j] < lo ? (KT)k[j] : lo;
                hi = (KT)k[j] > hi ? (KT)k[j] : hi;
            }
            sacc_add_warp(a, sel[j], (u32)((u64)k[j] & (KP - 1)), v[j]);
        }
    }
    __syncthreads();
    sacc_flush(a, KP, 0, gmod);
    struct Mn { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b < a ? b : a; } };
    struct Mx { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b > a ? b : a; } };
    const i64 lo64 = block_reduce<i64>((i64)lo, Mn(), RFB_INF_I64, red);
    const i64 hi64 = block_reduce<i64>((i64)hi, Mx(), NULL_I64, red);
    if (threadIdx.x == 0) {
        atomicMin((long long *)&mm[0], (long long)lo64);
        atomicMax((long long *)&mm[1], (long long)hi64);
    }
}

__global__ void __launch_bounds__(THREADS) k_mod_remap(Accums gmod, i64 kmin, i64 range, Accums a) {
    for (i64 s = (i64)blockIdx.x * THREADS + threadIdx.x; s < range; s += (i64)gridDim.x * THREADS) {
        const u64 m = ((u64)kmin + (u64)s) & (KP - 1);
        a.sum[s] = gmod.sum[m];
        a.cnt[s] = gmod.cnt[m];
        a.has_null[s] = gmod.has_null[m];
    }
}

// ---- accumulate, strategy 2: partition, then accumulate per partition
//
// A row's partition is its ABSOLUTE key bucket (key >> KP_LOG) mod 256 and its slot the low KP_LOG key bits, so the scatter
// pass needs no key bounds: it computes min/max itself, and the partitioning is valid iff the keys turn out to span at most
// 256 buckets (checked afterwards; a row sample decides beforehand whether it is worth trying).  Partition sizes are not
// known in advance either: partitioned rows live in blocks of PB rows handed out on demand.  cursor[b] counts the rows of
// bucket b; the tile whose run contains the first row of a block allocates it (one atomic on a block counter) and publishes
// it in the block table bt[b][i]; tiles that write into a block they did not allocate wait for that word.  The allocator
// has already passed its cursor atomic and publishes before it waits for anything itself, so the wait cannot deadlock.
constexpr int PB_LOG = 16, PB = 1 << PB_LOG;   // rows per block; a multiple of PTILE, so an accumulate unit never straddles blocks

struct PartStore {
    u32 *cursor;       // [MAX_PARTS] rows per bucket
    u32 *next_block;   // blocks handed out so far
    u32 *bt;           // [MAX_PARTS][bt_stride] physical block + 1 (0 = not yet allocated)
    u32 bt_stride;
    u64 *val;          // [blocks * PB]
    u16 *slot;         // [blocks * PB]
};

__device__ __forceinline__ u32 ld_relaxed_u32(const u32 *p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(u32 *p, u32 v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

struct ScatterSmem {
    u64 val[SC_TILE];
    u32 pk[SC_TILE];               // bucket << KP_LOG | slot
    uint4 desc[MAX_PARTS];       // per bucket: {x: physical - local offset before the block boundary, y: first local index past it, z: offset after it, w: local base}
    u32 cnt[MAX_PARTS];
    u32 wtot[MAX_PARTS / 32];
    u32 total;
    i64 red[32];
};

// scatter pass: every tile orders its selected rows by bucket in shared memory (a row's rank inside its bucket is what the
// returning shared atomic on the bucket's counter hands back), reserves its run in every bucket with one global atomic per
// bucket, and writes the runs out contiguously.  Row order inside a bucket is not preserved (integer sums and counts do
// not depend on it; first rows are claimed from the source columns).
template <typename FS>
__global__ void __launch_bounds__(SC_T, SC_CTAS)
k_part_scatter(FS fs, const i64 *__restrict__ val, i64 n, bool vec, PartStore ps, i64 *mm) {
    static_assert(SC_T >= MAX_PARTS, "one thread per bucket counter");
    __shared__ ScatterSmem sm;
    const int tid = threadIdx.x, lane = tid & 31;
    const i64 tiles = (n + SC_TILE - 1) / SC_TILE;
    typedef typename FS::key_t KT;   // running min/max in the key column's own width (register pressure)
    KT lo = sizeof(KT) == 4 ? (KT)0x7FFFFFFF : (KT)RFB_INF_I64, hi = sizeof(KT) == 4 ? (KT)NULL_I32 : (KT)NULL_I64;
    if (tid < MAX_PARTS) sm.cnt[tid] = 0;
    __syncthreads();
    for (i64 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        i64 k[SC_R], v[SC_R];
        bool sel[SC_R];
        load_tile<SC_T, true, SC_R>(fs, val, tile * SC_TILE, n, vec, k, v, sel);
        u32 pk[SC_R], pos[SC_R];
#pragma unroll
        for (int j = 0; j < SC_R; j++) {
            if (sel[j]) { lo = (KT)k[j] < lo ? (KT)k[j] : lo; hi = (KT)k[j] > hi ? (KT)k[j] : hi; }
            pk[j] = sel[j] ? (u32)((u64)k[j] & ((1u << (KP_LOG + 8)) - 1u)) : 0xFFFFFFFFu;
            const u32 part = sel[j] ? pk[j] >> KP_LOG : 0xFFFFFFFFu;
            const u32 p0 = __shfl_sync(0xffffffffu, part, 0);
            if (__all_sync(0xffffffffu, part == p0)) {      // the whole warp step goes to one bucket: one atomic
                u32 b = 0;
                if (lane == 0 && sel[j]) b = atomicAdd(&sm.cnt[part], 32u);
                pos[j] = __shfl_sync(0xffffffffu, b, 0) + lane;
            } else if (sel[j]) pos[j] = atomicAdd(&sm.cnt[part], 1u);
        }
        __syncthreads();
        u32 c = 0, incl = 0, start = 0, phys0 = 0, phys1 = 0;
        if (tid < MAX_PARTS) {
            c = sm.cnt[tid];
            incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            if (lane == 31) sm.wtot[tid >> 5] = incl;
            // reserve [start, start + c) of bucket `tid`, allocate the block(s) that begin inside the run, look up the two
            // blocks the run can touch
            if (c) {
                start = atomicAdd(&ps.cursor[tid], c);
                const u32 b0 = start >> PB_LOG, b1 = (start + c - 1) >> PB_LOG;
                u32 *row = ps.bt + (size_t)tid * ps.bt_stride;
                if ((start & (PB - 1)) == 0) { phys0 = atomicAdd(ps.next_block, 1u) + 1; st_relaxed_u32(row + b0, phys0); }
                if (b1 != b0) { phys1 = atomicAdd(ps.next_block, 1u) + 1; st_relaxed_u32(row + b1, phys1); }
                while (!phys0) phys0 = ld_relaxed_u32(row + b0);
                if (b1 == b0) phys1 = phys0;
            }
        }
        __syncthreads();
        if (tid < MAX_PARTS) {
            u32 before = 0;
            for (int w = 0; w < (tid >> 5); w++) before += sm.wtot[w];
            const u32 lbase = before + incl - c;
            const u32 in_block = start & (PB - 1), room = PB - in_block;           // rows left in the first block
            uint4 d;
            d.x = (phys0 - 1) * (u32)PB + in_block - lbase;
            d.y = lbase + room;
            d.z = (phys1 - 1) * (u32)PB - (lbase + room);
            d.w = lbase;
            sm.desc[tid] = d;
            sm.cnt[tid] = 0;                                   // for the next tile (this tile's counts live in registers now)
            if (tid == MAX_PARTS - 1) sm.total = before + incl;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SC_R; j++) {
            if (!sel[j]) continue;
            const u32 q = sm.desc[pk[j] >> KP_LOG].w + pos[j];
            sm.val[q] = (u64)v[j];
            sm.pk[q] = pk[j];
        }
        __syncthreads();
        const u32 total = sm.total;
#pragma unroll
        for (int step = 0; step < SC_R / 2; step++) {      // 2 independent elements per step: the shared-memory lookups overlap
            u32 w[2];
            u64 vv[2];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const u32 q = (step * 2 + j) * SC_T + tid;
                if (q < total) { w[j] = sm.pk[q]; vv[j] = sm.val[q]; }
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const u32 q = (step * 2 + j) * SC_T + tid;
                if (q < total) {
                    const uint4 dd = sm.desc[w[j] >> KP_LOG];
                    const u32 g = q + (q < dd.y ? dd.x : dd.z);
                    ps.val[g] = vv[j];
                    ps.slot[g] = (u16)(w[j] & (KP - 1));
                }
            }
        }
        __syncthreads();
    }
    struct Mn { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b < a ? b : a; } };
    struct Mx { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b > a ? b : a; } };
    // (nothing selected: lo > hi in either width, which is all the host tests)
    const i64 lo64 = block_reduce<i64>((i64)lo, Mn(), RFB_INF_I64, sm.red);
    const i64 hi64 = block_reduce<i64>((i64)hi, Mx(), NULL_I64, sm.red);
    if (tid == 0) {
        atomicMin((long long *)&mm[0], (long long)lo64);
        atomicMax((long long *)&mm[1], (long long)hi64);
    }
}

// accumulate pass: partition p = bucket ((kbase >> KP_LOG) + p) mod 256.  CTA b takes the flattened work units
// [b*U/G, (b+1)*U/G) (unit = PTILE consecutive rows of one partition), keeps the current partition's KP accumulators in
// shared memory and merges them into the device-wide arrays when the partition changes
__global__ void __launch_bounds__(PT, 2)
k_part_accum(PartStore ps, int P, i64 kbase, i64 kmin, Accums ga) {
    extern __shared__ u32 s_acc[];
    __shared__ u32 s_cnt[MAX_PARTS], s_ubase[MAX_PARTS + 1], s_wtot[MAX_PARTS / 32];
    const SAcc a{s_acc, s_acc + KP, s_acc + 2 * KP};
    sacc_zero(a, KP);
    const int tid = threadIdx.x, lane = tid & 31;
    const u32 bucket0 = (u32)(((u64)kbase >> KP_LOG) & 255u);
    u32 c = 0, units = 0, incl = 0;
    if (tid < MAX_PARTS) {
        c = tid < P ? ps.cursor[(bucket0 + tid) & 255u] : 0;
        s_cnt[tid] = c;
        units = (c + PTILE - 1) / PTILE;
        incl = units;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_wtot[tid >> 5] = incl;
    }
    __syncthreads();
    if (tid < MAX_PARTS) {
        u32 before = 0;
        for (int w = 0; w < (tid >> 5); w++) before += s_wtot[w];
        s_ubase[tid + 1] = before + incl;
        if (tid == 0) s_ubase[0] = 0;
    }
    __syncthreads();
    const u32 U = s_ubase[P];
    const u32 u0 = (u32)((u64)blockIdx.x * U / gridDim.x), u1 = (u32)((u64)(blockIdx.x + 1) * U / gridDim.x);
    int p = 0;
    while (p + 1 < P && s_ubase[p + 1] <= u0) p++;
    bool dirty = false;
    u32 have_blk = 0xFFFFFFFFu, phys = 0;   // block-table entry of the (partition, block) the previous unit was in
    for (u32 u = u0; u < u1; u++) {
        if (s_ubase[p + 1] <= u) {
            __syncthreads();
            if (dirty) sacc_flush(a, KP, (i64)((u64)kbase + (u64)p * KP - (u64)kmin), ga);
            __syncthreads();
            dirty = false;
            while (s_ubase[p + 1] <= u) p++;
            have_blk = 0xFFFFFFFFu;
        }
        const u32 r0 = (u - s_ubase[p]) * PTILE, cnt = s_cnt[p];
        const u32 rows = cnt - r0 < (u32)PTILE ? cnt - r0 : (u32)PTILE;
        const u32 bucket = (bucket0 + p) & 255u;
        if ((r0 >> PB_LOG) != have_blk) { have_blk = r0 >> PB_LOG; phys = ps.bt[(size_t)bucket * ps.bt_stride + have_blk] - 1; }
        const u64 base = (u64)phys * PB + (r0 & (PB - 1));
        dirty = true;
        if (rows == PTILE) {
            vec16 vv[4];
            u32 ss[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const u32 q = j * PT + threadIdx.x;
                vv[j] = ld_stream16(ps.val + base + 2 * q);
                ss[j] = __ldcs((const u32 *)(ps.slot + base) + q);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                sacc_add(a, ss[j] & 0xFFFFu, (i64)vv[j].lo);
                sacc_add(a, ss[j] >> 16, (i64)vv[j].hi);
            }
        } else {
            for (u32 r = threadIdx.x; r < rows; r += PT) sacc_add(a, ps.slot[base + r], (i64)ps.val[base + r]);
        }
    }
    __syncthreads();
    if (dirty) sacc_flush(a, KP, (i64)((u64)kbase + (u64)p * KP - (u64)kmin), ga);
}

// ---- the same accumulate pass with the partition data staged by the TMA unit (cp.async.bulk + mbarrier): one CTA per SM,
// 1024 threads, a 3-stage ring of 40 KB units (4096 values + 4096 slots) next to the 96 KB of accumulators.  Thread 0 issues
// the bulk copies two units ahead; the consumers wait on the stage's mbarrier phase and read the unit from shared memory.
// A CTA barrier per unit separates "everyone finished stage s" from "stage s is re-armed".
constexpr int AT = 1024, ASTAGES = 3;
struct __align__(16) AccumStage { u64 val[PTILE]; u16 slot[PTILE]; };

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__global__ void __launch_bounds__(AT, 1)
k_part_accum_tma(PartStore ps, int P, i64 kbase, i64 kmin, Accums ga) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    AccumStage *stage = (AccumStage *)s_dyn;                                   // ASTAGES x 40 KB
    u32 *s_acc = (u32 *)(s_dyn + ASTAGES * sizeof(AccumStage));                // lo[KP] | hi[KP] | cnt[KP]
    __shared__ u64 full[ASTAGES];
    __shared__ u32 s_cnt[MAX_PARTS], s_ubase[MAX_PARTS + 1], s_wtot[MAX_PARTS / 32];
    const SAcc a{s_acc, s_acc + KP, s_acc + 2 * KP};
    sacc_zero(a, KP);
    const int tid = threadIdx.x, lane = tid & 31;
    const u32 bucket0 = (u32)(((u64)kbase >> KP_LOG) & 255u);
    if (tid == 0) {
        for (int s = 0; s < ASTAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    u32 c = 0, units = 0, incl = 0;
    if (tid < MAX_PARTS) {
        c = tid < P ? ps.cursor[(bucket0 + tid) & 255u] : 0;
        s_cnt[tid] = c;
        units = (c + PTILE - 1) / PTILE;
        incl = units;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_wtot[tid >> 5] = incl;
    }
    __syncthreads();
    if (tid < MAX_PARTS) {
        u32 before = 0;
        for (int w = 0; w < (tid >> 5); w++) before += s_wtot[w];
        s_ubase[tid + 1] = before + incl;
        if (tid == 0) s_ubase[0] = 0;
    }
    __syncthreads();
    const u32 U = s_ubase[P];
    const u32 u0 = (u32)((u64)blockIdx.x * U / gridDim.x), u1 = (u32)((u64)(blockIdx.x + 1) * U / gridDim.x);
    // producer state (thread 0 only): partition cursor + cached block-table entry
    int pp = 0;
    u32 p_blk = 0xFFFFFFFFu, p_phys = 0;
    auto issue = [&](u32 u) {
        while (s_ubase[pp + 1] <= u) { pp++; p_blk = 0xFFFFFFFFu; }
        const u32 r0 = (u - s_ubase[pp]) * PTILE, cnt = s_cnt[pp];
        const u32 rows = cnt - r0 < (u32)PTILE ? cnt - r0 : (u32)PTILE, rows_up = (rows + 7u) & ~7u;
        const u32 bucket = (bucket0 + pp) & 255u;
        if ((r0 >> PB_LOG) != p_blk) { p_blk = r0 >> PB_LOG; p_phys = ps.bt[(size_t)bucket * ps.bt_stride + p_blk] - 1; }
        const u64 base = (u64)p_phys * PB + (r0 & (PB - 1));
        const int s = (int)((u - u0) % ASTAGES);
        mbar_expect_tx(&full[s], rows_up * 10u);
        bulk_g2s(stage[s].val, ps.val + base, rows_up * 8u, &full[s]);
        bulk_g2s(stage[s].slot, ps.slot + base, rows_up * 2u, &full[s]);
    };
    if (tid == 0)
        for (u32 u = u0; u < u1 && u < u0 + (ASTAGES - 1); u++) issue(u);
    int p = 0;
    while (p + 1 < P && s_ubase[p + 1] <= u0) p++;
    bool dirty = false;
    for (u32 u = u0; u < u1; u++) {
        const u32 k = u - u0;
        const int s = (int)(k % ASTAGES);
        __syncthreads();                                                       // unit u-1 (stage (s+2)%3) fully consumed
        if (tid == 0 && u + (ASTAGES - 1) < u1) issue(u + (ASTAGES - 1));
        if (s_ubase[p + 1] <= u) {
            if (dirty) sacc_flush(a, KP, (i64)((u64)kbase + (u64)p * KP - (u64)kmin), ga);
            __syncthreads();
            dirty = false;
            while (s_ubase[p + 1] <= u) p++;
        }
        const u32 r0 = (u - s_ubase[p]) * PTILE, cnt = s_cnt[p];
        const u32 rows = cnt - r0 < (u32)PTILE ? cnt - r0 : (u32)PTILE;
        mbar_wait(&full[s], (k / ASTAGES) & 1u);
        dirty = true;
        const AccumStage &st = stage[s];
#pragma unroll
        for (int j = 0; j < PTILE / (2 * AT); j++) {
            const u32 q = j * AT + tid;                                        // pair index
            if (2 * q + 1 < rows) {
                const ulonglong2 vv = *(const ulonglong2 *)&st.val[2 * q];
                const u32 ss = *(const u32 *)&st.slot[2 * q];
                sacc_add(a, ss & 0xFFFFu, (i64)vv.x);
                sacc_add(a, ss >> 16, (i64)vv.y);
            } else if (2 * q < rows) sacc_add(a, st.slot[2 * q], (i64)st.val[2 * q]);
        }
    }
    __syncthreads();
    if (dirty) sacc_flush(a, KP, (i64)((u64)kbase + (u64)p * KP - (u64)kmin), ga);
}

// min/max of the selected keys of the rows [r0, r1): the sample that decides whether the partitioned strategy is tried
template <typename FS>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM) k_fused_scope_rows(FS fs, i64 r0, i64 r1, i64 *mm) {
    __shared__ i64 red[32];
    i64 lo = RFB_INF_I64, hi = NULL_I64;
    for (i64 i = r0 + (i64)blockIdx.x * THREADS + threadIdx.x; i < r1; i += (i64)gridDim.x * THREADS)
        if (fs.selected(i)) { const i64 k = fs.key(i); lo = k < lo ? k : lo; hi = k > hi ? k : hi; }
    struct Mn { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b < a ? b : a; } };
    struct Mx { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b > a ? b : a; } };
    lo = block_reduce<i64>(lo, Mn(), RFB_INF_I64, red);
    hi = block_reduce<i64>(hi, Mx(), NULL_I64, red);
    if (threadIdx.x == 0) {
        atomicMin((long long *)&mm[0], (long long)lo);
        atomicMax((long long *)&mm[1], (long long)hi);
    }
}

template <typename FS>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_fused_emit(FS fs, i64 limit, i64 kmin, Accums a, i64 max_groups, i64 *out_keys, i64 *out_sums, i64 *out_counts, scan::TileCtl ctl) {
    __shared__ scan::TileSmem sm;
    scan::compact_rows<NUM_J>(
        limit, ctl, sm,
        [&](i64 r) { return fs.selected(r) && __ldcg(&a.first_row[(i64)((u64)fs.key(r) - (u64)kmin)]) == (u64)r; },
        [&](i64 r, i64 g) {
            if (g >= max_groups) return;
            const i64 k = fs.key(r), s = (i64)((u64)k - (u64)kmin);
            out_keys[g] = k;
            out_sums[g] = a.has_null[s] ? NULL_I64 : (i64)a.sum[s];
            out_counts[g] = (i64)a.cnt[s];
        });
}

// tuning knobs (environment): RFB_GROUP_STRATEGY = smem | part | l2 forces a strategy where it is applicable;
// RFB_PART_MIN_ROWS = smallest row count that takes the partitioned strategy
int group_strategy_forced() {
    const char *s = getenv("RFB_GROUP_STRATEGY");
    if (!s) return 0;
    return !strcmp(s, "smem") ? 1 : (!strcmp(s, "part") ? 2 : (!strcmp(s, "l2") ? 3 : 0));
}
i64 part_min_rows() {
    const char *s = getenv("RFB_PART_MIN_ROWS");
    return s ? atoll(s) : (1ll << 21);
}

template <typename FS>
int fused_run(rfb_ctx_t *ctx, FS fs, const i64 *val, i64 n, i64 max_groups, i64 *out_keys, i64 *out_sums, i64 *out_counts, i64 *groups) {
    i64 *mm = (i64 *)((char *)ctx->d_scratch + 32768);
    const int forced = group_strategy_forced();
    const bool part_able = n < 0xF0000000ll && (forced == 2 || (forced == 0 && n >= part_min_rows()));   // 32-bit row positions
    const int grid = rfb_grid_for(ctx, n, THREADS * 4, BLOCKS_PER_SM);
    const int sgrid = rfb_grid_for(ctx, n, STILE, 4);
    const bool vec = fs.vec_ok(val);
    const i64 tiles_max = (n + scan::RowTile<NUM_J>::TILE - 1) / scan::RowTile<NUM_J>::TILE;
    auto parts_of = [](i64 kmin, i64 kmax) { return (i64)(((u64)kmax - (u64)(kmin & ~(i64)(KP - 1))) >> KP_LOG) + 1; };
    i64 h[2];
    int rc;
    bool have_scope = false, scattered = false, modded = false;
    void *w = nullptr;
    PartStore ps{};
    Accums gmod{};
    // workspace of the partitioned strategy: accumulators for the largest range it accepts, then the block store
    const i64 max_range = (i64)MAX_PARTS * KP;
    const size_t pb8 = align256((size_t)max_range * 8), pb4 = align256((size_t)max_range * 4);
    const size_t p_acc_bytes = 3 * pb8 + pb4 + scan::tiles_bytes(tiles_max);
    if (part_able) {
        // 1. sample: three row windows; only a key range that needs more than one partition and fits MAX_PARTS is worth a scatter
        k_fused_scope_init<<<1, 32, 0, ctx->stream>>>(mm);
        RFB_CHECK_LAUNCH(ctx);
        const i64 win = 65536;
        const i64 starts[3] = {0, n / 2 > win ? n / 2 : 0, n > win ? n - win : 0};
        for (int s = 0; s < 3; s++) {
            const i64 r0 = starts[s], r1 = r0 + win < n ? r0 + win : n;
            k_fused_scope_rows<FS><<<64, THREADS, 0, ctx->stream>>>(fs, r0, r1, mm);
            RFB_CHECK_LAUNCH(ctx);
        }
        rc = d2h_sync(ctx, h, mm, 16);
        if (rc) return rc;
        const bool try_mod = h[0] <= h[1] && (forced == 0 || forced == 1) && (u64)h[1] - (u64)h[0] < (u64)KP;
        if (try_mod) {
            void *aux;
            rc = rfb_ensure_aux(ctx, (size_t)KP * 20, &aux);
            if (rc) return rc;
            gmod.first_row = nullptr;
            gmod.sum = (u64 *)aux;
            gmod.cnt = gmod.sum + KP;
            gmod.has_null = (u32 *)(gmod.cnt + KP);
            RFB_CUDA(cudaMemsetAsync(aux, 0, (size_t)KP * 20, ctx->stream));
            k_fused_scope_init<<<1, 32, 0, ctx->stream>>>(mm);
            RFB_CHECK_LAUNCH(ctx);
            const i64 ptiles = (n + PTILE - 1) / PTILE;
            const int pgrid = (int)(ptiles < 2ll * ctx->sm_count ? ptiles : 2ll * ctx->sm_count);
            RFB_CUDA(cudaFuncSetAttribute(k_fused_accum_mod<FS>, cudaFuncAttributeMaxDynamicSharedMemorySize, KP * 12));
            k_fused_accum_mod<FS><<<pgrid, PT, KP * 12, ctx->stream>>>(fs, val, n, vec, gmod, mm);
            RFB_CHECK_LAUNCH(ctx);
            have_scope = modded = true;
        }
        const bool try_part = !modded && h[0] <= h[1] && (forced == 2 || (i64)((u64)h[1] - (u64)h[0]) >= KP) && (u64)h[1] - (u64)h[0] < (u64)max_range &&
                              parts_of(h[0], h[1]) <= MAX_PARTS;
        if (try_part) {
            const u32 blocks = (u32)((n + PB - 1) / PB) + MAX_PARTS + 1;
            const u32 bt_stride = (u32)((n + PB - 1) / PB) + 1;
            const size_t bt_bytes = align256((size_t)MAX_PARTS * bt_stride * 4);
            const size_t ctl_bytes = align256((MAX_PARTS + 1) * 4);
            rc = rfb_ensure_work(ctx, p_acc_bytes + ctl_bytes + bt_bytes + (size_t)blocks * PB * 10, &w);
            if (rc) return rc;
            char *pw = (char *)w + p_acc_bytes;
            ps.cursor = (u32 *)pw;
            ps.next_block = ps.cursor + MAX_PARTS;
            ps.bt = (u32 *)(pw + ctl_bytes);
            ps.bt_stride = bt_stride;
            ps.val = (u64 *)(pw + ctl_bytes + bt_bytes);
            ps.slot = (u16 *)(pw + ctl_bytes + bt_bytes + (size_t)blocks * PB * 8);
            RFB_CUDA(cudaMemsetAsync(pw, 0, ctl_bytes + bt_bytes, ctx->stream));
            k_fused_scope_init<<<1, 32, 0, ctx->stream>>>(mm);
            RFB_CHECK_LAUNCH(ctx);
            k_part_scatter<FS><<<rfb_grid_for(ctx, n, SC_TILE, SC_CTAS), SC_T, 0, ctx->stream>>>(fs, val, n, vec, ps, mm);
            RFB_CHECK_LAUNCH(ctx);
            have_scope = scattered = true;
        }
    }
    if (!have_scope) {
        k_fused_scope_init<<<1, 32, 0, ctx->stream>>>(mm);
        RFB_CHECK_LAUNCH(ctx);
        k_fused_scope<FS><<<sgrid, ST, 0, ctx->stream>>>(fs, n, vec, mm);
        RFB_CHECK_LAUNCH(ctx);
    }
    rc = d2h_sync(ctx, h, mm, 16);
    if (rc) return rc;
    if (h[0] > h[1]) { *groups = 0; return RFB_OK; }   // nothing selected
    const i64 kmin = h[0], range = (i64)((u64)h[1] - (u64)h[0] + 1);
    if (range <= 0 || range > (1ll << 28)) {
        rfb_set_error("fused group-by: key range %lld is not a dense domain (use rfb_group_i64_dev + rfb_aggr_dev)", (long long)range);
        return RFB_ERR_ARG;
    }
    const i64 kbase = kmin & ~(i64)(KP - 1);            // floor to a multiple of KP (two's complement)
    const i64 P = parts_of(kmin, h[1]);                 // partitions of KP consecutive keys
    int strategy = 3;
    if (range <= KP && (n >= 65536 || forced == 1)) strategy = 1;
    else if (scattered && P <= MAX_PARTS) strategy = 2;
    if (forced == 3) strategy = 3;

    const size_t b8 = strategy == 2 ? pb8 : align256((size_t)range * 8), b4 = strategy == 2 ? pb4 : align256((size_t)range * 4);
    if (!scattered) {
        rc = rfb_ensure_work(ctx, 3 * b8 + b4 + scan::tiles_bytes(tiles_max), &w);
        if (rc) return rc;
    }
    Accums a;
    a.first_row = (u64 *)w;
    a.sum = (u64 *)((char *)w + b8);
    a.cnt = (u64 *)((char *)w + 2 * b8);
    a.has_null = (u32 *)((char *)w + 3 * b8);
    if (scattered && strategy != 2) {   // the sample misjudged the range: the accumulator layout of the other strategies must fit
        if (3 * align256((size_t)range * 8) + align256((size_t)range * 4) + scan::tiles_bytes(tiles_max) > ctx->work_bytes) {
            rc = rfb_ensure_work(ctx, 3 * b8 + b4 + scan::tiles_bytes(tiles_max), &w);
            if (rc) return rc;
            a.first_row = (u64 *)w; a.sum = (u64 *)((char *)w + b8); a.cnt = (u64 *)((char *)w + 2 * b8); a.has_null = (u32 *)((char *)w + 3 * b8);
        }
    }
    RFB_CUDA(cudaMemsetAsync(a.first_row, 0xFF, (size_t)range * 8, ctx->stream));
    RFB_CUDA(cudaMemsetAsync(a.sum, 0, (size_t)range * 8, ctx->stream));
    RFB_CUDA(cudaMemsetAsync(a.cnt, 0, (size_t)range * 8, ctx->stream));
    RFB_CUDA(cudaMemsetAsync(a.has_null, 0, (size_t)range * 4, ctx->stream));
    if (modded && range <= KP) {        // already accumulated by key residue: move the slots into key order
        k_mod_remap<<<rfb_grid_for(ctx, range, THREADS, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(gmod, kmin, range, a);
        RFB_CHECK_LAUNCH(ctx);
    } else if (strategy == 1) {
        const i64 ptiles = (n + PTILE - 1) / PTILE;
        const int pgrid = (int)(ptiles < 2ll * ctx->sm_count ? ptiles : 2ll * ctx->sm_count);
        RFB_CUDA(cudaFuncSetAttribute(k_fused_accum_smem<FS>, cudaFuncAttributeMaxDynamicSharedMemorySize, KP * 12));
        k_fused_accum_smem<FS><<<pgrid, PT, (size_t)range * 12, ctx->stream>>>(fs, val, n, vec, kmin, (int)range, a);
        RFB_CHECK_LAUNCH(ctx);
    } else if (strategy == 2) {
        const char *tma = getenv("RFB_ACCUM_TMA");       // "0": the register-staged kernel (128-bit loads) instead of the TMA ring
        if (!(tma && tma[0] == '0')) {
            const size_t smem = ASTAGES * sizeof(AccumStage) + KP * 12;
            RFB_CUDA(cudaFuncSetAttribute(k_part_accum_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_part_accum_tma<<<ctx->sm_count, AT, smem, ctx->stream>>>(ps, (int)P, kbase, kmin, a);
        } else {
            RFB_CUDA(cudaFuncSetAttribute(k_part_accum, cudaFuncAttributeMaxDynamicSharedMemorySize, KP * 12));
            k_part_accum<<<2 * ctx->sm_count, PT, KP * 12, ctx->stream>>>(ps, (int)P, kbase, kmin, a);
        }
        RFB_CHECK_LAUNCH(ctx);
    } else {
        k_fused_accum_l2<FS><<<grid, THREADS, 0, ctx->stream>>>(fs, val, n, kmin, a);
        RFB_CHECK_LAUNCH(ctx);
    }
    // first rows: claimed on a growing row prefix
    i64 r0 = 0, r1 = 32 * range > 65536 ? 32 * range : 65536;
    while (true) {
        if (r1 > n) r1 = n;
        k_fused_claim<FS><<<rfb_grid_for(ctx, r1 - r0, THREADS * 4, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(fs, r0, r1, kmin, a.first_row);
        RFB_CHECK_LAUNCH(ctx);
        if (r1 == n) break;
        RFB_CUDA(cudaMemsetAsync(mm + 3, 0, 16, ctx->stream));
        k_slot_census<<<rfb_grid_for(ctx, range, THREADS, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(a.first_row, a.cnt, range, mm);
        RFB_CHECK_LAUNCH(ctx);
        i64 census[2];
        rc = d2h_sync(ctx, census, mm + 3, 16);
        if (rc) return rc;
        if (census[0] == census[1]) break;   // every key that occurs has its first row
        r0 = r1;
        r1 = r1 * 4;
    }
    k_max_first<<<rfb_grid_for(ctx, range, THREADS, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(a.first_row, range, mm);
    RFB_CHECK_LAUNCH(ctx);
    i64 limit = 0;
    rc = d2h_sync(ctx, &limit, mm + 2, 8);
    if (rc) return rc;
    const i64 tiles = (limit + scan::RowTile<NUM_J>::TILE - 1) / scan::RowTile<NUM_J>::TILE;
    scan::TileCtl ctl;
    rc = scan::prepare_tiles(ctx, (char *)w + 3 * b8 + b4, tiles, ctx->h_count, &ctl);
    if (rc) return rc;
    k_fused_emit<FS><<<(unsigned)tiles, THREADS, 0, ctx->stream>>>(fs, limit, kmin, a, max_groups, out_keys, out_sums, out_counts, ctl);
    RFB_CHECK_LAUNCH(ctx);
    RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    *groups = *(volatile i64 *)ctx->h_count;
    if (*groups > max_groups) {
        rfb_set_error("fused group-by: %lld groups exceed the output capacity %lld", (long long)*groups, (long long)max_groups);
        return RFB_ERR_ARG;
    }
    return RFB_OK;
}

template <typename K, typename P>
int fused_pred(rfb_ctx_t *ctx, const void *keys, const void *pred, PredRange pr, const i64 *val, i64 n, i64 max_groups, i64 *ok,
               i64 *os, i64 *oc, i64 *groups) {
    FusedSrc<K, P, true> fs{(const K *)keys, (const P *)pred, pr};
    return fused_run(ctx, fs, val, n, max_groups, ok, os, oc, groups);
}

template <typename K>
int fused_key(rfb_ctx_t *ctx, const void *keys, int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k, const i64 *val,
              i64 n, i64 max_groups, i64 *ok, i64 *os, i64 *oc, i64 *groups) {
    if (!pred) {
        FusedSrc<K, i64, false> fs{(const K *)keys, nullptr, PredRange{0, 0, 0, 0}};
        return fused_run(ctx, fs, val, n, max_groups, ok, os, oc, groups);
    }
    PredRange pr;
    if (!rfb_make_pred(cmp_op, pred_type, k, &pr)) { rfb_set_error("fused group-by: unsupported predicate types"); return RFB_ERR_TYPE; }
    switch (rfb_kind_of(pred_type)) {
        case K_I32: return fused_pred<K, i32>(ctx, keys, pred, pr, val, n, max_groups, ok, os, oc, groups);
        case K_I64: return fused_pred<K, i64>(ctx, keys, pred, pr, val, n, max_groups, ok, os, oc, groups);
        case K_F64: return fused_pred<K, f64>(ctx, keys, pred, pr, val, n, max_groups, ok, os, oc, groups);
        default: rfb_set_error("fused group-by: unsupported predicate column type %d", pred_type); return RFB_ERR_TYPE;
    }
}

}  // namespace

extern "C" int rfb_group_sum_count_dev(rfb_ctx_t *ctx, int key_type, const void *keys, const int64_t *val, int64_t n,
                                       int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k,
                                       int64_t max_groups, int64_t *out_keys, int64_t *out_sums, int64_t *out_counts,
                                       int64_t *groups) {
    RFB_ARG(ctx && groups && n >= 0 && max_groups >= 0 && ((keys && val) || n == 0) && (!pred || k), "rfb_group_sum_count_dev");
    RFB_ARG((out_keys && out_sums && out_counts) || max_groups == 0, "rfb_group_sum_count_dev: outputs");
    *groups = 0;
    if (n == 0) return RFB_OK;
    switch (rfb_kind_of(key_type)) {
        case K_I32: return fused_key<i32>(ctx, keys, cmp_op, pred_type, pred, k, val, n, max_groups, out_keys, out_sums, out_counts, groups);
        case K_I64: return fused_key<i64>(ctx, keys, cmp_op, pred_type, pred, k, val, n, max_groups, out_keys, out_sums, out_counts, groups);
        default: rfb_set_error("fused group-by: key type %d (I32 or I64 keys)", key_type); return RFB_ERR_TYPE;
    }
}
