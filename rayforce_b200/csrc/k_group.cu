// k_group.cu — the group index (sm_100a): hash / perfect-hash grouping of one key column or a tuple of key columns, distinct keys.
//
//   rfb_group_i64_dev         index_group -> index_group_i64: scope + group numbering (reference core/index.c:402-435,
//                             2002-2092 perfect hash, 1777-1911 + core/hash.c open addressing)
//   rfb_group_keys_i64_dev    index_group_list (core/index.c:2731-2793): perfect-hash key fusion, else row hashing
//   rfb_distinct_i64_dev      ray_distinct -> index_distinct_i64, dense branch (core/index.c:551-577)
//   (grouped aggregates: k_aggr.cu; the fused group-by: k_fused_group.cu; shared device code: rfb_group.cuh)
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
#include "rfb_group.cuh"

namespace {


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
