// k_hash_group.cu — fused `select {s: (sum v) c: (count v) from t by k [where ...]}` on SPARSE key domains (sm_100a).
//
// Reference: index_group_i64_unscoped -> index_group_distribute (core/index.c:1959-1977, 1777-1911): per-chunk open-addressing
// tables (core/hash.c:35-148, linear probing, empty = NULL_I64), a merge of the chunk tables, then aggr_sum / aggr_count over
// the group ids.  The device keeps the same structure at two levels:
//   * every CTA (one per SM) owns an open-addressing table of HG_CAP slots in SHARED memory — key, 64-bit wrapping sum as two
//     32-bit words, row count with the sticky-null bit — fed by key / value / predicate tiles that the TMA unit stages into a
//     shared-memory ring (cp.async.bulk + mbarrier).  A row costs one probe (64-bit CAS only on first sight of a key) and the
//     usual two or three shared atomics; nothing leaves the SM while the CTA's distinct keys fit its table;
//   * when a table is about to run out of room for one more tile it is SPILLED: its entries are merged into the device-wide
//     open-addressing table (64-bit CAS on the key, L2 atomics on sum / count) and the table restarts empty; the same merge
//     runs at the end.  Low-cardinality sparse keys (symbol ids, hashes, timestamps rounded to a bar) therefore run at the
//     speed of the column scan, high-cardinality ones degrade to one device-wide probe per row.
// Groups come out in first-occurrence order (what the reference produces with one thread; with more threads its
// order is hash-slot order and parity is checked on key-sorted output): first rows are claimed on a growing row prefix
// exactly like the dense path does, the occupied slots are compacted and ordered by first row with the radix sort.
#include "rfb_group.cuh"
#include "rfb_tma.cuh"

namespace {

constexpr int HG_T = 1024;                 // threads per CTA; one CTA per SM
constexpr int HG_CAP = 8192;               // slots of the CTA table (+ 1 dedicated slot for the NULL_I64 key)
constexpr int HG_R = 2;                    // rows per thread and tile
constexpr int HG_TILE = HG_T * HG_R;       // rows per tile
constexpr int HG_STAGES = 2;
constexpr int HG_SPILL_AT = HG_CAP - HG_TILE - HG_CAP / 8;   // spill when a further tile could push the load past 7/8
constexpr u64 HG_EMPTY = (u64)NULL_I64;

struct GTable {
    u64 *keys;       // [cap + 1] EMPTY = NULL_I64; slot cap = the NULL_I64 key itself
    u64 *sum;        // wrapping sums of the non-null values
    u64 *cnt;        // rows
    u64 *first;      // first row (NO_ROW until claimed)
    u32 *has_null;   // sticky-null marker of the sum
    u32 *ctl;        // [0] groups inserted, [1] overflow (more than max_groups), [2] NULL key seen
    u64 mask;
    i64 cap;
    u32 max_groups;
};

// device-wide insert (or find) of a key: the slot it lives in.  *overflow is set when this very insertion took the table past
// the number of groups the caller has room for (the table itself has slack for the insertions still in flight).
__device__ __forceinline__ i64 g_insert(const GTable &g, i64 key, bool *overflow) {
    if (key == NULL_I64) {
        if (atomicCAS(&g.ctl[2], 0u, 1u) == 0u && atomicAdd(&g.ctl[0], 1u) >= g.max_groups) { g.ctl[1] = 1u; *overflow = true; }
        return g.cap;
    }
    u64 s = mix64((u64)key) & g.mask;
    while (true) {
        const u64 cur = scan::ld_relaxed(&g.keys[s]);
        if (cur == (u64)key) return (i64)s;
        if (cur == HG_EMPTY) {
            const u64 old = atomicCAS((unsigned long long *)&g.keys[s], (unsigned long long)HG_EMPTY, (unsigned long long)key);
            if (old == HG_EMPTY) {
                if (atomicAdd(&g.ctl[0], 1u) >= g.max_groups) { g.ctl[1] = 1u; *overflow = true; }
                return (i64)s;
            }
            if (old == (u64)key) return (i64)s;
        }
        s = (s + 1) & g.mask;
    }
}
__device__ __forceinline__ i64 g_lookup(const GTable &g, i64 key) {   // a key known to be present
    if (key == NULL_I64) return g.cap;
    u64 s = mix64((u64)key) & g.mask;
    while (__ldg(&g.keys[s]) != (u64)key) s = (s + 1) & g.mask;
    return (i64)s;
}
// one row straight into the device-wide table
__device__ __forceinline__ void g_add(const GTable &g, i64 key, i64 v, bool *overflow) {
    const i64 gs = g_insert(g, key, overflow);
    if (v == NULL_I64) g.has_null[gs] = 1u; else atomicAdd((unsigned long long *)g.sum + gs, (unsigned long long)v);
    atomicAdd((unsigned long long *)g.cnt + gs, 1ULL);
}

struct CtaTable {
    u64 *keys;       // [HG_CAP + 1]
    u32 *lo, *hi, *cnt;
};

__device__ __forceinline__ u32 hg_hash(u64 key) { return (((u32)key ^ (u32)(key >> 32)) * 0x9E3779B1u) >> (32 - 13); }
static_assert(HG_CAP == 1 << 13, "hg_hash yields 13 bits");

// merge the CTA table into the device-wide table and clear it (all threads of the CTA); returns this thread's entry count
__device__ __forceinline__ u32 hg_spill(const CtaTable &t, const GTable &g, bool *overflow) {
    u32 entries = 0;
    for (int s = threadIdx.x; s <= HG_CAP; s += HG_T) {
        const u32 c = t.cnt[s];
        if (c) {
            entries++;
            const i64 key = s == HG_CAP ? NULL_I64 : (i64)t.keys[s];
            const i64 gs = g_insert(g, key, overflow);
            const u64 sum = ((u64)t.hi[s] << 32) | t.lo[s];
            if (sum) atomicAdd((unsigned long long *)g.sum + gs, (unsigned long long)sum);
            atomicAdd((unsigned long long *)g.cnt + gs, (unsigned long long)(c & ~NULL_FLAG));
            if (c & NULL_FLAG) g.has_null[gs] = 1u;
            t.lo[s] = 0; t.hi[s] = 0; t.cnt[s] = 0;
        }
        if (s < HG_CAP) t.keys[s] = HG_EMPTY;
    }
    return entries;
}

template <typename K> struct __align__(16) HgStage {
    K key[HG_TILE];
    i64 val[HG_TILE];
};

// Every CTA starts with its shared-memory table.  After a spill it looks at what the table bought: when fewer than
// HG_MIN_ROWS_PER_ENTRY rows were folded per spilled entry (a column with about as many distinct keys as rows) the CTA goes
// DIRECT for the rest of its rows — one device-wide probe and two L2 atomics per row, no table upkeep.
constexpr u32 HG_MIN_ROWS_PER_ENTRY = 4;

template <typename K, typename P, bool HAS_PRED, bool TMA>
__global__ void __launch_bounds__(HG_T, 1)
k_hg_accum(const K *__restrict__ keys, const i64 *__restrict__ val, const P *__restrict__ pred, PredRange pr, i64 n, GTable g) {
    typedef HgStage<K> Stage;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    Stage *stage = (Stage *)s_dyn;
    CtaTable t;
    t.keys = (u64 *)(s_dyn + HG_STAGES * sizeof(Stage));
    t.lo = (u32 *)(t.keys + HG_CAP + 1);
    t.hi = t.lo + HG_CAP + 1;
    t.cnt = t.hi + HG_CAP + 1;
    __shared__ u64 full[HG_STAGES];
    __shared__ u32 s_occupied, s_spilled, s_stop, s_direct;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int s = tid; s <= HG_CAP; s += HG_T) { t.keys[s] = HG_EMPTY; t.lo[s] = 0; t.hi[s] = 0; t.cnt[s] = 0; }
    if (tid == 0) {
        s_occupied = 0; s_spilled = 0; s_stop = 0; s_direct = 0;
        for (int s = 0; s < HG_STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const i64 tiles = (n + HG_TILE - 1) / HG_TILE;
    auto issue = [&](i64 tile, int s) {     // thread 0: stage the keys and values of a FULL tile with the TMA unit
        if (!TMA || (tile + 1) * HG_TILE > n) return;
        mbar_expect_tx(&full[s], HG_TILE * (u32)(sizeof(K) + 8));
        bulk_g2s(stage[s].key, keys + tile * HG_TILE, HG_TILE * (u32)sizeof(K), &full[s]);
        bulk_g2s(stage[s].val, val + tile * HG_TILE, HG_TILE * 8u, &full[s]);
    };
    if (tid == 0 && (i64)blockIdx.x < tiles) issue(blockIdx.x, 0);
    i64 it = 0;
    u32 tiles_since_spill = 0;
    bool overflow = false;
    for (i64 tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
        const int s = (int)(it % HG_STAGES);
        __syncthreads();                                   // the previous tile is consumed: its stage may be re-armed, the counters are final
        if (s_stop) break;                                 // more groups than the caller has room for (uniform: written before the barrier)
        if (!s_direct && s_occupied > (u32)HG_SPILL_AT) {
            const u32 e = hg_spill(t, g, &overflow);
            const u32 we = __reduce_add_sync(0xffffffffu, e);
            if (lane == 0 && we) atomicAdd(&s_spilled, we);
            if (overflow) s_stop = 1;
            __syncthreads();
            if (tid == 0) {
                if (tiles_since_spill * (u32)HG_TILE < HG_MIN_ROWS_PER_ENTRY * s_spilled) s_direct = 1;
                s_occupied = 0;
                s_spilled = 0;
            }
            tiles_since_spill = 0;
            __syncthreads();
        }
        tiles_since_spill++;
        if (tid == 0 && tile + gridDim.x < tiles) issue(tile + gridDim.x, (int)((it + 1) % HG_STAGES));
        const i64 base = tile * HG_TILE;
        const bool staged = TMA && base + HG_TILE <= n, direct = s_direct != 0;
        if (staged) mbar_wait(&full[s], (u32)(it / HG_STAGES) & 1u);
        u32 fresh = 0;
#pragma unroll
        for (int j = 0; j < HG_R; j++) {
            const int q = j * HG_T + tid;
            const i64 r = base + q;
            if (r >= n) continue;
            if constexpr (HAS_PRED) { if (!pred_test(pred_key<P>(ld_stream(pred + r)), pr)) continue; }
            const i64 key = staged ? (i64)stage[s].key[q] : (i64)ld_stream(keys + r);
            const i64 v = staged ? stage[s].val[q] : ld_stream(val + r);
            if (direct) { g_add(g, key, v, &overflow); continue; }
            u32 slot;
            if (key == NULL_I64) slot = HG_CAP;
            else {
                slot = hg_hash((u64)key);
                while (true) {
                    const u64 cur = *(volatile u64 *)&t.keys[slot];
                    if (cur == (u64)key) break;
                    if (cur == HG_EMPTY) {
                        const u64 old = atomicCAS((unsigned long long *)&t.keys[slot], (unsigned long long)HG_EMPTY, (unsigned long long)key);
                        if (old == HG_EMPTY) { fresh++; break; }
                        if (old == (u64)key) break;
                    }
                    slot = (slot + 1) & (HG_CAP - 1);
                }
            }
            sacc_add(SAcc{t.lo, t.hi, t.cnt}, slot, v);
        }
        if (!direct) {
            const u32 nf = __reduce_add_sync(0xffffffffu, fresh);
            if (lane == 0 && nf) atomicAdd(&s_occupied, nf);
        } else if (overflow) s_stop = 1;
    }
    __syncthreads();
    hg_spill(t, g, &overflow);
}

// first rows on the row range [r0, r1): every group is already in the table
template <typename K, typename P, bool HAS_PRED>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_hg_claim(const K *__restrict__ keys, const P *__restrict__ pred, PredRange pr, i64 r0, i64 r1, GTable g) {
    for (i64 i = r0 + (i64)blockIdx.x * THREADS + threadIdx.x; i < r1; i += (i64)gridDim.x * THREADS) {
        if constexpr (HAS_PRED) { if (!pred_test(pred_key<P>(ld_stream(pred + i)), pr)) continue; }
        claim_first(g.first, g_lookup(g, (i64)ld_stream(keys + i)), i);
    }
}

// occupied slots -> dense arrays (slot, first row), in slot order
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_hg_compact(GTable g, i64 *out_slot, i64 *out_first, scan::TileCtl ctl) {
    __shared__ scan::TileSmem sm;
    scan::compact_rows<NUM_J>(
        g.cap + 1, ctl, sm, [&](i64 s) { return s == g.cap ? g.ctl[2] != 0u : g.keys[s] != HG_EMPTY; },
        [&](i64 s, i64 i) { out_slot[i] = s; out_first[i] = (i64)g.first[s]; });
}

__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_hg_emit(GTable g, const i64 *__restrict__ slot_of, const i64 *__restrict__ perm, i64 groups, i64 *out_keys, i64 *out_sums, i64 *out_counts) {
    for (i64 i = (i64)blockIdx.x * THREADS + threadIdx.x; i < groups; i += (i64)gridDim.x * THREADS) {
        const i64 s = slot_of[perm[i]];
        out_keys[i] = s == g.cap ? NULL_I64 : (i64)g.keys[s];
        out_sums[i] = g.has_null[s] ? NULL_I64 : (i64)g.sum[s];
        out_counts[i] = (i64)g.cnt[s];
    }
}

template <typename K, typename P, bool HAS_PRED>
int hg_run(rfb_ctx_t *ctx, const K *keys, const i64 *val, const P *pred, PredRange pr, i64 n, i64 max_groups, i64 *out_keys, i64 *out_sums,
           i64 *out_counts, i64 *groups) {
    if (n >= 0xFFFFFFF0ll || max_groups >= 0x7FFFFFFFll) { rfb_set_error("sparse group-by: more than 2^32 rows / 2^31 groups"); return RFB_ERR_ARG; }
    // device-wide table: twice the groups the caller has room for, plus slack for the inserts in flight when the overflow is noticed
    i64 cap = 1 << 16;
    while (cap < 2 * max_groups + (1 << 19)) cap <<= 1;
    const size_t b8 = align256((size_t)(cap + 1) * 8), b4 = align256((size_t)(cap + 1) * 4);
    const i64 ctiles = (cap + 1 + scan::RowTile<NUM_J>::TILE - 1) / scan::RowTile<NUM_J>::TILE;
    const size_t bg = align256((size_t)(max_groups > 0 ? max_groups : 1) * 8);
    void *aux;
    int rc = rfb_ensure_aux(ctx, 4 * b8 + b4 + 256 + scan::tiles_bytes(ctiles) + 3 * bg, &aux);
    if (rc) return rc;
    char *w = (char *)aux;
    GTable g;
    g.keys = (u64 *)w;
    g.sum = (u64 *)(w + b8);
    g.cnt = (u64 *)(w + 2 * b8);
    g.first = (u64 *)(w + 3 * b8);
    g.has_null = (u32 *)(w + 4 * b8);
    g.ctl = (u32 *)(w + 4 * b8 + b4);
    g.mask = (u64)cap - 1;
    g.cap = cap;
    g.max_groups = (u32)max_groups;
    char *tiles_mem = w + 4 * b8 + b4 + 256;
    i64 *c_slot = (i64 *)(tiles_mem + scan::tiles_bytes(ctiles)), *c_first = (i64 *)((char *)c_slot + bg), *c_perm = (i64 *)((char *)c_slot + 2 * bg);
    rc = fill<u64>(ctx, g.keys, cap + 1, HG_EMPTY);
    if (rc) return rc;
    RFB_CUDA(cudaMemsetAsync(g.sum, 0, 2 * b8, ctx->stream));                 // sum, cnt
    RFB_CUDA(cudaMemsetAsync(g.first, 0xFF, b8, ctx->stream));
    RFB_CUDA(cudaMemsetAsync(g.has_null, 0, b4 + 256, ctx->stream));          // has_null, ctl

    const size_t smem = HG_STAGES * sizeof(HgStage<K>) + (size_t)(HG_CAP + 1) * 20 + 16;
    const bool tma = aligned16(keys) && aligned16(val);
    const i64 tiles = (n + HG_TILE - 1) / HG_TILE;
    const int grid = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
    if (tma) {
        RFB_CUDA(cudaFuncSetAttribute(k_hg_accum<K, P, HAS_PRED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_hg_accum<K, P, HAS_PRED, true><<<grid, HG_T, smem, ctx->stream>>>(keys, val, pred, pr, n, g);
    } else {
        RFB_CUDA(cudaFuncSetAttribute(k_hg_accum<K, P, HAS_PRED, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_hg_accum<K, P, HAS_PRED, false><<<grid, HG_T, smem, ctx->stream>>>(keys, val, pred, pr, n, g);
    }
    RFB_CHECK_LAUNCH(ctx);
    u32 ctl[4];
    rc = d2h_sync(ctx, ctl, g.ctl, 16);
    if (rc) return rc;
    const i64 G = ctl[0];
    *groups = G;
    if (ctl[1] || G > max_groups) {
        rfb_set_error("fused group-by: more than %lld groups (the output capacity)", (long long)max_groups);
        return RFB_ERR_ARG;
    }
    if (G == 0) return RFB_OK;
    // first rows: claimed on a growing row prefix until every group has one
    i64 *mm = (i64 *)((char *)ctx->d_scratch + 32768);
    i64 r0 = 0, r1 = 32 * G > 65536 ? 32 * G : 65536;
    while (true) {
        if (r1 > n) r1 = n;
        k_hg_claim<K, P, HAS_PRED><<<rfb_grid_for(ctx, r1 - r0, THREADS * 4, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(keys, pred, pr, r0, r1, g);
        RFB_CHECK_LAUNCH(ctx);
        if (r1 == n) break;
        RFB_CUDA(cudaMemsetAsync(mm + 4, 0, 8, ctx->stream));
        k_claimed_count<<<rfb_grid_for(ctx, cap + 1, THREADS, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(g.first, cap + 1, mm);
        RFB_CHECK_LAUNCH(ctx);
        i64 claimed = 0;
        rc = d2h_sync(ctx, &claimed, mm + 4, 8);
        if (rc) return rc;
        if (claimed == G) break;
        r0 = r1;
        r1 *= 4;
    }
    scan::TileCtl tctl;
    rc = scan::prepare_tiles(ctx, tiles_mem, ctiles, ctx->h_count, &tctl);
    if (rc) return rc;
    k_hg_compact<<<(unsigned)ctiles, THREADS, 0, ctx->stream>>>(g, c_slot, c_first, tctl);
    RFB_CHECK_LAUNCH(ctx);
    rc = rfb_sort_dev(ctx, RFB_I64, c_first, G, 0, c_perm);   // groups in first-occurrence order (workspace: ctx->d_work, ours is d_aux)
    if (rc) return rc;
    k_hg_emit<<<rfb_grid_for(ctx, G, THREADS, BLOCKS_PER_SM), THREADS, 0, ctx->stream>>>(g, c_slot, c_perm, G, out_keys, out_sums, out_counts);
    RFB_CHECK_LAUNCH(ctx);
    RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return RFB_OK;
}

template <typename K>
int hg_key(rfb_ctx_t *ctx, const void *keys, const i64 *val, int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k, i64 n, i64 max_groups,
           i64 *ok, i64 *os, i64 *oc, i64 *groups) {
    if (!pred) return hg_run<K, i64, false>(ctx, (const K *)keys, val, nullptr, PredRange{0, 0, 0, 0}, n, max_groups, ok, os, oc, groups);
    PredRange pr;
    if (!rfb_make_pred(cmp_op, pred_type, k, &pr)) { rfb_set_error("fused group-by: unsupported predicate types"); return RFB_ERR_TYPE; }
    switch (rfb_kind_of(pred_type)) {
        case K_I32: return hg_run<K, i32, true>(ctx, (const K *)keys, val, (const i32 *)pred, pr, n, max_groups, ok, os, oc, groups);
        case K_I64: return hg_run<K, i64, true>(ctx, (const K *)keys, val, (const i64 *)pred, pr, n, max_groups, ok, os, oc, groups);
        case K_F64: return hg_run<K, f64, true>(ctx, (const K *)keys, val, (const f64 *)pred, pr, n, max_groups, ok, os, oc, groups);
        default: rfb_set_error("fused group-by: unsupported predicate column type %d", pred_type); return RFB_ERR_TYPE;
    }
}

}  // namespace

// the sparse-domain body of rfb_group_sum_count_dev (k_fused_group.cu)
int rfb_hash_group_sum_count(rfb_ctx_t *ctx, int key_type, const void *keys, const int64_t *val, int64_t n, int cmp_op, int pred_type,
                             const void *pred, const rfb_scalar_t *k, int64_t max_groups, int64_t *out_keys, int64_t *out_sums,
                             int64_t *out_counts, int64_t *groups) {
    switch (rfb_kind_of(key_type)) {
        case K_I32: return hg_key<i32>(ctx, keys, val, cmp_op, pred_type, pred, k, n, max_groups, out_keys, out_sums, out_counts, groups);
        case K_I64: return hg_key<i64>(ctx, keys, val, cmp_op, pred_type, pred, k, n, max_groups, out_keys, out_sums, out_counts, groups);
        default: rfb_set_error("fused group-by: key type %d (I32 or I64 keys)", key_type); return RFB_ERR_TYPE;
    }
}
