// k_select.cu — selection vectors and gathers (sm_100a).
//
//   rfb_where_dev       ray_where -> ops_where: byte mask -> ascending i64 row ids   (reference core/ops.c:255-273,
//                       two sequential passes on one thread there)
//   rfb_cmp_where_dev   ray_where(ray_<cmp>(x, k)) without ever writing the mask
//   rfb_gather_dev      filter_collect -> at_ids: out[i] = col[ids[i]]              (core/rayforce.c:1036-1159)
//
// Compaction is ONE pass: a chained scan with decoupled look-back (rfb_scan.cuh).  Each CTA counts its selected rows with
// warp ballots, publishes (AGGREGATE | count) in a 64-bit status word, walks back over its predecessors' words 32 at a
// time until it meets an INCLUSIVE one, publishes its own inclusive prefix and then writes its row ids at that offset.
// Tiles are 16 K rows so that one look-back (a CTA barrier around a chain of L2 round trips) is amortised.  Order is preserved:
// within a warp the rank of a row is a popcount over the ballot masks of lower rows.  Traffic: mask bytes (or the 8-byte
// predicate column) read once + 8 bytes written per selected row.
#include "rfb_scan.cuh"

namespace {

using scan::THREADS;
using scan::WARPS;
using scan::TileCtl;
using scan::TileSmem;
constexpr int BLOCKS_PER_SM = 4;

// ---- byte mask -> ids.  Lane owns 16 consecutive mask bytes per step (one 128-bit load), J steps per tile.
// The ids of one warp-step (<= 512) are staged in shared memory in output order and written out by the whole warp as
// contiguous 256-byte stores: writing them straight from the lanes touches 32 different sectors per instruction (each lane
// owns a ~64-byte run), which quadrupled the L2 write transactions.
// Tile size: one look-back per tile (a CTA barrier around a chain of L2 round trips) is the kernels' fixed cost, so long columns
// take LARGE tiles (mask: 48 K rows, MASK_J = 12; typed predicate: 24 sub-tiles = 96 K rows of an 8-byte column) — measured per
// 1e9 rows: where 1.42 -> 1.10 ms, cmp+where 2.49 -> 1.98 ms (0.93 of the HBM peak) against the 16 K-row tiles, which columns that
// would not fill the GPU with large tiles keep, with a 32 K-row tier in between (mask J = 2 / 4 / 8 / 12: 1.73 / 1.42 / 1.17 / 1.10 ms; SUB = 2 / 4 / 8 / 16 / 24 / 40:
// 2.95 / 2.49 / 2.15 / 2.03 / 1.98 / 2.00 ms).
constexpr int MASK_J_SMALL = 4, MASK_J_MEDIUM = 8, MASK_J_LARGE = 12, CMP_SUB_SMALL = 4, CMP_SUB_MEDIUM = 8, CMP_SUB_LARGE = 24;
template <int MASK_J> struct MaskTile { static constexpr int TILE = THREADS * 16 * MASK_J; };

template <int MASK_J>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_where_mask(const u8 *__restrict__ mask, i64 n, i64 *__restrict__ ids, TileCtl ctl) {
    constexpr int MASK_TILE = MaskTile<MASK_J>::TILE;
    __shared__ TileSmem sm;
    __shared__ i64 stage[WARPS][512];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 tile = blockIdx.x;
    const i64 wbase = (i64)tile * MASK_TILE + (i64)warp * (32 * 16 * MASK_J);
    u32 bits[MASK_J];    // this lane's 16 flags per step
    u32 excl[MASK_J];    // selected rows of this warp-step before this lane's first row
    u32 tot[MASK_J];     // selected rows of this warp-step
    u32 warp_total = 0;
    Vec16<u8> v[MASK_J];
    const bool full = wbase + 32 * 16 * MASK_J <= n;
    if (full) {
#pragma unroll
        for (int j = 0; j < MASK_J; j++) v[j].raw = ld_stream16(mask + wbase + ((i64)j * 32 + lane) * 16);
    }
#pragma unroll
    for (int j = 0; j < MASK_J; j++) {
        const i64 r0 = wbase + ((i64)j * 32 + lane) * 16;
        u32 b = 0;
        if (full) {
#pragma unroll
            for (int e = 0; e < 16; e++) b |= (v[j].e[e] != 0 ? 1u : 0u) << e;
        } else {
            for (int e = 0; e < 16; e++)
                if (r0 + e < n && mask[r0 + e] != 0) b |= 1u << e;
        }
        bits[j] = b;
        u32 c = __popc(b), incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        excl[j] = incl - c;
        tot[j] = __shfl_sync(0xffffffffu, incl, 31);
        warp_total += tot[j];
    }
    u64 obase = scan::tile_offsets<WARPS>(ctl, tile, warp_total, sm);
#pragma unroll
    for (int j = 0; j < MASK_J; j++) {
        const i64 r0 = wbase + ((i64)j * 32 + lane) * 16;
        i64 *st = &stage[warp][excl[j]];
        u32 b = bits[j];
        while (b) {
            const int e = __ffs(b) - 1;
            b &= b - 1;
            *st++ = r0 + e;
        }
        __syncwarp();
        for (u32 q = lane; q < tot[j]; q += 32) ids[obase + q] = stage[warp][q];
        __syncwarp();
        obase += tot[j];
    }
}

// ---- predicate on a typed column -> ids.  Lane owns R = 16/sizeof(P) consecutive rows per step; a warp walks SUB
// sub-tiles of J steps each, keeping only the selection bits (J*R per sub-tile), so one look-back serves 16 K rows.
// The store phase recomputes the in-warp ranks from the same ballots instead of keeping them in registers.
template <typename P, int SUB_> struct CmpTile {
    static constexpr int R = 16 / (int)sizeof(P);
    static constexpr int J = (R <= 4) ? 8 : 32 / R;          // J*R <= 32 selection bits per sub-tile
    static constexpr int SUB = SUB_;
    static constexpr int SROWS = 32 * R * J;                 // rows per warp per sub-tile
    static constexpr int WROWS = SROWS * SUB;
    static constexpr int TILE = WARPS * WROWS;
};

template <typename P, int SUBT>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_where_cmp(const P *__restrict__ x, PredRange pr, i64 n, bool vec_ok, i64 *__restrict__ ids, TileCtl ctl) {
    typedef CmpTile<P, SUBT> CT;
    constexpr int R = CT::R, J = CT::J, SUB = CT::SUB;
    __shared__ TileSmem sm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 tile = blockIdx.x;
    const u32 lt = (1u << lane) - 1u;
    const i64 wbase = (i64)tile * CT::TILE + (i64)warp * CT::WROWS;
    __shared__ u32 bits[SUB][THREADS];   // bit (j*R + e) = row (sub, j, lane, e) selected; thread-private column
    u32 warp_total = 0;
#pragma unroll 1
    for (int s = 0; s < SUB; s++) {
        const i64 sbase = wbase + (i64)s * CT::SROWS;
        Vec16<P> v[J];
        const bool full = vec_ok && sbase + CT::SROWS <= n;
        if (full) {
#pragma unroll
            for (int j = 0; j < J; j++) v[j].raw = ld_stream16(x + sbase + ((i64)j * 32 + lane) * R);
        } else {
#pragma unroll
            for (int j = 0; j < J; j++)
#pragma unroll
                for (int e = 0; e < R; e++) {
                    const i64 r = sbase + ((i64)j * 32 + lane) * R + e;
                    v[j].e[e] = r < n ? x[r] : P();
                }
        }
        u32 b = 0;
#pragma unroll
        for (int j = 0; j < J; j++)
#pragma unroll
            for (int e = 0; e < R; e++) {
                const i64 r = sbase + ((i64)j * 32 + lane) * R + e;
                const bool sel = (full || r < n) && pred_test(pred_key<P>(v[j].e[e]), pr);
                b |= (sel ? 1u : 0u) << (j * R + e);
            }
        bits[s][threadIdx.x] = b;
        warp_total += __reduce_add_sync(0xffffffffu, (u32)__popc(b));
    }
    u64 obase = scan::tile_offsets<WARPS>(ctl, tile, warp_total, sm);
#pragma unroll 1
    for (int s = 0; s < SUB; s++) {
        const i64 sbase = wbase + (i64)s * CT::SROWS;
        const u32 mybits = bits[s][threadIdx.x];
#pragma unroll
        for (int j = 0; j < J; j++) {
            // rows of step j in row order: lane-major, element-minor -> rank = selected elements of lower lanes + own lower ones
            u32 before = 0, tot = 0;
#pragma unroll
            for (int e = 0; e < R; e++) {
                const u32 m = __ballot_sync(0xffffffffu, (mybits >> (j * R + e)) & 1u);
                before += __popc(m & lt);
                tot += __popc(m);
            }
            i64 *o = ids + obase + before;
            const i64 r0 = sbase + ((i64)j * 32 + lane) * R;
#pragma unroll
            for (int e = 0; e < R; e++)
                if ((mybits >> (j * R + e)) & 1u) *o++ = r0 + e;
            obase += tot;
        }
    }
}

int prepare_tiles(rfb_ctx_t *ctx, i64 tiles, TileCtl *ctl) {
    void *w;
    int rc = rfb_ensure_work(ctx, scan::tiles_bytes(tiles), &w);
    if (rc) return rc;
    return scan::prepare_tiles(ctx, w, tiles, ctx->h_count, ctl);
}

int finish_count(rfb_ctx_t *ctx, i64 *count) {
    RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    *count = *(volatile i64 *)ctx->h_count;
    return RFB_OK;
}

template <typename P, int SUB>
int where_cmp_launch(rfb_ctx_t *ctx, const void *x, PredRange pr, i64 n, i64 *ids) {
    const i64 tiles = (n + CmpTile<P, SUB>::TILE - 1) / CmpTile<P, SUB>::TILE;
    TileCtl ctl;
    int rc = prepare_tiles(ctx, tiles, &ctl);
    if (rc) return rc;
    k_where_cmp<P, SUB><<<(unsigned)tiles, THREADS, 0, ctx->stream>>>((const P *)x, pr, n, aligned16(x), ids, ctl);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}
// large tiles when they still give every SM several rounds of CTAs
inline bool large_tiles(const rfb_ctx_t *ctx, i64 n, i64 large_tile) { return n / large_tile >= (i64)ctx->sm_count * BLOCKS_PER_SM * 2; }

template <typename P>
int where_cmp_t(rfb_ctx_t *ctx, const void *x, PredRange pr, i64 n, i64 *ids) {
    if (large_tiles(ctx, n, CmpTile<P, CMP_SUB_LARGE>::TILE)) return where_cmp_launch<P, CMP_SUB_LARGE>(ctx, x, pr, n, ids);
    if (large_tiles(ctx, n, CmpTile<P, CMP_SUB_MEDIUM>::TILE)) return where_cmp_launch<P, CMP_SUB_MEDIUM>(ctx, x, pr, n, ids);
    return where_cmp_launch<P, CMP_SUB_SMALL>(ctx, x, pr, n, ids);
}

// ---- gather
template <typename T>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM)
k_gather(const T *__restrict__ col, const i64 *__restrict__ ids, i64 m, T *__restrict__ out) {
    constexpr int UNROLL = 8;
    const i64 stride = (i64)gridDim.x * THREADS;
    i64 i = (i64)blockIdx.x * THREADS + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < m; i += UNROLL * stride) {
        i64 id[UNROLL];
        T v[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; j++) id[j] = ld_stream(ids + i + j * stride);
#pragma unroll
        for (int j = 0; j < UNROLL; j++) v[j] = __ldg(col + id[j]);
#pragma unroll
        for (int j = 0; j < UNROLL; j++) __stcs(out + i + j * stride, v[j]);
    }
    for (; i < m; i += stride) out[i] = __ldg(col + ld_stream(ids + i));
}

template <typename T> int gather_t(rfb_ctx_t *ctx, const void *col, const i64 *ids, i64 m, void *out) {
    const int grid = rfb_grid_for(ctx, m, THREADS * 8, BLOCKS_PER_SM * 2);
    k_gather<T><<<grid, THREADS, 0, ctx->stream>>>((const T *)col, ids, m, (T *)out);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

}  // namespace

extern "C" int rfb_where_dev(rfb_ctx_t *ctx, const uint8_t *mask, int64_t n, int64_t *ids, int64_t *count) {
    RFB_ARG(ctx && count && n >= 0 && ((mask && ids) || n == 0), "rfb_where_dev");
    *count = 0;
    if (n == 0) return RFB_OK;
    TileCtl ctl;
    if (aligned16(mask)) {
        const int tier = large_tiles(ctx, n, MaskTile<MASK_J_LARGE>::TILE) ? 2 : large_tiles(ctx, n, MaskTile<MASK_J_MEDIUM>::TILE) ? 1 : 0;
        const i64 tile = tier == 2 ? MaskTile<MASK_J_LARGE>::TILE : tier == 1 ? MaskTile<MASK_J_MEDIUM>::TILE : MaskTile<MASK_J_SMALL>::TILE;
        const i64 tiles = (n + tile - 1) / tile;
        int rc = prepare_tiles(ctx, tiles, &ctl);
        if (rc) return rc;
        if (tier == 2) k_where_mask<MASK_J_LARGE><<<(unsigned)tiles, THREADS, 0, ctx->stream>>>(mask, n, ids, ctl);
        else if (tier == 1) k_where_mask<MASK_J_MEDIUM><<<(unsigned)tiles, THREADS, 0, ctx->stream>>>(mask, n, ids, ctl);
        else k_where_mask<MASK_J_SMALL><<<(unsigned)tiles, THREADS, 0, ctx->stream>>>(mask, n, ids, ctl);
        RFB_CHECK_LAUNCH(ctx);
    } else {
        // unaligned payload: treat the bytes as a U8 column and select != 0 with the typed kernel's scalar path
        PredRange pr = make_pred_range(RFB_NE, key_of_i64(0));
        const i64 tiles = (n + CmpTile<u8, CMP_SUB_SMALL>::TILE - 1) / CmpTile<u8, CMP_SUB_SMALL>::TILE;
        int rc = prepare_tiles(ctx, tiles, &ctl);
        if (rc) return rc;
        k_where_cmp<u8, CMP_SUB_SMALL><<<(unsigned)tiles, THREADS, 0, ctx->stream>>>(mask, pr, n, false, ids, ctl);
        RFB_CHECK_LAUNCH(ctx);
    }
    return finish_count(ctx, count);
}

extern "C" int rfb_cmp_where_dev(rfb_ctx_t *ctx, int op, int type, const void *x, int64_t n, const rfb_scalar_t *k,
                                 int64_t *ids, int64_t *count) {
    RFB_ARG(ctx && count && k && n >= 0 && ((x && ids) || n == 0), "rfb_cmp_where_dev");
    *count = 0;
    PredRange pr;
    if (!rfb_make_pred(op, type, k, &pr)) { rfb_set_error("cmp+where: unsupported column/constant types %d, %d", type, k->type); return RFB_ERR_TYPE; }
    if (n == 0) return RFB_OK;
    int rc;
    switch (rfb_kind_of(type)) {
        case K_U8: rc = where_cmp_t<u8>(ctx, x, pr, n, ids); break;
        case K_I16: rc = where_cmp_t<i16>(ctx, x, pr, n, ids); break;
        case K_I32: rc = where_cmp_t<i32>(ctx, x, pr, n, ids); break;
        case K_I64: rc = where_cmp_t<i64>(ctx, x, pr, n, ids); break;
        default: rc = where_cmp_t<f64>(ctx, x, pr, n, ids); break;
    }
    if (rc) return rc;
    return finish_count(ctx, count);
}

extern "C" int rfb_gather_dev(rfb_ctx_t *ctx, int type, const void *col, const int64_t *ids, int64_t m, void *out) {
    RFB_ARG(ctx && m >= 0 && ((col && ids && out) || m == 0), "rfb_gather_dev");
    if (m == 0) return RFB_OK;
    switch (rfb_type_size(type)) {
        case 1: return gather_t<u8>(ctx, col, ids, m, out);
        case 2: return gather_t<i16>(ctx, col, ids, m, out);
        case 4: return gather_t<i32>(ctx, col, ids, m, out);
        case 8: return gather_t<i64>(ctx, col, ids, m, out);
        default: rfb_set_error("gather: unsupported type %d", type); return RFB_ERR_TYPE;
    }
}
