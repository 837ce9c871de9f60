// rfb_scan.cuh — single-pass order-preserving compaction (chained scan with decoupled look-back), shared by the
// selection-vector kernels (k_select.cu) and the group-numbering kernels (k_group.cu).
//
// Tile index = blockIdx.x.  A tile only waits on lower-numbered tiles, and the hardware work distributor hands out the
// CTAs of a 1-D grid in increasing index order, so every predecessor has started by the time a tile spins on its status
// word (the same assumption CUB's DeviceScan makes; an atomic ticket per tile cost ~25 % of the kernel on B200: one
// global round trip plus a CTA barrier before the first load could be issued).  Each CTA counts its selected rows with
// warp ballots, publishes (AGGREGATE | count) in a 64-bit status word, walks back over its predecessors' words 32 at a
// time until it meets an INCLUSIVE one, publishes its own inclusive prefix and then writes its outputs at that offset.
#pragma once
#include "rfb_common.cuh"

namespace scan {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;

constexpr u64 ST_AGG = 1ULL << 62, ST_INC = 2ULL << 62, ST_FLAGS = 3ULL << 62, ST_VAL = ~ST_FLAGS;

__device__ __forceinline__ u64 ld_relaxed(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u64 *p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

// Exclusive prefix of this tile over all previous tiles.  Called by every thread of warp 0; returns it in all lanes.
__device__ __forceinline__ u64 lookback(u64 *state, u32 tile, u64 block_total) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) st_relaxed(&state[tile], (tile == 0 ? ST_INC : ST_AGG) | block_total);
    if (tile == 0) return 0;
    u64 exclusive = 0;
    i64 look = (i64)tile - 1;
    while (true) {
        const i64 idx = look - lane;
        u64 s;
        do {
            s = idx >= 0 ? ld_relaxed(&state[idx]) : ST_INC;   // before tile 0: an inclusive prefix of 0
        } while (__any_sync(0xffffffffu, (s & ST_FLAGS) == 0));
        const u32 inc = __ballot_sync(0xffffffffu, (s & ST_FLAGS) == ST_INC);
        u64 v = s & ST_VAL;
        if (inc) {
            const int first = __ffs(inc) - 1;   // nearest predecessor that already knows its inclusive prefix
            v = lane <= first ? v : 0;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        exclusive += v;
        if (inc) break;
        look -= 32;
    }
    if (lane == 0) st_relaxed(&state[tile], ST_INC | (exclusive + block_total));
    return exclusive;
}

struct TileCtl {
    u64 *state;     // one status word per tile (zeroed before the launch)
    i64 *total;     // device-visible: receives the number of outputs (written by the last tile)
    u32 tiles;
};

template <int NWARPS> struct TileSmemT {
    u64 warp[NWARPS];
    u64 prefix;
};
typedef TileSmemT<WARPS> TileSmem;

// warp totals -> CTA offsets + global prefix.  Returns the global output offset of this warp's first selected row.
template <int NWARPS>
__device__ __forceinline__ u64 tile_offsets(const TileCtl &ctl, u32 tile, u32 warp_total, TileSmemT<NWARPS> &sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sm.warp[warp] = warp_total;
    __syncthreads();
    if (warp == 0) {
        u64 t = lane < NWARPS ? sm.warp[lane] : 0, incl = t;
#pragma unroll
        for (int d = 1; d < NWARPS; d <<= 1) {
            const u64 o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        const u64 block_total = __shfl_sync(0xffffffffu, incl, NWARPS - 1);
        if (lane < NWARPS) sm.warp[lane] = incl - t;   // exclusive offset of each warp inside the tile
        const u64 excl = lookback(ctl.state, tile, block_total);
        if (lane == 0) {
            sm.prefix = excl;
            if (tile == ctl.tiles - 1) { *ctl.total = (i64)(excl + block_total); __threadfence_system(); }
        }
    }
    __syncthreads();
    return sm.prefix + sm.warp[warp];
}

// Generic row compaction, one row per lane per step: rows [0, n); flag(row) decides, emit(row, index) consumes the
// selected rows in ascending row order with their dense output index.
template <int J> struct RowTile { static constexpr int WROWS = 32 * J, TILE = WARPS * WROWS; };

template <int J, typename FlagFn, typename EmitFn>
__device__ __forceinline__ void compact_rows(i64 n, const TileCtl &ctl, TileSmem &sm, FlagFn flag, EmitFn emit) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 tile = blockIdx.x;
    const i64 wbase = (i64)tile * RowTile<J>::TILE + (i64)warp * RowTile<J>::WROWS;
    bool f[J];
#pragma unroll
    for (int j = 0; j < J; j++) {
        const i64 r = wbase + j * 32 + lane;
        f[j] = r < n && flag(r);
    }
    u32 excl[J], warp_total = 0;
#pragma unroll
    for (int j = 0; j < J; j++) {
        const u32 m = __ballot_sync(0xffffffffu, f[j]);
        excl[j] = warp_total + __popc(m & ((1u << lane) - 1u));
        warp_total += __popc(m);
    }
    const u64 obase = tile_offsets(ctl, tile, warp_total, sm);
#pragma unroll
    for (int j = 0; j < J; j++)
        if (f[j]) emit(wbase + j * 32 + lane, (i64)(obase + excl[j]));
}

// host: carve (ticket, state[tiles]) out of the context workspace at `offset` bytes and zero it
static inline int prepare_tiles(rfb_ctx_t *ctx, void *work, i64 tiles, i64 *total, TileCtl *ctl) {
    const size_t bytes = (size_t)tiles * 8 + 64;
    RFB_CUDA(cudaMemsetAsync(work, 0, bytes, ctx->stream));
    ctl->state = (u64 *)((char *)work + 64);
    ctl->total = total;
    ctl->tiles = (u32)tiles;
    return RFB_OK;
}
static inline size_t tiles_bytes(i64 tiles) { return (((size_t)tiles * 8 + 64) + 255) & ~(size_t)255; }

}  // namespace scan
