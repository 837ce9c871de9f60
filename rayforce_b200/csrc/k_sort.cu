// k_sort.cu — key sort (sm_100a): stable LSD radix sort returning the i64 permutation.
//
//   rfb_sort_dev   ray_sort_asc / ray_sort_desc (reference core/sort.c:183-428 ascending, :481-689 descending; single
//                  threaded there: counting sort for 8/16-bit keys, 16-bit-digit LSD radix for 32/64-bit keys)
//
// Any stable sort on the reference's key order yields the same permutation, so the device is free to pick its own digit
// width: 8-bit digits, least significant first.  Keys are mapped to unsigned integers exactly like the reference
// (core/sort.c:266-285,313): integers flip the sign bit (nulls first), doubles flip sign / all bits with every NaN -> 0
// (NaN first, -0.0 before +0.0); descending sorts the complemented key, which keeps equal keys in original order.
//
// Per pass:  HIST    G persistent CTAs, CTA b owns the contiguous chunk b of the current sequence: 256-bin counts
//            SCAN    exclusive scan of the digit-major (256 x G) count matrix  -> first output slot of (digit, chunk)
//            SCATTER CTA b walks its chunk tile by tile (3584 rows): warp-level multisplit (per-bit ballots) ranks rows of equal
//                    digit in (warp, step, lane) order, warp counts are scanned across the CTA, the tile is ordered by
//                    digit in shared memory and leaves as one contiguous run per digit; a running per-digit base carried
//                    in shared memory keeps tiles of one chunk in order => stable.
// One census kernel up front takes the bitwise OR and AND of all keys; a pass whose digit is constant over the column is skipped
// (e.g. the upper bytes of small integers).  The first executed pass reads the typed column and synthesises the row ids;
// the last one writes only the permutation.  HBM traffic per executed pass: 8N (hist) + 16N read + 16N written.
#include "rfb_tma.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
// scatter geometry, measured on B200 (1e8 i64 rows, 8 passes; ITEMS x CTAs/SM x early row-id loads): 12x3x0 13.3 ms (before the
// cross-tile pipeline), 12x3x1 14.1 (row ids spill), 12x2x0 12.8, 12x2x1 11.7, 14x2x1 11.3, 16x2x1 11.7, 18x2x1 12.2;
// 512-thread CTAs (RFB_SORT_THREADS=512, 64 registers, twice the warps per SM): 7x2 11.5, 6x2 11.8, 8x2 12.2 — no gain
#ifndef RFB_SORT_EARLY_RID
#define RFB_SORT_EARLY_RID 1
#endif
#ifndef RFB_SORT_ITEMS
#define RFB_SORT_ITEMS 14
#define RFB_SORT_CTAS 2
#endif
#ifndef RFB_SORT_THREADS
#define RFB_SORT_THREADS 256
#endif
constexpr int SCT = RFB_SORT_THREADS, SWARPS = SCT / 32;   // scatter kernel CTA size (the other kernels use THREADS)
constexpr int ITEMS = RFB_SORT_ITEMS;
constexpr int TILE = SCT * ITEMS;  // 3584 rows: 56 KB of staged (key, row id) pairs per CTA
constexpr int RADIX = 256;

template <typename T> __device__ __forceinline__ u64 sortable(T v);
template <> __device__ __forceinline__ u64 sortable<u8>(u8 v) { return v; }
template <> __device__ __forceinline__ u64 sortable<i16>(i16 v) { return (u64)(unsigned short)((unsigned short)v ^ 0x8000u); }
template <> __device__ __forceinline__ u64 sortable<i32>(i32 v) { return (u64)((u32)v ^ 0x80000000u); }
template <> __device__ __forceinline__ u64 sortable<i64>(i64 v) { return (u64)v ^ 0x8000000000000000ULL; }
template <> __device__ __forceinline__ u64 sortable<f64>(f64 v) { return f64_sort_key(v); }

// where a pass reads its (key, row id) pairs from
template <typename T> struct ColumnSrc {      // first pass: the typed column itself
    const T *col;
    u64 flip;  // 0 ascending, all-ones (within the key width) descending
    __device__ __forceinline__ u64 key(i64 i) const { return sortable<T>(ld_stream(col + i)) ^ flip; }
    __device__ __forceinline__ i64 val(i64 i) const { return i; }
};
struct PairSrc {                              // later passes: the previous pass's output
    const u64 *keys;
    const i64 *vals;
    __device__ __forceinline__ u64 key(i64 i) const { return ld_stream(keys + i); }
    __device__ __forceinline__ i64 val(i64 i) const { return ld_stream(vals + i); }
};

// ---- census: bitwise OR and AND of all keys.  A pass whose digit is the same in every key (OR byte == AND byte) leaves
// the sequence unchanged and is skipped.
template <typename T>
__global__ void __launch_bounds__(THREADS) k_census(ColumnSrc<T> src, i64 n, unsigned long long *or_and) {
    u64 o = 0, a = ~0ULL;
    constexpr int U = 4;
    const i64 stride = (i64)gridDim.x * THREADS;
    i64 i = (i64)blockIdx.x * THREADS + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        u64 k[U];
#pragma unroll
        for (int j = 0; j < U; j++) k[j] = src.key(i + j * stride);
#pragma unroll
        for (int j = 0; j < U; j++) { o |= k[j]; a &= k[j]; }
    }
    for (; i < n; i += stride) { const u64 k = src.key(i); o |= k; a &= k; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        o |= __shfl_xor_sync(0xffffffffu, o, d);
        a &= __shfl_xor_sync(0xffffffffu, a, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicOr(&or_and[0], (unsigned long long)o);
        atomicAnd(&or_and[1], (unsigned long long)a);
    }
}

// ---- per-chunk digit histogram
template <typename Src>
__global__ void __launch_bounds__(THREADS) k_hist(Src src, i64 n, i64 chunk, int shift, u32 *bh /* [256][G] */) {
    __shared__ u32 h[RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    const i64 lo = (i64)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
    for (i64 i = lo + threadIdx.x; i < hi; i += THREADS) atomicAdd(&h[(src.key(i) >> shift) & 255], 1u);
    __syncthreads();
    bh[(i64)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// ---- exclusive scan of the (256 x G) matrix in row-major (digit-major) order; one CTA
__global__ void __launch_bounds__(1024) k_scan_counts(const u32 *bh, i64 *offs, int total) {
    __shared__ i64 wsum[32];
    __shared__ i64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < total; base += 1024) {
        const int i = base + threadIdx.x;
        const i64 v = i < total ? (i64)bh[i] : 0;
        i64 incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const i64 o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            i64 w = wsum[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const i64 o = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += o;
            }
            wsum[lane] = wi - w;
        }
        __syncthreads();
        const i64 c = carry;
        if (i < total) offs[i] = c + wsum[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + wsum[31] + incl;
        __syncthreads();
    }
}

// ---- stable scatter of one chunk.  Each tile is first ordered by digit in shared memory (stable: rank = rows of the same
// digit in lower warps + lower steps/lanes of the own warp), then streamed out so that the rows of one digit leave as one
// contiguous run — direct per-lane stores hit up to 32 different sectors per instruction.
template <typename Src, bool WRITE_KEYS>
__global__ void __launch_bounds__(SCT, RFB_SORT_CTAS)
k_scatter(Src src, i64 n, i64 chunk, int shift, const i64 *__restrict__ offs /* [256][G] */, u64 *__restrict__ keys_out,
          i64 *__restrict__ vals_out) {
    __shared__ u32 whist[SWARPS][RADIX];
    __shared__ i64 base[RADIX];     // next free output slot of each digit for this chunk
    __shared__ u32 dstart[RADIX];   // tile-local position of the first row of each digit
    __shared__ u32 wsum[RADIX / 32];
    extern __shared__ u64 stage_dyn[];            // TILE keys then TILE row ids (48 KB: above the static limit)
    u64 *skeys = stage_dyn;
    i64 *svals = (i64 *)(stage_dyn + TILE);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1u;
    if (threadIdx.x < RADIX) base[threadIdx.x] = offs[(i64)threadIdx.x * gridDim.x + blockIdx.x];
    const i64 lo = (i64)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
    // software pipeline across tiles, without extra registers: a tile's keys are dead once they sit in shared memory, so the
    // NEXT tile's keys are loaded into the same registers right before the write-out (their latency hides behind it), and the
    // row ids of the current tile are requested right after the ranking, before the two barriers of the digit scan
    u64 key[ITEMS];
    {
        const i64 wb0 = lo + (i64)warp * (32 * ITEMS);
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const i64 i = wb0 + j * 32 + lane;
            key[j] = i < hi ? src.key(i) : ~0ULL;
        }
    }
    for (i64 t0 = lo; t0 < hi; t0 += TILE) {
        for (int idx = threadIdx.x; idx < SWARPS * RADIX; idx += SCT) (&whist[0][0])[idx] = 0;
        __syncthreads();
        u32 rank[ITEMS];
        const i64 wb = t0 + (i64)warp * (32 * ITEMS);
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const i64 i = wb + j * 32 + lane;
            const bool ok = i < hi;
            const u32 d = (u32)(key[j] >> shift) & 255u;
            const u32 vmask = __ballot_sync(0xffffffffu, ok);
            // lanes holding the same digit: eight ballots (one per digit bit) instead of __match_any_sync — MATCH runs on the
            // ADU pipe at ~17 cycles per warp instruction on B200 and bounded this kernel; VOTE + LOP3 issue at full rate
            u32 peers = vmask;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const bool bit = (d >> b) & 1u;
                const u32 bal = __ballot_sync(0xffffffffu, bit);
                peers &= bit ? bal : ~bal;
            }
            const u32 prior = whist[warp][d];
            __syncwarp();
            if (ok && (peers & lt) == 0) whist[warp][d] = prior + __popc(peers);   // lowest peer lane publishes
            __syncwarp();
            rank[j] = prior + __popc(peers & lt);
        }
#if RFB_SORT_EARLY_RID
        i64 rid[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const i64 i = wb + j * 32 + lane;
            rid[j] = i < hi ? src.val(i) : 0;
        }
#endif
        __syncthreads();
        u32 cnt = 0, incl = 0;
        if (threadIdx.x < RADIX) {   // thread d: digit d's counts over the warps -> exclusive warp offsets; then the tile-local start of each digit
            const int d = threadIdx.x;
            u32 sum = 0;
#pragma unroll
            for (int w = 0; w < SWARPS; w++) { const u32 c = whist[w][d]; whist[w][d] = sum; sum += c; }
            cnt = sum;
            incl = cnt;
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) {
                const u32 o = __shfl_up_sync(0xffffffffu, incl, k);
                if (lane >= k) incl += o;
            }
            if (lane == 31) wsum[warp] = incl;
        }
        __syncthreads();
        if (threadIdx.x < RADIX) {
            u32 woff = 0;
#pragma unroll
            for (int w = 0; w < RADIX / 32; w++) woff += (w < warp) ? wsum[w] : 0u;
            dstart[threadIdx.x] = woff + incl - cnt;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const i64 i = wb + j * 32 + lane;
            if (i < hi) {
                const u32 d = (u32)(key[j] >> shift) & 255u;
                const u32 lp = dstart[d] + whist[warp][d] + rank[j];
                skeys[lp] = key[j];
#if RFB_SORT_EARLY_RID
                svals[lp] = rid[j];
#else
                svals[lp] = src.val(i);
#endif
            }
        }
        __syncthreads();
        if (t0 + TILE < hi) {      // next tile's keys
            const i64 nb = wb + TILE;
#pragma unroll
            for (int j = 0; j < ITEMS; j++) {
                const i64 i = nb + j * 32 + lane;
                key[j] = i < hi ? src.key(i) : ~0ULL;
            }
        }
        const int tile_n = (int)((hi - t0) < TILE ? (hi - t0) : TILE);
        for (int i = threadIdx.x; i < tile_n; i += SCT) {
            const u64 k = skeys[i];
            const u32 d = (u32)(k >> shift) & 255u;
            const i64 pos = base[d] + (i64)(i - (int)dstart[d]);
            if (WRITE_KEYS) keys_out[pos] = k;
            vals_out[pos] = svals[i];
        }
        __syncthreads();
        if (threadIdx.x < RADIX) base[threadIdx.x] += cnt;
    }
}

__global__ void k_iota(i64 *p, i64 n) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) p[i] = i;
}

inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

template <typename Src>
int run_pass(rfb_ctx_t *ctx, Src src, i64 n, int G, i64 chunk, int shift, u32 *bh, i64 *offs, u64 *keys_out, i64 *vals_out, bool last) {
    k_hist<Src><<<G, THREADS, 0, ctx->stream>>>(src, n, chunk, shift, bh);
    RFB_CHECK_LAUNCH(ctx);
    k_scan_counts<<<1, 1024, 0, ctx->stream>>>(bh, offs, RADIX * G);
    RFB_CHECK_LAUNCH(ctx);
    constexpr int STAGE_BYTES = TILE * 16;
    static bool opted_in = false;   // per template instantiation
    if (!opted_in) {
        RFB_CUDA(cudaFuncSetAttribute(k_scatter<Src, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE_BYTES));
        RFB_CUDA(cudaFuncSetAttribute(k_scatter<Src, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE_BYTES));
        opted_in = true;
    }
    if (last) k_scatter<Src, false><<<G, SCT, STAGE_BYTES, ctx->stream>>>(src, n, chunk, shift, offs, keys_out, vals_out);
    else k_scatter<Src, true><<<G, SCT, STAGE_BYTES, ctx->stream>>>(src, n, chunk, shift, offs, keys_out, vals_out);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

// =====================================================================================================================
// Single-sweep passes ("onesweep"): what sorts run on for n < 2^32 rows.
//   k_os_hist   ONE read of the typed column: the 256-bin histogram of every key byte (global digit counts of all passes).
//   k_os_scan   exclusive scan per pass -> where each digit's run starts in that pass's output.
//   k_os_pass   per executed pass: tiles are claimed in order from an atomic counter; a tile ranks its rows (ballot multisplit),
//               publishes its 256 digit counts as (tag | AGGREGATE | count) words, resolves its exclusive prefix per digit by
//               decoupled look-back over the preceding tiles' words (thread d chases digit d) while the other warps already
//               order the tile by digit in shared memory, publishes (INCLUSIVE | prefix + count) and streams the tile out as
//               one contiguous run per digit.  No per-pass histogram kernel, no (256 x chunks) offset matrix: a pass reads
//               its input once.  Tiles leave in tile order within a digit and rows keep their order inside a tile => stable.
// Row ids travel as 32-bit words between passes (12 B per row moved instead of 16); the last pass widens them into `perm`.
// A pass whose digit is constant over the column (one histogram bin holds every row) is skipped, as before.
// Per executed pass: 12N read + 12N written (first pass: sizeof(T) N read; last pass: 8N written).
#ifndef RFB_OS_MATCH
#define RFB_OS_MATCH 0
#endif
#ifndef RFB_OS_LB
#define RFB_OS_LB 4
#endif
#ifndef RFB_OS_CTAS32
#define RFB_OS_CTAS32 2       // CTAs per SM of the passes over 32-bit key words (16 x 3 / 20 x 2 / 24 x 2 / 20 x 3 / 12 x 4: i32 3.66 / 3.46 / 3.26 / 3.65 / 4.41 ms per 1e8 rows)
#endif
#ifndef RFB_OS_SLEEP
#define RFB_OS_SLEEP 0
#endif

#ifndef RFB_OS_STAGE_RIDS
#define RFB_OS_STAGE_RIDS 1
#endif
constexpr int OS_LB = RFB_OS_LB;
#ifndef RFB_OS_ITEMS
#define RFB_OS_ITEMS 16
#define RFB_OS_CTAS 2
#endif
constexpr int OS_T = 256, OS_W = OS_T / 32;
#ifndef RFB_OS_ITEMS32
#define RFB_OS_ITEMS32 24     // 32-bit key words: 6144-row tiles (96 KB of stages, 2 CTAs per SM)
#endif
// rows per thread (tile = 256 x ITEMS rows) by key word
template <typename K> struct OsGeom { static constexpr int ITEMS = sizeof(K) == 4 ? RFB_OS_ITEMS32 : RFB_OS_ITEMS, TILE = OS_T * ITEMS; };
constexpr u64 OS_AGG = 1ULL << 54, OS_INC = 2ULL << 54, OS_FLAGS = 3ULL << 54, OS_COUNT = (1ULL << 54) - 1;

// the key word a column travels as between passes: 32 bits for columns of up to 4 bytes (8 B per row moved, three CTAs per SM)
template <typename T> struct OsKey { typedef u64 type; };
template <> struct OsKey<u8> { typedef u32 type; };
template <> struct OsKey<i16> { typedef u32 type; };
template <> struct OsKey<i32> { typedef u32 type; };

// first pass: the typed column; the row id is the row number.  K: the key word (OsKey<T>, or u32 for an 8-byte column whose
// varying bytes fit a 32-bit window: `base_shift` drops the constant bytes below it, the constant bytes above fall off the cast)
template <typename T, typename K> struct OsColumnSrc {
    typedef T raw_t;
    typedef K key_t;
    static constexpr bool HAS_RIDS = false;
    const T *col;
    u64 flip;
    int base_shift;
    __host__ __device__ __forceinline__ const T *raw(i64 i) const { return col + i; }
    __device__ __forceinline__ K key_of(T v) const { return (K)((sortable<T>(v) ^ flip) >> base_shift); }
    __device__ __forceinline__ K key(i64 i) const { return key_of(ld_stream(col + i)); }
    __device__ __forceinline__ u64 key64(i64 i) const { return sortable<T>(ld_stream(col + i)) ^ flip; }
    __device__ __forceinline__ u32 rid(i64 i) const { return (u32)i; }
};
template <typename K> struct OsPairSrc {        // later passes: the previous pass's (key, 32-bit row id) pairs
    typedef K raw_t;
    typedef K key_t;
    static constexpr bool HAS_RIDS = RFB_OS_STAGE_RIDS != 0;
    const K *keys;
    const u32 *rids;
    __host__ __device__ __forceinline__ const K *raw(i64 i) const { return keys + i; }
    __host__ __device__ __forceinline__ const u32 *raw_rids(i64 i) const { return rids + i; }
    __device__ __forceinline__ K key_of(K v) const { return v; }
    __device__ __forceinline__ K key(i64 i) const { return ld_stream(keys + i); }
    __device__ __forceinline__ u32 rid(i64 i) const { return ld_stream(rids + i); }
};

template <typename T>
__global__ void __launch_bounds__(THREADS, 4) k_os_hist(OsColumnSrc<T, u64> src, i64 n, unsigned long long *ghist /* [NPASS][256] */) {
    constexpr int NPASS = (int)sizeof(T);
    __shared__ u32 h[NPASS][RADIX];
    for (int i = threadIdx.x; i < NPASS * RADIX; i += THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    constexpr int U = 4;
    const i64 stride = (i64)gridDim.x * THREADS;
    // whole warps stay in the loop (the votes below need all 32 lanes): iterate on the warp's first row
    for (i64 w0 = (i64)blockIdx.x * THREADS + (threadIdx.x & ~31); w0 < n; w0 += U * stride) {
        u64 k[U];
        bool ok[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            const i64 i = w0 + j * stride + (threadIdx.x & 31);
            ok[j] = i < n;
            k[j] = ok[j] ? src.key64(i) : 0;
        }
#pragma unroll
        for (int j = 0; j < U; j++) {
#pragma unroll
            for (int p = 0; p < NPASS; p++) {
                const u32 d = (u32)(k[j] >> (8 * p)) & 255u;
                // a byte that is the same in all 32 rows (the upper bytes of small integers) would be a 32-way same-address atomic
                const u32 d0 = __shfl_sync(0xffffffffu, d, 0);
                if (__all_sync(0xffffffffu, ok[j] && d == d0)) { if ((threadIdx.x & 31) == 0) atomicAdd(&h[p][d], 32u); }
                else if (ok[j]) atomicAdd(&h[p][d], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NPASS * RADIX; i += THREADS) {
        const u32 c = (&h[0][0])[i];
        if (c) atomicAdd(&ghist[i], (unsigned long long)c);
    }
}

__global__ void __launch_bounds__(RADIX) k_os_scan(const unsigned long long *ghist, i64 *gbase) {
    __shared__ i64 wtot[RADIX / 32];
    const int d = threadIdx.x, lane = d & 31, warp = d >> 5;
    const i64 c = (i64)ghist[blockIdx.x * RADIX + d];
    i64 incl = c;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        const i64 o = __shfl_up_sync(0xffffffffu, incl, k);
        if (lane >= k) incl += o;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    i64 woff = 0;
    for (int w = 0; w < warp; w++) woff += wtot[w];
    gbase[blockIdx.x * RADIX + d] = woff + incl - c;
}

template <typename Src, bool LAST>
__global__ void __launch_bounds__(OS_T, (sizeof(typename Src::key_t) == 4 ? RFB_OS_CTAS32 : RFB_OS_CTAS))
k_os_pass(Src src, i64 n, u32 tiles, int shift, u64 tag, const i64 *__restrict__ gbase /* [256] */, unsigned long long *__restrict__ status /* [tiles][256] */,
          u32 *__restrict__ tile_counter, typename Src::key_t *__restrict__ keys_out, u32 *__restrict__ rids_out, i64 *__restrict__ perm_out, bool staged) {
    typedef typename Src::raw_t raw_t;
    typedef typename Src::key_t K;
    constexpr int OS_ITEMS = OsGeom<K>::ITEMS, OS_TILE = OsGeom<K>::TILE;
    __shared__ u32 whist[OS_W][RADIX];
    __shared__ i64 base[RADIX];     // output slot of the tile-local position 0 of each digit's run
    __shared__ u32 wsum[RADIX / 32];
    __shared__ u32 s_tile[2];
    __shared__ __align__(8) u64 full;             // mbarrier: the staged keys of the NEXT tile have landed
    extern __shared__ __align__(16) unsigned char stage_raw[];   // OS_TILE sorted keys | staged raw keys | sorted 32-bit row ids | staged row ids
    K *skeys = (K *)stage_raw;
    const raw_t *inkeys = (const raw_t *)(stage_raw + OS_TILE * sizeof(K));
    u32 *srids = (u32 *)(stage_raw + OS_TILE * (sizeof(K) + sizeof(raw_t)));
    const u32 *inrids = srids + OS_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1u;
    // A tile's keys are staged by the TMA unit (one cp.async.bulk, completion counted on `full`) while the PREVIOUS tile is
    // ranked, ordered and written out: thread 0 claims the next tile and issues its copy as soon as every warp has moved the
    // current tile's keys from the stage into registers.  Only whole tiles of a 16-byte aligned source are staged.
    if (threadIdx.x == 0) {
        mbar_init(&full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const u32 t = atomicAdd(tile_counter, 1u);
        s_tile[0] = t;
        if (staged && t < tiles && (i64)(t + 1) * OS_TILE <= n) {
            mbar_expect_tx(&full, OS_TILE * (u32)(sizeof(raw_t) + (Src::HAS_RIDS ? 4 : 0)));
            bulk_g2s((void *)inkeys, src.raw((i64)t * OS_TILE), OS_TILE * (u32)sizeof(raw_t), &full);
            if constexpr (Src::HAS_RIDS) bulk_g2s((void *)inrids, src.raw_rids((i64)t * OS_TILE), OS_TILE * 4u, &full);
        }
    }
    u32 phase = 0;
    for (u32 it = 0;; it++) {
        for (int idx = threadIdx.x; idx < OS_W * RADIX; idx += OS_T) (&whist[0][0])[idx] = 0;
        __syncthreads();
        const u32 tile = s_tile[it & 1];
        if (tile >= tiles) break;
        const i64 t0 = (i64)tile * OS_TILE;
        const i64 wb = t0 + (i64)warp * (32 * OS_ITEMS);
        K key[OS_ITEMS];
        u32 rid[OS_ITEMS];
        if (staged && t0 + OS_TILE <= n) {
            mbar_wait(&full, phase & 1u);
            phase++;
#pragma unroll
            for (int j = 0; j < OS_ITEMS; j++) {
                const int li = warp * (32 * OS_ITEMS) + j * 32 + lane;
                key[j] = src.key_of(inkeys[li]);
                if constexpr (Src::HAS_RIDS) rid[j] = inrids[li];
                else rid[j] = src.rid(t0 + li);
            }
        } else {
#pragma unroll
            for (int j = 0; j < OS_ITEMS; j++) {
                const i64 i = wb + j * 32 + lane;
                key[j] = i < n ? src.key(i) : (K)~0ULL;
                rid[j] = i < n ? src.rid(i) : 0u;
            }
        }
        __syncthreads();                          // the stage is free again
        if (threadIdx.x == 0) {
            const u32 t = atomicAdd(tile_counter, 1u);
            s_tile[(it + 1) & 1] = t;
            if (staged && t < tiles && (i64)(t + 1) * OS_TILE <= n) {
                mbar_expect_tx(&full, OS_TILE * (u32)(sizeof(raw_t) + (Src::HAS_RIDS ? 4 : 0)));
                bulk_g2s((void *)inkeys, src.raw((i64)t * OS_TILE), OS_TILE * (u32)sizeof(raw_t), &full);
                if constexpr (Src::HAS_RIDS) bulk_g2s((void *)inrids, src.raw_rids((i64)t * OS_TILE), OS_TILE * 4u, &full);
            }
        }
        // rank of a row among the rows of its warp with the same digit, in (step, lane) order: the lanes holding the same digit
        // come from eight ballots; the lowest of them adds the group's size to the warp's counter with a returning atomic (the
        // counter's value before = rows of that digit in earlier steps) and hands the answer to its peers with a shuffle.
        // Measured alternatives (1e8 i64 rows, same box): match.any instead of the ballots 10.0 vs 8.5 ms (ADU pipe); the votes
        // of 2 / 4 rows interleaved by hand with all atomics issued before the first shuffle 8.20 / 8.22 vs 7.78 ms.
        u32 rank[OS_ITEMS];
#pragma unroll
        for (int j = 0; j < OS_ITEMS; j++) {
            const i64 i = wb + j * 32 + lane;
            const bool ok = i < n;
            const u32 d = (u32)(key[j] >> shift) & 255u;
            u32 peers = __ballot_sync(0xffffffffu, ok);
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const bool bit = (d >> b) & 1u;
                const u32 bal = __ballot_sync(0xffffffffu, bit);
                peers &= bit ? bal : ~bal;
            }
            const int leader = ok ? __ffs(peers) - 1 : lane;
            u32 prior = 0;
            if (ok && lane == leader) prior = atomicAdd(&whist[warp][d], (u32)__popc(peers));
            prior = __shfl_sync(0xffffffffu, prior, leader);
            rank[j] = prior + __popc(peers & lt);
        }
        __syncthreads();
        // thread d: digit d's counts over the warps -> exclusive warp offsets, tile count, tile-local start of the digit's run
        const int d = threadIdx.x;
        u32 cnt = 0;
#pragma unroll
        for (int w = 0; w < OS_W; w++) { const u32 c = whist[w][d]; whist[w][d] = cnt; cnt += c; }
        u32 incl = cnt;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const u32 o = __shfl_up_sync(0xffffffffu, incl, k);
            if (lane >= k) incl += o;
        }
        if (lane == 31) wsum[warp] = incl;
        unsigned long long *mine = status + (size_t)tile * RADIX + d;
        *(volatile unsigned long long *)mine = tag | (tile == 0 ? OS_INC : OS_AGG) | (u64)cnt;
        __syncthreads();
        u32 woff = 0;
#pragma unroll
        for (int w = 0; w < RADIX / 32; w++) woff += (w < warp) ? wsum[w] : 0u;
        const u32 ds = woff + incl - cnt;          // tile-local start of digit d's run
#pragma unroll
        for (int w = 0; w < OS_W; w++) whist[w][d] += ds;     // tile-local slot of the first row of (warp w, digit d)
        __syncthreads();
        // order the tile by digit in shared memory ...
#pragma unroll
        for (int j = 0; j < OS_ITEMS; j++) {
            const i64 i = wb + j * 32 + lane;
            if (i < n) {
                const u32 dj = (u32)(key[j] >> shift) & 255u;
                const u32 lp = whist[warp][dj] + rank[j];
                skeys[lp] = key[j];
                srids[lp] = rid[j];
            }
        }
        // ... then resolve digit d's exclusive prefix over the preceding tiles (their words have had the whole ordering step to land)
        u64 excl = 0;
        if (tile > 0) {
            // OS_LB predecessor words per round trip (independent loads), consumed nearest first; a word that is not there yet
            // ends the batch and the walk resumes from it
            i64 pt = (i64)tile - 1;
            for (bool done = false; !done;) {
                u64 v[OS_LB];
#pragma unroll
                for (int k = 0; k < OS_LB; k++)
                    v[k] = pt - k >= 0 ? *(const volatile unsigned long long *)(status + (size_t)(pt - k) * RADIX + d) : (tag | OS_INC);
                int used = 0;
#pragma unroll
                for (int k = 0; k < OS_LB; k++) {
                    if (done || used != k) continue;
                    if ((v[k] & ~(OS_FLAGS | OS_COUNT)) != tag || (v[k] & OS_FLAGS) == 0) continue;   // not published in this pass yet
                    excl += v[k] & OS_COUNT;
                    used = k + 1;
                    done = (v[k] & OS_FLAGS) == OS_INC;
                }
                pt -= used;
#if RFB_OS_SLEEP > 0
                if (!done && used == 0) __nanosleep(RFB_OS_SLEEP);   // measured: no effect (0 / 100 / 400 ns: 8.29 / 8.32 / 8.34 ms)
#endif
            }
            *(volatile unsigned long long *)mine = tag | OS_INC | (excl + cnt);
        }
        base[d] = gbase[d] + (i64)excl - (i64)ds;
        __syncthreads();
        const int tile_n = (int)((n - t0) < OS_TILE ? (n - t0) : OS_TILE);
#pragma unroll 4
        for (int i = threadIdx.x; i < tile_n; i += OS_T) {
            const K k = skeys[i];
            const i64 pos = base[(u32)(k >> shift) & 255u] + i;
            if (LAST) perm_out[pos] = (i64)srids[i];
            else { keys_out[pos] = k; rids_out[pos] = srids[i]; }
        }
        // the next round's first barrier (after whist is zeroed) also orders these reads of the sorted stage before its next writes
    }
}

template <typename Src>
int os_run_pass(rfb_ctx_t *ctx, Src src, i64 n, u32 tiles, int shift, u64 tag, const i64 *gbase, unsigned long long *status, u32 *counter,
                typename Src::key_t *keys_out, u32 *rids_out, i64 *perm_out, bool last) {
    constexpr int KB = (int)sizeof(typename Src::key_t), RB = (int)sizeof(typename Src::raw_t), OS_TILE = OsGeom<typename Src::key_t>::TILE;
    constexpr int STAGE_BYTES = OS_TILE * (KB + RB + (Src::HAS_RIDS ? 8 : 4));   // sorted keys + staged raw keys + sorted row ids (+ staged row ids)
    constexpr int CTAS = KB == 4 ? RFB_OS_CTAS32 : RFB_OS_CTAS;
    const bool staged = aligned16(src.raw(0));
    static bool opted_in = false;   // per template instantiation
    if (!opted_in) {
        RFB_CUDA(cudaFuncSetAttribute(k_os_pass<Src, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE_BYTES));
        RFB_CUDA(cudaFuncSetAttribute(k_os_pass<Src, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE_BYTES));
        opted_in = true;
    }
    const u32 resident = (u32)ctx->sm_count * CTAS;
    const u32 grid = tiles < resident ? tiles : resident;
    if (last) k_os_pass<Src, true><<<grid, OS_T, STAGE_BYTES, ctx->stream>>>(src, n, tiles, shift, tag, gbase, status, counter, keys_out, rids_out, perm_out, staged);
    else k_os_pass<Src, false><<<grid, OS_T, STAGE_BYTES, ctx->stream>>>(src, n, tiles, shift, tag, gbase, status, counter, keys_out, rids_out, perm_out, staged);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

// the executed passes over key words of type K (`base`: the key byte that became byte 0 of the word)
template <typename T, typename K>
int os_passes(rfb_ctx_t *ctx, const T *x, u64 flip, i64 n, const int *passes, int np, int base, const i64 *gbase, unsigned long long *status,
              u32 *counters, void *keys_a, void *keys_b, u32 *ridsA, u32 *ridsB, i64 *perm) {
    const u32 tiles = (u32)((n + OsGeom<K>::TILE - 1) / OsGeom<K>::TILE);
    OsColumnSrc<T, K> col{x, flip, 8 * base};
    K *keysA = (K *)keys_a, *keysB = (K *)keys_b;
    K *kin = nullptr, *kout = keysA;
    u32 *rin = nullptr, *rout = ridsA;
    for (int q = 0; q < np; q++) {
        const bool last = (q == np - 1);
        const int shift = 8 * (passes[q] - base);
        const u64 tag = (u64)(q + 1) << 56;
        const i64 *gb = gbase + passes[q] * RADIX;
        int rc;
        if (q == 0) rc = os_run_pass(ctx, col, n, tiles, shift, tag, gb, status, counters + q, kout, rout, perm, last);
        else rc = os_run_pass(ctx, OsPairSrc<K>{kin, rin}, n, tiles, shift, tag, gb, status, counters + q, kout, rout, perm, last);
        if (rc) return rc;
        kin = kout; kout = (kout == keysA) ? keysB : keysA;
        rin = rout; rout = (rout == ridsA) ? ridsB : ridsA;
    }
    return RFB_OK;
}

template <typename T>
int sort_onesweep(rfb_ctx_t *ctx, const void *x, i64 n, int descending, i64 *perm) {
    constexpr int NPASS = (int)sizeof(T);
    const u64 width_mask = NPASS == 8 ? ~0ULL : ((1ULL << (8 * NPASS)) - 1);
    typedef typename OsKey<T>::type K;
    const u64 flip = descending ? width_mask : 0ULL;
    constexpr int MIN_TILE = OS_T * (RFB_OS_ITEMS < RFB_OS_ITEMS32 ? RFB_OS_ITEMS : RFB_OS_ITEMS32);
    const u32 tiles = (u32)((n + MIN_TILE - 1) / MIN_TILE);             // status words: sized for the smaller tile
    // workspace: ghist[8][256] u64 | gbase[8][256] i64 | counters[8] u32 | status[tiles][256] u64 | keysA[n] | keysB[n] | ridsA[n] | ridsB[n]
    const size_t b_hist = 8 * RADIX * 8, b_base = 8 * RADIX * 8, b_cnt = 256, b_status = align256((size_t)tiles * RADIX * 8),
                 b_k = align256((size_t)n * sizeof(K)), b_r = align256((size_t)n * 4);
    void *w;
    int rc = rfb_ensure_work(ctx, b_hist + b_base + b_cnt + b_status + 2 * b_k + 2 * b_r, &w);
    if (rc) return rc;
    char *p = (char *)w;
    unsigned long long *ghist = (unsigned long long *)p; p += b_hist;
    i64 *gbase = (i64 *)p; p += b_base;
    u32 *counters = (u32 *)p; p += b_cnt;
    unsigned long long *status = (unsigned long long *)p; p += b_status;
    void *keysA = p; p += b_k;
    void *keysB = p; p += b_k;
    u32 *ridsA = (u32 *)p; p += b_r;
    u32 *ridsB = (u32 *)p;
    RFB_CUDA(cudaMemsetAsync(w, 0, b_hist + b_base + b_cnt + b_status, ctx->stream));
    k_os_hist<T><<<rfb_grid_for(ctx, n, THREADS * 4, 4), THREADS, 0, ctx->stream>>>(OsColumnSrc<T, u64>{(const T *)x, flip, 0}, n, ghist);
    RFB_CHECK_LAUNCH(ctx);
    k_os_scan<<<NPASS, RADIX, 0, ctx->stream>>>(ghist, gbase);
    RFB_CHECK_LAUNCH(ctx);
    static thread_local unsigned long long hh[8 * RADIX];
    RFB_CUDA(cudaMemcpyAsync(hh, ghist, (size_t)NPASS * RADIX * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    int passes[8], np = 0;
    for (int q = 0; q < NPASS; q++) {
        bool constant = false;
        for (int d = 0; d < RADIX; d++) if (hh[q * RADIX + d] == (unsigned long long)n) { constant = true; break; }
        if (!constant) passes[np++] = q;
    }
    if (np == 0) {  // all keys equal: the identity permutation (stable)
        k_iota<<<rfb_grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(perm, n);
        RFB_CHECK_LAUNCH(ctx);
        return RFB_OK;
    }
    // an 8-byte column whose varying bytes fit a 32-bit window (ids, dates, small integers, doubles of a narrow range) travels as
    // 32-bit key words like the narrow types: 8 B per row moved and three CTAs per SM instead of 12 B and two
    if (sizeof(K) == 8 && passes[np - 1] - passes[0] < 4)
        return os_passes<T, u32>(ctx, (const T *)x, flip, n, passes, np, passes[0], gbase, status, counters, keysA, keysB, ridsA, ridsB, perm);
    return os_passes<T, K>(ctx, (const T *)x, flip, n, passes, np, 0, gbase, status, counters, keysA, keysB, ridsA, ridsB, perm);
}

template <typename T>
int sort_lsd(rfb_ctx_t *ctx, const void *x, i64 n, int descending, i64 *perm) {
    constexpr int NPASS = (int)sizeof(T);
    const u64 width_mask = NPASS == 8 ? ~0ULL : ((1ULL << (8 * NPASS)) - 1);
    ColumnSrc<T> col{(const T *)x, descending ? width_mask : 0ULL};
    // persistent chunks: G CTAs, chunk a multiple of the tile
    int G = ctx->sm_count * 8;
    i64 chunk = ((n + G - 1) / G + TILE - 1) / TILE * TILE;
    G = (int)((n + chunk - 1) / chunk);
    // workspace: census {or, and} | bh[256*G] u32 | offs[256*G] i64 | keysA[n] | keysB[n] | valsA[n]
    const size_t b_census = 256, b_bh = align256((size_t)RADIX * G * 4), b_offs = align256((size_t)RADIX * G * 8),
                 b_n = align256((size_t)n * 8);
    void *w;
    int rc = rfb_ensure_work(ctx, b_census + b_bh + b_offs + 3 * b_n, &w);
    if (rc) return rc;
    unsigned long long *census = (unsigned long long *)w;
    u32 *bh = (u32 *)((char *)w + b_census);
    i64 *offs = (i64 *)((char *)w + b_census + b_bh);
    u64 *keysA = (u64 *)((char *)w + b_census + b_bh + b_offs), *keysB = (u64 *)((char *)keysA + b_n);
    i64 *valsA = (i64 *)((char *)keysB + b_n);
    RFB_CUDA(cudaMemsetAsync(census, 0, 8, ctx->stream));
    RFB_CUDA(cudaMemsetAsync(census + 1, 0xFF, 8, ctx->stream));
    k_census<T><<<rfb_grid_for(ctx, n, THREADS * 4, 4), THREADS, 0, ctx->stream>>>(col, n, census);
    RFB_CHECK_LAUNCH(ctx);
    unsigned long long hc[2];
    RFB_CUDA(cudaMemcpyAsync(hc, census, 16, cudaMemcpyDeviceToHost, ctx->stream));
    RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    int passes[8], np = 0;
    for (int p = 0; p < NPASS; p++)
        if ((((hc[0] ^ hc[1]) >> (8 * p)) & 255ULL) != 0) passes[np++] = p;
    if (np == 0) {  // all keys equal: the identity permutation (stable)
        k_iota<<<rfb_grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(perm, n);
        RFB_CHECK_LAUNCH(ctx);
        return RFB_OK;
    }
    // ping-pong so that the LAST pass lands in `perm`: vals alternate between perm and valsA
    u64 *kin = nullptr, *kout = keysA;
    i64 *vin = nullptr;
    i64 *vout = (np % 2 == 1) ? perm : valsA;
    for (int q = 0; q < np; q++) {
        const bool last = (q == np - 1);
        const int shift = 8 * passes[q];
        if (q == 0) rc = run_pass(ctx, col, n, G, chunk, shift, bh, offs, kout, vout, last);
        else rc = run_pass(ctx, PairSrc{kin, vin}, n, G, chunk, shift, bh, offs, kout, vout, last);
        if (rc) return rc;
        kin = kout;
        kout = (kout == keysA) ? keysB : keysA;
        vin = vout;
        vout = (vout == perm) ? valsA : perm;
    }
    return RFB_OK;
}

// the single-sweep passes carry 32-bit row ids; longer columns (and RFB_SORT_ALGO=lsd) take the histogram + scatter passes
template <typename T>
int sort_t(rfb_ctx_t *ctx, const void *x, i64 n, int descending, i64 *perm) {
    if (n < (1ll << 32) && rfb_options()->sort_algo == 0) return sort_onesweep<T>(ctx, x, n, descending, perm);
    return sort_lsd<T>(ctx, x, n, descending, perm);
}

}  // namespace

extern "C" int rfb_sort_dev(rfb_ctx_t *ctx, int type, const void *x, int64_t n, int descending, int64_t *perm) {
    RFB_ARG(ctx && n >= 0 && ((x && perm) || n == 0), "rfb_sort_dev");
    if (type == RFB_SYMBOL || !rfb_kind_of(type)) { rfb_set_error("sort: unsupported type %d", type); return RFB_ERR_TYPE; }
    if (n == 0) return RFB_OK;
    switch (rfb_kind_of(type)) {
        case K_U8: return sort_t<u8>(ctx, x, n, descending, perm);
        case K_I16: return sort_t<i16>(ctx, x, n, descending, perm);
        case K_I32: return sort_t<i32>(ctx, x, n, descending, perm);
        case K_I64: return sort_t<i64>(ctx, x, n, descending, perm);
        default: return sort_t<f64>(ctx, x, n, descending, perm);
    }
}
