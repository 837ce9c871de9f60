// k_fused_group.cu — rfb_group_sum_count_dev (sm_100a): `select {s: (sum v) c: (count v) from t by k [where ...]}` fused, never
// materialising group ids.
#include "rfb_group.cuh"
#include "rfb_tma.cuh"

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

// k_hash_group.cu: the same query on a sparse key domain (shared-memory open-addressing tables + device-wide merge)
int rfb_hash_group_sum_count(rfb_ctx_t *ctx, int key_type, const void *keys, const int64_t *val, int64_t n, int cmp_op, int pred_type,
                             const void *pred, const rfb_scalar_t *k, int64_t max_groups, int64_t *out_keys, int64_t *out_sums,
                             int64_t *out_counts, int64_t *groups);

namespace {

constexpr int RFB_SPARSE_DOMAIN = 1;              // internal: the dense strategies decline, the caller takes the hash path
constexpr i64 DENSE_RANGE_MAX = 1ll << 28;        // largest key range addressed directly

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
    typedef P pred_t;
    static constexpr bool has_pred = HAS_PRED;
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
    bool tma_ok(const i64 *val) const { return aligned16(keys) && aligned16(val) && (!HAS_PRED || aligned16(pred)); }   // bulk copies: 16-byte aligned sources
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

constexpr int PCUR_STRIDE = 32;   // u32 words between two bucket cursors: one cache line each (atomics on one line serialise in its L2 slice)
struct PartStore {
    u32 *cursor;       // [MAX_PARTS * PCUR_STRIDE] rows per bucket at [b * PCUR_STRIDE]
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
    u32 cnt[2 * MAX_PARTS];      // two counters per bucket (even / odd lanes): halves the same-address conflicts of the ranking atomics
    u32 lb[2 * MAX_PARTS];       // tile-local base of each (bucket, copy)
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
    if (tid < MAX_PARTS) { sm.cnt[2 * tid] = 0; sm.cnt[2 * tid + 1] = 0; }
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
                if (lane == 0 && sel[j]) b = atomicAdd(&sm.cnt[2 * part], 32u);
                pos[j] = __shfl_sync(0xffffffffu, b, 0) + lane;
                pk[j] &= ~(1u << 31);                       // (copy 0)
            } else if (sel[j]) {
                pos[j] = atomicAdd(&sm.cnt[2 * part + (lane & 1)], 1u);
                pk[j] |= (u32)(lane & 1) << 31;             // remember the copy for the staging step
            }
        }
        __syncthreads();
        u32 c = 0, c0 = 0, incl = 0, start = 0, phys0 = 0, phys1 = 0;
        if (tid < MAX_PARTS) {
            c0 = sm.cnt[2 * tid];
            c = c0 + sm.cnt[2 * tid + 1];
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
                start = atomicAdd(&ps.cursor[tid * PCUR_STRIDE], c);
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
            sm.lb[2 * tid] = lbase;
            sm.lb[2 * tid + 1] = lbase + c0;
            sm.cnt[2 * tid] = 0;                               // for the next tile (this tile's counts live in registers now)
            sm.cnt[2 * tid + 1] = 0;
            if (tid == MAX_PARTS - 1) sm.total = before + incl;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SC_R; j++) {
            if (!sel[j]) continue;
            const u32 w = pk[j] & 0x7FFFFFFFu;
            const u32 q = sm.lb[2 * (w >> KP_LOG) + (pk[j] >> 31)] + pos[j];
            sm.val[q] = (u64)v[j];
            sm.pk[q] = w;
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
        c = tid < P ? ps.cursor[((bucket0 + tid) & 255u) * PCUR_STRIDE] : 0;
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
        c = tid < P ? ps.cursor[((bucket0 + tid) & 255u) * PCUR_STRIDE] : 0;
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


// ---- accumulate, strategy 2n ("narrow"): the same partition + accumulate idea for key ranges of at most 32 partitions, which
// is what the 1e5-key workload needs.  Three things differ from strategy 2:
//   * ranking by warp ballots instead of returning shared atomics: a row's rank among the rows of its warp step that go to the
//     same partition comes from five ballots over the partition bits; only the first lane of each such group touches the
//     partition's counter (distinct addresses -> one conflict-free atomic per warp step instead of 32 conflicting ones)
//   * records are PACKED: a 32-bit record holds the slot (KPL bits) and the value (32 - KPL bits) when the value fits, a
//     64-bit record the slot and a 48-bit signed value: 4 or 8 bytes per row instead of 10, written once and read once
//   * rows whose value does not fit the record (nulls included) are EXCEPTIONS: appended to a small side list of (key, value)
//     pairs that a tiny kernel folds with L2 atomics.  A row sample decides which record format is tried; when the guess was
//     wrong (the list overflows, or the keys span more than 32 partitions) the pass aborts early and the caller falls back.
constexpr int NP = 32;                          // partitions of the narrow path
#ifndef RFB_MS_T
#define RFB_MS_T 256
#define RFB_MS_R 8
#define RFB_MS_CTAS 3
#endif
constexpr int MS_T = RFB_MS_T, MS_CTAS = RFB_MS_CTAS, MS_WARPS = MS_T / 32;
// rows per thread by key width: 8-byte keys (i64 columns, group ids in rfb_narrow_sums) take 6 — their two 16 B/row input stages
// plus the staging area then fit three CTAs per SM like the 4-byte keys' (8 rows: 2 CTAs; measured 6.56 -> 6.14 ms per 1e9 rows,
// aggr_sum / aggr_avg over group ids 6.42 -> 6.01 ms; 4-byte keys with 6 rows: 5.37 -> 5.61 ms)
#ifndef RFB_MS_R8
#define RFB_MS_R8 6
#endif
template <typename KT> struct MsGeom { static constexpr int R = sizeof(KT) == 8 ? RFB_MS_R8 : RFB_MS_R, TILE = MS_T * R; };
static_assert(MS_T * RFB_MS_R <= PB && MS_T * RFB_MS_R8 <= PB, "a tile's run touches at most two blocks");

template <typename REC, int KPL> struct RecFmt;
template <int KPL> struct RecFmt<u32, KPL> {
    static constexpr int VB = 32 - KPL;
    __host__ __device__ static bool fits(i64 v) { return (u64)v < (1ull << VB); }
    __device__ __forceinline__ static u32 pack(u32 slot, i64 v) { return (slot << VB) | (u32)v; }
    __device__ __forceinline__ static u32 slot(u32 r) { return r >> VB; }
    __device__ __forceinline__ static i64 val(u32 r) { return (i64)(r & ((1u << VB) - 1u)); }
};
template <int KPL> struct RecFmt<u64, KPL> {
    static constexpr int VB = 48;
    __host__ __device__ static bool fits(i64 v) { return (u64)v + (1ull << 47) < (1ull << 48); }
    __device__ __forceinline__ static u64 pack(u32 slot, i64 v) { return ((u64)slot << 48) | ((u64)v & 0xFFFFFFFFFFFFull); }
    __device__ __forceinline__ static u32 slot(u64 r) { return (u32)(r >> 48); }
    __device__ __forceinline__ static i64 val(u64 r) { return (i64)(r << 16) >> 16; }
};

// Every tile adds to the cursor of every partition it holds rows for: ~25 atomics per 2048-row tile.  Atomics on one cache line
// are serialised by the L2 slice that owns it (measured on B200: ~1.2 atomics/ns per line — with all cursors on one line the
// whole scatter pass ran at the speed of that one slice), so every cursor gets its own 256 bytes.
constexpr int CUR_STRIDE = 64;                  // u32 words between the cursors of two partitions
struct RecStore {
    u32 *cursor;       // [NP * CUR_STRIDE] rows per partition at [p * CUR_STRIDE]
    u32 *next_block;   // blocks handed out so far
    u32 *exc_count;    // exceptions appended so far (may run past exc_cap: then the pass is void)
    u32 *bt;           // [NP][bt_stride] physical block + 1 (0 = not yet allocated)
    u32 bt_stride;
    u32 exc_cap;
    void *rec;         // [blocks * PB] records
    i64 *exc;          // [exc_cap][2] (key, value)
};

// ballot of (x & mask) != 0 over the warp.  Written in PTX so that the test stays one LOP3 with a predicate result (the
// compiler's canonical form of the C expression is shift + and + compare: three ALU instructions per ballot instead of one)
__device__ __forceinline__ u32 ballot_bits(u32 x, u32 mask) {
    u32 r;
    asm volatile("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %1, %2;\nsetp.ne.u32 p, t, 0;\nvote.sync.ballot.b32 %0, p, 0xffffffff;\n}" : "=r"(r) : "r"(x), "r"(mask));
    return r;
}

// One tile of the scatter pass: ranking (ballots) -> barrier #1 (the tile's partition counts are final) -> every warp derives the
// tile-local layout itself (a 32-lane scan, lane = partition) and stages its rows; warp 0 has posted the global reservations
// of all partitions right after the barrier, so their latency hides behind the staging, and turns them into per-partition
// output descriptors -> barrier #2 -> all threads write the staged tile out linearly (a byte per staged row names its
// partition).  Staging area, counters and descriptors are double-buffered: no third barrier before the next tile starts.
// INPUT is staged by the TMA unit: thread 0 issues the bulk copies of the NEXT tile's key / value / predicate slices into a
// two-stage shared-memory ring at the top of every tile (its stage was last read before the previous tile's barriers), so
// the DRAM latency of a whole tile is hidden without holding a single register for it.
// The kernel takes FULL tiles of 16-byte aligned columns only; the rows past the last full tile go through k_ms_tail,
// unaligned columns do not take the narrow path at all.
template <typename FS> struct MsIn {   // one input stage
    typedef typename FS::key_t KT;
    typedef typename FS::pred_t PT_;
    static constexpr int MS_TILE = MsGeom<KT>::TILE;
    static constexpr size_t KEY_BYTES = MS_TILE * sizeof(KT), VAL_BYTES = MS_TILE * 8, PRED_BYTES = FS::has_pred ? MS_TILE * sizeof(PT_) : 0;
    static constexpr size_t BYTES = KEY_BYTES + VAL_BYTES + PRED_BYTES;
    // a predicate on the aggregated column itself (`where (< v k)` next to `(sum v)`) reads the staged values: no third slice
    __host__ __device__ static bool pred_is_val(const FS &fs, const i64 *val) { return FS::has_pred && sizeof(PT_) == 8 && (const void *)fs.pred == (const void *)val; }
    __host__ __device__ static size_t bytes(const FS &fs, const i64 *val) { return pred_is_val(fs, val) ? KEY_BYTES + VAL_BYTES : BYTES; }
};

// Warp roles: MS_WARPS data warps rank, stage and write out rows; one extra CONTROL warp (lane = partition) owns everything
// that waits on the memory system — the TMA issue, the global reservations of the tile's runs (one atomic per partition), the
// block-table lookups and the output descriptors.  The write-out of tile i is delayed until after tile i+1's ranking, so the
// control warp's two dependent L2 round trips overlap a whole tile of data-warp work instead of stalling a barrier.
// The data warps and the control warp meet at two barriers per tile.  Both roles reach the SAME two __syncthreads() instructions
// (the role-specific work sits between them): compute-sanitizer synccheck rejects barriers taken from different code paths.
template <typename FS, typename REC, int KPL>
__global__ void __launch_bounds__(MS_T + 32, MS_CTAS)
k_ms_scatter(FS fs, const i64 *__restrict__ val, i64 tiles, RecStore rs, i64 *mm) {
    typedef RecFmt<REC, KPL> F;
    typedef MsIn<FS> In;
    typedef typename FS::key_t KT;
    typedef typename FS::pred_t PT_;
    constexpr u32 KPN = 1u << KPL;
    constexpr int MS_R = MsGeom<KT>::R, MS_TILE = MsGeom<KT>::TILE;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    const bool pred_alias = In::pred_is_val(fs, val);
    const size_t in_bytes = In::bytes(fs, val);
    unsigned char *const s_in = s_dyn;                                                        // [2] input stages: keys | values | predicate column
    REC (*const s_rec)[MS_TILE] = (REC (*)[MS_TILE])(s_dyn + 2 * in_bytes);                   // [2][MS_TILE] staged records
    u8 (*const s_part)[MS_TILE] = (u8 (*)[MS_TILE])(s_dyn + 2 * in_bytes + 2 * MS_TILE * sizeof(REC));   // [2][MS_TILE] partition of every staged record
    __shared__ u64 full[2];
    __shared__ u32 s_cnt[2][NP];
    __shared__ u32 s_wb[MS_WARPS][NP];   // per data warp: where its rows of each partition start in the tile's staging area
    __shared__ uint4 s_desc[2][NP];      // per partition: {x: global - local index before the block boundary, y: first local index past it, z: global - local after it}
    __shared__ u32 s_total[2];
    __shared__ u32 s_abort;
    __shared__ i64 red[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool ctrl = warp == MS_WARPS;
    const u32 lt_mask = (1u << lane) - 1u;
    u32 cb[5];                           // lane L collects the lanes of partition L: cb[b] flips ballot b where bit b of L is clear
#pragma unroll
    for (int b = 0; b < 5; b++) cb[b] = ((lane >> b) & 1) ? 0u : 0xFFFFFFFFu;
    KT lo = sizeof(KT) == 4 ? (KT)0x7FFFFFFF : (KT)RFB_INF_I64, hi = sizeof(KT) == 4 ? (KT)NULL_I32 : (KT)NULL_I64;
    REC *const grec = (REC *)rs.rec;
    if (tid < 2 * NP) s_cnt[tid >> 5][tid & 31] = 0;
    if (tid == 0) {
        s_abort = 0;
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](i64 tile, int st) {          // one lane: the three slices of one tile
        unsigned char *dst = s_in + (size_t)st * in_bytes;
        mbar_expect_tx(&full[st], (u32)in_bytes);
        bulk_g2s(dst, fs.keys + tile * MS_TILE, (u32)In::KEY_BYTES, &full[st]);
        bulk_g2s(dst + In::KEY_BYTES, val + tile * MS_TILE, (u32)In::VAL_BYTES, &full[st]);
        if constexpr (FS::has_pred)
            if (!pred_alias) bulk_g2s(dst + In::KEY_BYTES + In::VAL_BYTES, fs.pred + tile * MS_TILE, (u32)In::PRED_BYTES, &full[st]);
    };
    auto write_out = [&](int b) {                 // data warps: the staged tile in buffer b, linearly
        const u32 total = s_total[b];
#pragma unroll
        for (int j = 0; j < MS_R; j++) {
            const u32 q = j * MS_T + tid;
            if (q < total) {
                const uint4 d = s_desc[b][s_part[b][q]];
                grec[q + (q < d.y ? d.x : d.z)] = s_rec[b][q];
            }
        }
    };
    if (ctrl && lane == 0 && (i64)blockIdx.x < tiles) issue(blockIdx.x, 0);
    int buf = 0;
    u32 it = 0;
    // data warps, first half of a tile: wait for the staged input, rank the thread's MS_R rows among the rows of their partition
    // inside the warp, reserve the warp's runs in the tile's staging area (wbase), send misfits to the side list
    auto rank_tile = [&](REC (&rec)[MS_R], u32 (&pp)[MS_R], u32 &wbase) {
        mbar_wait(&full[buf], (it >> 1) & 1u);
        const unsigned char *in = s_in + (size_t)buf * in_bytes;
        const KT *const sk = (const KT *)in;
        const i64 *const sv = (const i64 *)(in + In::KEY_BYTES);
        // rows of this thread: pairs (j2 * MS_T + tid) of the tile: conflict-free 8 / 16-byte shared-memory reads
        KT k[MS_R];
        i64 v[MS_R];
        bool sel[MS_R];
#pragma unroll
        for (int j2 = 0; j2 < MS_R / 2; j2++) {
            const int pair = j2 * MS_T + tid;
            if constexpr (sizeof(KT) == 4) {
                const uint2 w = *(const uint2 *)(sk + 2 * pair);
                k[2 * j2] = (KT)w.x; k[2 * j2 + 1] = (KT)w.y;
            } else {
                const ulonglong2 w = *(const ulonglong2 *)(sk + 2 * pair);
                k[2 * j2] = (KT)w.x; k[2 * j2 + 1] = (KT)w.y;
            }
            const ulonglong2 w = *(const ulonglong2 *)(sv + 2 * pair);
            v[2 * j2] = (i64)w.x; v[2 * j2 + 1] = (i64)w.y;
            if constexpr (FS::has_pred) {
                const PT_ *const sp = (const PT_ *)(in + In::KEY_BYTES + (pred_alias ? 0 : In::VAL_BYTES));
                sel[2 * j2] = pred_test(pred_key<PT_>(sp[2 * pair]), fs.pr);
                sel[2 * j2 + 1] = pred_test(pred_key<PT_>(sp[2 * pair + 1]), fs.pr);
            } else { sel[2 * j2] = sel[2 * j2 + 1] = true; }
        }
        u32 exc_mask = 0;                // rows of this thread that go to the side list
        u32 wcount = 0;                  // rows of partition `lane` this warp has ranked so far in this tile
#pragma unroll
        for (int j = 0; j < MS_R; j++) {
            const KT kj = k[j];
            lo = (sel[j] && kj < lo) ? kj : lo;
            hi = (sel[j] && kj > hi) ? kj : hi;
            const bool ok = sel[j] && F::fits(v[j]);
            exc_mask |= (u32)(sel[j] && !ok) << j;
            const u32 kb = (u32)kj;
            const u32 part = (kb >> KPL) & (NP - 1);
            rec[j] = F::pack(kb & (KPN - 1u), v[j]);
            // Ranking by ballots: lane L ends up with the set of lanes whose row goes to partition L (five ballots over
            // the partition bits, each flipped where L's bit is clear); a row then fetches its own partition's set and
            // running count from lane `part`.  No shared-memory traffic and no atomics per row.
            u32 pl = __ballot_sync(0xffffffffu, ok);
#pragma unroll
            for (int b = 0; b < 5; b++) pl &= ballot_bits(kb, 1u << (KPL + b)) ^ cb[b];
            const u32 peers = __shfl_sync(0xffffffffu, pl, part);
            const u32 before = __shfl_sync(0xffffffffu, wcount, part);
            wcount += __popc(pl);
            pp[j] = ((before + __popc(peers & lt_mask)) << 8) | (ok ? part : 32u);
        }
        if (wcount) wbase = atomicAdd(&s_cnt[buf][lane], wcount);   // one atomic per (warp, partition) and tile, distinct addresses
        if (exc_mask) {                                              // rare
#pragma unroll
            for (int j = 0; j < MS_R; j++)
                if (exc_mask & (1u << j)) {
                    const u32 e = atomicAdd(rs.exc_count, 1u);
                    if (e < rs.exc_cap) { rs.exc[2 * (size_t)e] = (i64)k[j]; rs.exc[2 * (size_t)e + 1] = v[j]; }
                }
        }
    };
    for (i64 tile = blockIdx.x; tile < tiles; tile += gridDim.x, buf ^= 1, it++) {
        // both roles walk the same two barriers per tile (one __syncthreads() instruction each, reached by the whole block)
        REC rec[MS_R];
        u32 pp[MS_R];                        // rank among the warp's rows of the partition << 8 | partition (32 = not staged)
        u32 wbase = 0;
        if (ctrl) {
            // the other input stage was last read before barrier A of the previous tile: refill it now
            if (lane == 0 && tile + gridDim.x < tiles) issue(tile + gridDim.x, buf ^ 1);
        } else {
            rank_tile(rec, pp, wbase);
        }
        __syncthreads();                                             // A: the tile's partition counts are final
        if (ctrl) {
            const u32 c = s_cnt[buf][lane];
            u32 incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            const u32 lbase = incl - c;
            // reserve [start, start + c) of the partition's stream; blocks of PB records are handed out on demand: the run that
            // contains a block's first record allocates it and publishes it in the block table, everybody else waits for that
            // word (same protocol as k_part_scatter)
            u32 start = 0, phys0 = 1, phys1 = 1;
            if (c) {
                start = atomicAdd(&rs.cursor[lane * CUR_STRIDE], c);
                const u32 b0 = start >> PB_LOG, b1 = (start + c - 1) >> PB_LOG;
                u32 *row = rs.bt + (size_t)lane * rs.bt_stride;
                phys0 = phys1 = 0;
                if ((start & (PB - 1)) == 0) { phys0 = atomicAdd(rs.next_block, 1u) + 1; st_relaxed_u32(row + b0, phys0); }
                if (b1 != b0) { phys1 = atomicAdd(rs.next_block, 1u) + 1; st_relaxed_u32(row + b1, phys1); }
                // (measured: caching the partition's last block in the lane to skip this second dependent L2 round trip makes the
                // whole pass 3 % SLOWER, 5.27 -> 5.43 ms — the control warp then reaches barrier B early and the data warps' write-out
                // loses the overlap with it)
                while (!phys0) phys0 = ld_relaxed_u32(row + b0);
                if (b1 == b0) phys1 = phys0;
            }
            const u32 in_block = start & (PB - 1), room = PB - in_block;
            uint4 d;
            d.x = (phys0 - 1) * (u32)PB + in_block - lbase;
            d.y = lbase + room;
            d.z = (phys1 - 1) * (u32)PB - (lbase + room);
            d.w = 0;
            s_desc[buf][lane] = d;
            s_cnt[buf ^ 1][lane] = 0;                                // the next tile's counters: last read before the previous barrier B
            if (lane == 31) s_total[buf] = incl;
            if (lane == 0 && ld_relaxed_u32(rs.exc_count) > rs.exc_cap) s_abort = 1;
        } else {
            const u32 c = s_cnt[buf][lane];
            u32 incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            s_wb[warp][lane] = incl - c + wbase;                     // start of partition `lane` in the staging area + this warp's offset in it
            __syncwarp();
#pragma unroll
            for (int j = 0; j < MS_R; j++)
                if (!(pp[j] & 32u)) {
                    const u32 q = s_wb[warp][pp[j] & 31u] + (pp[j] >> 8);
                    s_rec[buf][q] = rec[j];
                    s_part[buf][q] = (u8)(pp[j] & 31u);
                }
            if (it > 0) write_out(buf ^ 1);                          // the PREVIOUS tile: its descriptors were complete at its barrier B
        }
        __syncthreads();                                             // B: this tile is staged and described
        if (s_abort) break;                                          // written before B, uniform: the record format was a bad guess
    }
    if (!ctrl && it > 0 && !s_abort) write_out(buf ^ 1);             // the last tile
    struct Mn { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b < a ? b : a; } };
    struct Mx { __device__ __forceinline__ i64 operator()(i64 a, i64 b) const { return b > a ? b : a; } };
    const i64 lo64 = block_reduce<i64>((i64)lo, Mn(), RFB_INF_I64, red);
    const i64 hi64 = block_reduce<i64>((i64)hi, Mx(), NULL_I64, red);
    if (tid == 0) {
        atomicMin((long long *)&mm[0], (long long)lo64);
        atomicMax((long long *)&mm[1], (long long)hi64);
    }
}

// the rows past the last full tile: every selected one goes to the side list (fewer than MS_TILE rows, the list holds >= 4096)
template <typename FS>
__global__ void __launch_bounds__(THREADS) k_ms_tail(FS fs, const i64 *__restrict__ val, i64 r0, i64 n, RecStore rs, i64 *mm) {
    for (i64 i = r0 + threadIdx.x; i < n; i += THREADS)
        if (fs.selected(i)) {
            const i64 k = fs.key(i);
            atomicMin((long long *)&mm[0], (long long)k);
            atomicMax((long long *)&mm[1], (long long)k);
            const u32 e = atomicAdd(rs.exc_count, 1u);
            if (e < rs.exc_cap) { rs.exc[2 * (size_t)e] = k; rs.exc[2 * (size_t)e + 1] = val[i]; }
        }
}

// non-null values that fit the record: the low word takes the value, a wrap carries into the high word
__device__ __forceinline__ void sacc_add_small(const SAcc &a, u32 s, i64 v) {
    const u32 lo = (u32)(u64)v;
    u32 hi = (u32)((u64)v >> 32);
    const u32 old = atomicAdd(&a.lo[s], lo);
    hi += (u32)((u32)(old + lo) < lo);
    if (hi) atomicAdd(&a.hi[s], hi);
    atomicAdd(&a.cnt[s], 1u);
}

// 32-bit records (values below 2^VB <= 2^20): ONE shared atomic per record.  The slot's first word packs the record count (top 8
// bits) over the value sum (low 24 bits): a record adds 2^24 + v.  With T the exact total of what was added, the word is
// T mod 2^32.  The adder whose value overflowed the 24-bit sum field sees it in the returned old word (the low 24 bits evolve as
// sum mod 2^24 whatever the upper bits do) and counts a CARRY in the second word; the adder that wrapped the whole word counts a
// WRAP in the third.  Both events are exact functions of (old, increment), so at flush time
//     T = wraps * 2^32 + word,   sum = carries * 2^24 + (word mod 2^24),   count = (T >> 24) - carries
// hold exactly.  A carry needs ~2^24 / v records (>= 16), a wrap <= 256: 1.07 atomics per record instead of 2.
__device__ __forceinline__ void sacc_add_packed(const SAcc &a, u32 s, u32 v) {
    const u32 inc = (1u << 24) + v;
    const u32 old = atomicAdd(&a.lo[s], inc);
    if ((old & 0xFFFFFFu) + v >= (1u << 24)) atomicAdd(&a.hi[s], 1u);
    if ((u32)(old + inc) < inc) atomicAdd(&a.cnt[s], 1u);
}
__device__ __forceinline__ void sacc_flush_packed(const SAcc &a, int slots, i64 slot0, const Accums &ga) {
    for (int s = threadIdx.x; s < slots; s += blockDim.x) {
        const u32 w = a.lo[s], carries = a.hi[s], wraps = a.cnt[s];
        if (!(w | carries | wraps)) continue;
        const i64 g = slot0 + s;
        const u64 total = ((u64)wraps << 32) | w;
        const u64 sum = ((u64)carries << 24) + (w & 0xFFFFFFu);
        if (sum) atomicAdd((unsigned long long *)ga.sum + g, (unsigned long long)sum);
        atomicAdd((unsigned long long *)ga.cnt + g, (unsigned long long)((total >> 24) - carries));
        a.lo[s] = 0; a.hi[s] = 0; a.cnt[s] = 0;
    }
}
template <typename REC> __device__ __forceinline__ void msa_add(const SAcc &a, u32 s, i64 v) {
    if constexpr (sizeof(REC) == 4) sacc_add_packed(a, s, (u32)v); else sacc_add_small(a, s, v);
}
template <typename REC> __device__ __forceinline__ void msa_flush(const SAcc &a, int slots, i64 slot0, const Accums &ga) {
    if constexpr (sizeof(REC) == 4) sacc_flush_packed(a, slots, slot0, ga); else sacc_flush(a, slots, slot0, ga);
}

// accumulate pass of the narrow path: one CTA per SM, the partition's records staged by the TMA unit (cp.async.bulk into an
// mbarrier ring) next to the partition's 2^KPL accumulators.  Partition p = absolute bucket ((kbase >> KPL) + p) mod 32.
constexpr int MS_UNIT = 4096;                   // records per work unit
// Geometry of the accumulate pass.  The pass is bound by shared-memory atomics (two per record, ~3.5-way bank conflicts on random
// slots) and by the latency of the returning one: two 512-thread CTAs per SM (each with its own copy of the partition's
// accumulators and a 3-stage ring) hide it better than one 1024-thread CTA — measured 1.49 -> 1.19 ms per 1e9 records.
#ifndef RFB_MSA_T
#define RFB_MSA_T 512
#define RFB_MSA_CTAS 2
#define RFB_MSA_STAGES 3
#endif
// 8192-slot partitions (96 KB of accumulators) leave room for one CTA per SM only: 1024 threads, 4 / 3 stages.
template <typename REC, int KPL> struct MsaCfg {
    static constexpr bool TWO = KPL <= 12 && RFB_MSA_CTAS > 1;
    static constexpr int T = TWO ? RFB_MSA_T : 1024, CTAS = TWO ? RFB_MSA_CTAS : 1;
    static constexpr int STAGES = TWO ? RFB_MSA_STAGES : (sizeof(REC) == 4 ? 4 : 3);
};

template <typename REC, int KPL>
__global__ void __launch_bounds__((MsaCfg<REC, KPL>::T), (MsaCfg<REC, KPL>::CTAS))
k_ms_accum_tma(RecStore rs, int P, i64 kbase, i64 kmin, Accums ga) {
    typedef RecFmt<REC, KPL> F;
    constexpr int KPN = 1 << KPL, STAGES = MsaCfg<REC, KPL>::STAGES, MSA_T = MsaCfg<REC, KPL>::T, PER = 4;   // records per thread and step
    static_assert(MS_UNIT % (MSA_T * PER) == 0, "whole steps of four records per thread");
    extern __shared__ __align__(16) unsigned char s_dyn[];
    REC *stage = (REC *)s_dyn;                                                  // STAGES x MS_UNIT records
    u32 *s_acc = (u32 *)(s_dyn + (size_t)STAGES * MS_UNIT * sizeof(REC));       // lo[KPN] | hi[KPN] | cnt[KPN]
    __shared__ u64 full[STAGES];
    __shared__ u32 s_cnt[NP], s_ubase[NP + 1];
    const SAcc a{s_acc, s_acc + KPN, s_acc + 2 * KPN};
    sacc_zero(a, KPN);
    const int tid = threadIdx.x, lane = tid & 31;
    const u32 bucket0 = (u32)(((u64)kbase >> KPL) & (NP - 1));
    const REC *const grec = (const REC *)rs.rec;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        const u32 c = tid < P ? rs.cursor[((bucket0 + tid) & (NP - 1)) * CUR_STRIDE] : 0;
        s_cnt[tid] = c;
        u32 incl = (c + MS_UNIT - 1) / MS_UNIT;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        s_ubase[tid + 1] = incl;
        if (tid == 0) s_ubase[0] = 0;
    }
    __syncthreads();
    const u32 U = s_ubase[P];
    const u32 u0 = (u32)((u64)blockIdx.x * U / gridDim.x), u1 = (u32)((u64)(blockIdx.x + 1) * U / gridDim.x);
    int pp = 0;                                   // producer state (thread 0 only)
    u32 p_blk = 0xFFFFFFFFu, p_phys = 0;
    auto issue = [&](u32 u) {
        while (s_ubase[pp + 1] <= u) { pp++; p_blk = 0xFFFFFFFFu; }
        const u32 r0 = (u - s_ubase[pp]) * MS_UNIT, cnt = s_cnt[pp];
        const u32 rows = cnt - r0 < (u32)MS_UNIT ? cnt - r0 : (u32)MS_UNIT;
        const u32 bytes = (rows * (u32)sizeof(REC) + 15u) & ~15u;
        const u32 bucket = (bucket0 + pp) & (NP - 1);
        if ((r0 >> PB_LOG) != p_blk) { p_blk = r0 >> PB_LOG; p_phys = rs.bt[(size_t)bucket * rs.bt_stride + p_blk] - 1; }
        const u64 base = (u64)p_phys * PB + (r0 & (PB - 1));
        const int s = (int)((u - u0) % STAGES);
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(stage + (size_t)s * MS_UNIT, grec + base, bytes, &full[s]);
    };
    if (tid == 0)
        for (u32 u = u0; u < u1 && u < u0 + (STAGES - 1); u++) issue(u);
    int p = 0;
    while (p + 1 < P && s_ubase[p + 1] <= u0) p++;
    bool dirty = false;
    for (u32 u = u0; u < u1; u++) {
        const u32 k = u - u0;
        const int s = (int)(k % STAGES);
        __syncthreads();                                                       // unit u-1 fully consumed: its stage may be re-armed
        if (tid == 0 && u + (STAGES - 1) < u1) issue(u + (STAGES - 1));
        if (s_ubase[p + 1] <= u) {
            if (dirty) msa_flush<REC>(a, KPN, (i64)((u64)kbase + (u64)p * KPN - (u64)kmin), ga);
            __syncthreads();
            dirty = false;
            while (s_ubase[p + 1] <= u) p++;
        }
        const u32 r0 = (u - s_ubase[p]) * MS_UNIT, cnt = s_cnt[p];
        const u32 rows = cnt - r0 < (u32)MS_UNIT ? cnt - r0 : (u32)MS_UNIT;
        mbar_wait(&full[s], (k / STAGES) & 1u);
        dirty = true;
        const REC *st = stage + (size_t)s * MS_UNIT;
#pragma unroll
        for (u32 q0 = (u32)tid * PER; q0 < (u32)MS_UNIT; q0 += MSA_T * PER)
        if (q0 + PER <= rows) {
            REC r[PER];
            if constexpr (sizeof(REC) == 4) {
                const uint4 w = *(const uint4 *)(st + q0);
                r[0] = w.x; r[1] = w.y; r[2] = w.z; r[3] = w.w;
            } else {
                const ulonglong2 w0 = *(const ulonglong2 *)(st + q0), w1 = *(const ulonglong2 *)(st + q0 + 2);
                r[0] = w0.x; r[1] = w0.y; r[2] = w1.x; r[3] = w1.y;
            }
#pragma unroll
            for (int j = 0; j < PER; j++) msa_add<REC>(a, F::slot(r[j]), F::val(r[j]));
        } else {
            for (u32 q = q0; q < rows && q < q0 + PER; q++) msa_add<REC>(a, F::slot(st[q]), F::val(st[q]));
        }
    }
    __syncthreads();
    if (dirty) msa_flush<REC>(a, KPN, (i64)((u64)kbase + (u64)p * KPN - (u64)kmin), ga);
}


// the exception rows of the narrow path: device-wide accumulators, L2 atomics
__global__ void __launch_bounds__(THREADS) k_ms_exceptions(const i64 *__restrict__ exc, const u32 *__restrict__ exc_count, i64 kmin, Accums a, bool nulls_count) {
    const u32 m = *exc_count;
    for (u32 i = blockIdx.x * THREADS + threadIdx.x; i < m; i += gridDim.x * THREADS) {
        const i64 k = exc[2 * (size_t)i], v = exc[2 * (size_t)i + 1];
        const i64 s = (i64)((u64)k - (u64)kmin);
        if (v == NULL_I64) a.has_null[s] = 1u; else atomicAdd((unsigned long long *)a.sum + s, (unsigned long long)v);
        if (v != NULL_I64 || nulls_count) atomicAdd((unsigned long long *)a.cnt + s, 1ULL);
    }
}

// value census of the rows [r0, r1): how many selected rows do not fit a 19-bit / 20-bit unsigned / 48-bit signed record
template <typename FS>
__global__ void __launch_bounds__(THREADS, BLOCKS_PER_SM) k_val_census_rows(FS fs, const i64 *__restrict__ val, i64 r0, i64 r1, i64 *out) {
    __shared__ i64 red[32];
    i64 c19 = 0, c20 = 0, c48 = 0, rows = 0;
    for (i64 i = r0 + (i64)blockIdx.x * THREADS + threadIdx.x; i < r1; i += (i64)gridDim.x * THREADS)
        if (fs.selected(i)) {
            const i64 v = ld_stream(val + i);
            rows++;
            c19 += !RecFmt<u32, 13>::fits(v);
            c20 += !RecFmt<u32, 12>::fits(v);
            c48 += !RecFmt<u64, 13>::fits(v);
        }
    c19 = block_reduce<i64>(c19, OpAddWrap(), 0, red);
    c20 = block_reduce<i64>(c20, OpAddWrap(), 0, red);
    c48 = block_reduce<i64>(c48, OpAddWrap(), 0, red);
    rows = block_reduce<i64>(rows, OpAddWrap(), 0, red);
    if (threadIdx.x == 0) {
        atomicAdd((unsigned long long *)&out[0], (unsigned long long)c19);
        atomicAdd((unsigned long long *)&out[1], (unsigned long long)c20);
        atomicAdd((unsigned long long *)&out[2], (unsigned long long)c48);
        atomicAdd((unsigned long long *)&out[3], (unsigned long long)rows);
    }
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

// tuning knobs (environment, read once — rfb_options_reload() re-reads them): RFB_GROUP_STRATEGY = smem | part | narrow | l2 | hash
// forces a strategy where it is applicable; RFB_PART_MIN_ROWS = smallest row count that takes the partitioned strategy
int group_strategy_forced() { return rfb_options()->group_strategy; }
i64 part_min_rows() { return rfb_options()->part_min_rows; }


// ---- host side of the narrow path
struct NarrowPlan { int rec_bytes; int kpl; };   // rec_bytes 0 = not applicable

static inline i64 buckets_spanned(i64 kmin, i64 kmax, int kpl) { return (kmax >> kpl) - (kmin >> kpl) + 1; }   // arithmetic shifts: floor

template <typename FS, typename REC, int KPL>
int ms_scatter_launch(rfb_ctx_t *ctx, FS fs, const i64 *val, i64 n, const RecStore &rs, i64 *mm) {
    constexpr int MS_TILE = MsGeom<typename FS::key_t>::TILE;
    const size_t smem = 2 * MsIn<FS>::bytes(fs, val) + 2 * (size_t)MS_TILE * (sizeof(REC) + 1);
    const i64 tiles = n / MS_TILE;                 // full tiles; the rest goes through the side list
    if (tiles > 0) {
        RFB_CUDA(cudaFuncSetAttribute(k_ms_scatter<FS, REC, KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = (int)((size_t)(220 << 10) / (smem + 2048));   // CTAs per SM the shared memory allows
        if (per_sm > MS_CTAS) per_sm = MS_CTAS;
        if (per_sm < 1) per_sm = 1;
        k_ms_scatter<FS, REC, KPL><<<rfb_grid_for(ctx, tiles * MS_TILE, MS_TILE, per_sm), MS_T + 32, smem, ctx->stream>>>(fs, val, tiles, rs, mm);
        RFB_CHECK_LAUNCH(ctx);
    }
    if (tiles * MS_TILE < n) {
        k_ms_tail<FS><<<1, THREADS, 0, ctx->stream>>>(fs, val, tiles * MS_TILE, n, rs, mm);
        RFB_CHECK_LAUNCH(ctx);
    }
    return RFB_OK;
}
template <typename REC, int KPL>
int ms_accum_launch(rfb_ctx_t *ctx, const RecStore &rs, int P, i64 kbase, i64 kmin, const Accums &a) {
    typedef MsaCfg<REC, KPL> C;
    const size_t smem = (size_t)C::STAGES * MS_UNIT * sizeof(REC) + (size_t)(1 << KPL) * 12;
    RFB_CUDA(cudaFuncSetAttribute(k_ms_accum_tma<REC, KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ms_accum_tma<REC, KPL><<<ctx->sm_count * C::CTAS, C::T, smem, ctx->stream>>>(rs, P, kbase, kmin, a);
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

template <typename FS>
int fused_run(rfb_ctx_t *ctx, FS fs, const i64 *val, i64 n, i64 max_groups, i64 *out_keys, i64 *out_sums, i64 *out_counts, i64 *groups) {
    i64 *mm = (i64 *)((char *)ctx->d_scratch + 32768);
    const int forced = group_strategy_forced();
    if (forced == 5) return RFB_SPARSE_DOMAIN;
    const bool part_able = n < 0xF0000000ll && (forced == 2 || forced == 4 || (forced == 0 && n >= part_min_rows()));   // 32-bit row positions
    const int grid = rfb_grid_for(ctx, n, THREADS * 4, BLOCKS_PER_SM);
    const int sgrid = rfb_grid_for(ctx, n, STILE, 4);
    const bool vec = fs.vec_ok(val);
    const i64 tiles_max = (n + scan::RowTile<NUM_J>::TILE - 1) / scan::RowTile<NUM_J>::TILE;
    auto parts_of = [](i64 kmin, i64 kmax) { return (i64)(((u64)kmax - (u64)(kmin & ~(i64)(KP - 1))) >> KP_LOG) + 1; };
    i64 h[2];
    int rc;
    bool have_scope = false, scattered = false, modded = false, narrowed = false;
    void *w = nullptr;
    PartStore ps{};
    RecStore rs{};
    NarrowPlan plan{0, 0};
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
        if (h[0] <= h[1] && (u64)h[1] - (u64)h[0] >= (u64)DENSE_RANGE_MAX) return RFB_SPARSE_DOMAIN;   // no point in a full scope pass
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
        // narrow path: at most 32 partitions, packed records chosen from a value census of the same row windows
        if (!modded && vec && fs.tma_ok(val) && h[0] <= h[1] && (forced == 0 || forced == 4) && (forced == 4 || (i64)((u64)h[1] - (u64)h[0]) >= KP) &&
            (u64)h[1] - (u64)h[0] < (u64)NP * KP) {
            i64 *census = mm + 8;
            RFB_CUDA(cudaMemsetAsync(census, 0, 32, ctx->stream));
            for (int s = 0; s < 3; s++) {
                const i64 r0 = starts[s], r1 = r0 + win < n ? r0 + win : n;
                k_val_census_rows<FS><<<64, THREADS, 0, ctx->stream>>>(fs, val, r0, r1, census);
                RFB_CHECK_LAUNCH(ctx);
            }
            i64 c[4];
            rc = d2h_sync(ctx, c, census, 32);
            if (rc) return rc;
            const i64 tol = c[3] / 64;
            if (buckets_spanned(h[0], h[1], 12) <= NP && c[1] <= tol) plan = NarrowPlan{4, 12};
            else if (buckets_spanned(h[0], h[1], 13) <= NP && c[0] <= tol) plan = NarrowPlan{4, 13};
            else if (buckets_spanned(h[0], h[1], 13) <= NP && c[2] <= tol) plan = NarrowPlan{8, 13};
        }
        const i64 hs[2] = {h[0], h[1]};
        if (plan.rec_bytes) {
            const u32 blocks = (u32)((n + PB - 1) / PB) + NP + 1;
            const u32 bt_stride = (u32)((n + PB - 1) / PB) + 1;
            i64 cap = n / 64 > 4096 ? n / 64 : 4096;
            if (cap > (1ll << 23)) cap = 1ll << 23;
            const size_t ctl_bytes = align256((size_t)(NP + 2) * CUR_STRIDE * 4), bt_bytes = align256((size_t)NP * bt_stride * 4);
            const size_t rec_bytes = align256((size_t)blocks * PB * plan.rec_bytes), exc_bytes = align256((size_t)cap * 16);
            rc = rfb_ensure_work(ctx, p_acc_bytes + ctl_bytes + bt_bytes + rec_bytes + exc_bytes, &w);
            if (rc) return rc;
            char *pw = (char *)w + p_acc_bytes;
            rs.cursor = (u32 *)pw;
            rs.next_block = rs.cursor + NP * CUR_STRIDE;
            rs.exc_count = rs.cursor + (NP + 1) * CUR_STRIDE;
            rs.bt = (u32 *)(pw + ctl_bytes);
            rs.bt_stride = bt_stride;
            rs.exc_cap = (u32)cap;
            rs.rec = pw + ctl_bytes + bt_bytes;
            rs.exc = (i64 *)(pw + ctl_bytes + bt_bytes + rec_bytes);
            RFB_CUDA(cudaMemsetAsync(pw, 0, ctl_bytes + bt_bytes, ctx->stream));
            k_fused_scope_init<<<1, 32, 0, ctx->stream>>>(mm);
            RFB_CHECK_LAUNCH(ctx);
            if (plan.rec_bytes == 4 && plan.kpl == 12) rc = ms_scatter_launch<FS, u32, 12>(ctx, fs, val, n, rs, mm);
            else if (plan.rec_bytes == 4) rc = ms_scatter_launch<FS, u32, 13>(ctx, fs, val, n, rs, mm);
            else rc = ms_scatter_launch<FS, u64, 13>(ctx, fs, val, n, rs, mm);
            if (rc) return rc;
            u32 exc_n = 0;
            RFB_CUDA(cudaMemcpyAsync(&exc_n, rs.exc_count, 4, cudaMemcpyDeviceToHost, ctx->stream));
            rc = d2h_sync(ctx, h, mm, 16);
            if (rc) return rc;
            if (exc_n <= rs.exc_cap && (h[0] > h[1] || buckets_spanned(h[0], h[1], plan.kpl) <= NP)) have_scope = narrowed = true;
            else if (exc_n <= rs.exc_cap) have_scope = true;     // a complete pass: the key bounds are exact, only the partitioning is void
            else { h[0] = hs[0]; h[1] = hs[1]; }                 // aborted early: back to what the sample said
        }
        const bool try_part = !modded && !narrowed && h[0] <= h[1] && (forced == 2 || (i64)((u64)h[1] - (u64)h[0]) >= KP) && (u64)h[1] - (u64)h[0] < (u64)max_range &&
                              parts_of(h[0], h[1]) <= MAX_PARTS;
        if (try_part) {
            const u32 blocks = (u32)((n + PB - 1) / PB) + MAX_PARTS + 1;
            const u32 bt_stride = (u32)((n + PB - 1) / PB) + 1;
            const size_t bt_bytes = align256((size_t)MAX_PARTS * bt_stride * 4);
            const size_t ctl_bytes = align256((size_t)(MAX_PARTS + 1) * PCUR_STRIDE * 4);
            rc = rfb_ensure_work(ctx, p_acc_bytes + ctl_bytes + bt_bytes + (size_t)blocks * PB * 10, &w);
            if (rc) return rc;
            char *pw = (char *)w + p_acc_bytes;
            ps.cursor = (u32 *)pw;
            ps.next_block = ps.cursor + MAX_PARTS * PCUR_STRIDE;
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
    if (range <= 0 || range > DENSE_RANGE_MAX) return RFB_SPARSE_DOMAIN;   // open-addressing tables instead of direct addressing (k_hash_group.cu)
    const i64 kbase = kmin & ~(i64)(KP - 1);            // floor to a multiple of KP (two's complement)
    const i64 P = parts_of(kmin, h[1]);                 // partitions of KP consecutive keys
    int strategy = 3;
    if (narrowed) strategy = 4;
    else if (range <= KP && (n >= 65536 || forced == 1)) strategy = 1;
    else if (scattered && P <= MAX_PARTS) strategy = 2;
    if (forced == 3) strategy = 3;
    const bool part_layout = strategy == 2 || strategy == 4;   // accumulators sized for the largest range the partitioned passes accept

    const size_t b8 = part_layout ? pb8 : align256((size_t)range * 8), b4 = part_layout ? pb4 : align256((size_t)range * 4);
    if (!scattered && !plan.rec_bytes) {
        rc = rfb_ensure_work(ctx, 3 * b8 + b4 + scan::tiles_bytes(tiles_max), &w);
        if (rc) return rc;
    }
    Accums a;
    a.first_row = (u64 *)w;
    a.sum = (u64 *)((char *)w + b8);
    a.cnt = (u64 *)((char *)w + 2 * b8);
    a.has_null = (u32 *)((char *)w + 3 * b8);
    if ((scattered || plan.rec_bytes) && !part_layout) {   // the sample misjudged the range: the accumulator layout of the other strategies must fit
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
    } else if (strategy == 4) {
        const i64 kb = kmin & ~(((i64)1 << plan.kpl) - 1);
        const int np = (int)buckets_spanned(kmin, h[1], plan.kpl);
        if (plan.rec_bytes == 4 && plan.kpl == 12) rc = ms_accum_launch<u32, 12>(ctx, rs, np, kb, kmin, a);
        else if (plan.rec_bytes == 4) rc = ms_accum_launch<u32, 13>(ctx, rs, np, kb, kmin, a);
        else rc = ms_accum_launch<u64, 13>(ctx, rs, np, kb, kmin, a);
        if (rc) return rc;
        k_ms_exceptions<<<rfb_grid_for(ctx, rs.exc_cap, THREADS, 2), THREADS, 0, ctx->stream>>>(rs.exc, rs.exc_count, kmin, a, true);
        RFB_CHECK_LAUNCH(ctx);
    } else if (strategy == 2) {
        if (rfb_options()->accum_tma) {                  // RFB_ACCUM_TMA=0: the register-staged kernel (128-bit loads) instead of the TMA ring
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

// carve the narrow path's block store out of a workspace region
size_t narrow_store_bytes(i64 n, int rec_bytes) {
    const u32 blocks = (u32)((n + PB - 1) / PB) + NP + 1, bt_stride = (u32)((n + PB - 1) / PB) + 1;
    i64 cap = n / 64 > 4096 ? n / 64 : 4096;
    if (cap > (1ll << 23)) cap = 1ll << 23;
    return align256((size_t)(NP + 2) * CUR_STRIDE * 4) + align256((size_t)NP * bt_stride * 4) + align256((size_t)blocks * PB * rec_bytes) + align256((size_t)cap * 16);
}
int narrow_store_init(rfb_ctx_t *ctx, char *pw, i64 n, int rec_bytes, RecStore *rs) {
    const u32 blocks = (u32)((n + PB - 1) / PB) + NP + 1, bt_stride = (u32)((n + PB - 1) / PB) + 1;
    i64 cap = n / 64 > 4096 ? n / 64 : 4096;
    if (cap > (1ll << 23)) cap = 1ll << 23;
    const size_t ctl_bytes = align256((size_t)(NP + 2) * CUR_STRIDE * 4), bt_bytes = align256((size_t)NP * bt_stride * 4);
    const size_t rec_total = align256((size_t)blocks * PB * rec_bytes);
    rs->cursor = (u32 *)pw;
    rs->next_block = rs->cursor + NP * CUR_STRIDE;
    rs->exc_count = rs->cursor + (NP + 1) * CUR_STRIDE;
    rs->bt = (u32 *)(pw + ctl_bytes);
    rs->bt_stride = bt_stride;
    rs->exc_cap = (u32)cap;
    rs->rec = pw + ctl_bytes + bt_bytes;
    rs->exc = (i64 *)(pw + ctl_bytes + bt_bytes + rec_total);
    RFB_CUDA(cudaMemsetAsync(pw, 0, ctl_bytes + bt_bytes, ctx->stream));
    return RFB_OK;
}

}  // namespace

// ---- grouped sums over DENSE group ids through the narrow partitioned passes (rfb_aggr_dev at 1e4 .. 2.6e5 groups: two
// shared-memory atomics per row in the accumulate pass instead of two L2 atomics).  sum / cnt / has_null are the caller's
// zero-initialised device-wide arrays of `groups` slots; cnt counts the rows of each group (nulls_count) or its non-null rows.
// *done = false: not applicable (small input, too many groups, unaligned columns, values that fit no record format) — nothing
// was written, the caller takes its own path.
size_t rfb_narrow_sums_bytes(i64 n) { return narrow_store_bytes(n, 8); }

int rfb_narrow_sums(rfb_ctx_t *ctx, const i64 *gid, const i64 *val, i64 n, i64 groups, void *work, u64 *sum, u64 *cnt, u32 *has_null,
                    bool nulls_count, bool *done) {
    *done = false;
    const int forced = group_strategy_forced();
    if (forced == 3 || n < part_min_rows() || n >= 0xF0000000ll || groups > ((i64)NP << 13) || !aligned16(gid) || !aligned16(val)) return RFB_OK;
    typedef FusedSrc<i64, i64, false> FS;
    FS fs{gid, nullptr, PredRange{0, 0, 0, 0}};
    i64 *mm = (i64 *)((char *)ctx->d_scratch + 32768), *census = mm + 8;
    RFB_CUDA(cudaMemsetAsync(census, 0, 32, ctx->stream));
    const i64 win = 65536;
    const i64 starts[3] = {0, n / 2 > win ? n / 2 : 0, n > win ? n - win : 0};
    for (int s = 0; s < 3; s++) {
        const i64 r0 = starts[s], r1 = r0 + win < n ? r0 + win : n;
        k_val_census_rows<FS><<<64, THREADS, 0, ctx->stream>>>(fs, val, r0, r1, census);
        RFB_CHECK_LAUNCH(ctx);
    }
    i64 c[4], h[2];
    int rc = d2h_sync(ctx, c, census, 32);
    if (rc) return rc;
    const i64 tol = c[3] / 64;
    NarrowPlan plan{0, 0};
    if (groups <= ((i64)NP << 12) && c[1] <= tol) plan = NarrowPlan{4, 12};
    else if (c[0] <= tol) plan = NarrowPlan{4, 13};
    else if (c[2] <= tol) plan = NarrowPlan{8, 13};
    else return RFB_OK;
    RecStore rs{};
    rc = narrow_store_init(ctx, (char *)work, n, plan.rec_bytes, &rs);
    if (rc) return rc;
    k_fused_scope_init<<<1, 32, 0, ctx->stream>>>(mm);
    RFB_CHECK_LAUNCH(ctx);
    if (plan.rec_bytes == 4 && plan.kpl == 12) rc = ms_scatter_launch<FS, u32, 12>(ctx, fs, val, n, rs, mm);
    else if (plan.rec_bytes == 4) rc = ms_scatter_launch<FS, u32, 13>(ctx, fs, val, n, rs, mm);
    else rc = ms_scatter_launch<FS, u64, 13>(ctx, fs, val, n, rs, mm);
    if (rc) return rc;
    u32 exc_n = 0;
    RFB_CUDA(cudaMemcpyAsync(&exc_n, rs.exc_count, 4, cudaMemcpyDeviceToHost, ctx->stream));
    rc = d2h_sync(ctx, h, mm, 16);
    if (rc) return rc;
    if (exc_n > rs.exc_cap) return RFB_OK;               // the sample misjudged the values: aborted early, nothing accumulated yet
    if (h[0] <= h[1] && (h[0] < 0 || h[1] >= groups)) { rfb_set_error("grouped aggregate: group id outside [0, %lld)", (long long)groups); return RFB_ERR_ARG; }
    Accums a;
    a.first_row = nullptr; a.sum = sum; a.cnt = cnt; a.has_null = has_null;
    const int np = (int)((groups + ((i64)1 << plan.kpl) - 1) >> plan.kpl);
    if (plan.rec_bytes == 4 && plan.kpl == 12) rc = ms_accum_launch<u32, 12>(ctx, rs, np, 0, 0, a);
    else if (plan.rec_bytes == 4) rc = ms_accum_launch<u32, 13>(ctx, rs, np, 0, 0, a);
    else rc = ms_accum_launch<u64, 13>(ctx, rs, np, 0, 0, a);
    if (rc) return rc;
    k_ms_exceptions<<<rfb_grid_for(ctx, rs.exc_cap, THREADS, 2), THREADS, 0, ctx->stream>>>(rs.exc, rs.exc_count, 0, a, nulls_count);
    RFB_CHECK_LAUNCH(ctx);
    *done = true;
    return RFB_OK;
}

extern "C" int rfb_group_sum_count_dev(rfb_ctx_t *ctx, int key_type, const void *keys, const int64_t *val, int64_t n,
                                       int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k,
                                       int64_t max_groups, int64_t *out_keys, int64_t *out_sums, int64_t *out_counts,
                                       int64_t *groups) {
    RFB_ARG(ctx && groups && n >= 0 && max_groups >= 0 && ((keys && val) || n == 0) && (!pred || k), "rfb_group_sum_count_dev");
    RFB_ARG((out_keys && out_sums && out_counts) || max_groups == 0, "rfb_group_sum_count_dev: outputs");
    *groups = 0;
    if (n == 0) return RFB_OK;
    int rc;
    switch (rfb_kind_of(key_type)) {
        case K_I32: rc = fused_key<i32>(ctx, keys, cmp_op, pred_type, pred, k, val, n, max_groups, out_keys, out_sums, out_counts, groups); break;
        case K_I64: rc = fused_key<i64>(ctx, keys, cmp_op, pred_type, pred, k, val, n, max_groups, out_keys, out_sums, out_counts, groups); break;
        default: rfb_set_error("fused group-by: key type %d (I32 or I64 keys)", key_type); return RFB_ERR_TYPE;
    }
    if (rc == RFB_SPARSE_DOMAIN)
        rc = rfb_hash_group_sum_count(ctx, key_type, keys, val, n, cmp_op, pred_type, pred, k, max_groups, out_keys, out_sums, out_counts, groups);
    return rc;
}
