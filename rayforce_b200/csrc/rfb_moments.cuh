// rfb_moments.cuh — per-group fp64 moments (sum, sum of squares, non-null count) for the grouped f64 sum / avg and the
// grouped deviation (k_group.cu, k_stats.cu).
//
// fp64 accumulators cannot use the 32-bit shared-atomic trick of the integer sums, and an fp64 atomicAdd — shared (a
// compare-and-swap loop on sm_100a) or device-wide (L2) — collapses when many lanes hit few addresses: 100 groups over
// 1e9 rows took 312 ms with one L2 atomic per row.  Low-cardinality groupings therefore accumulate in shared memory:
//   groups <= MOM_WARP_GROUPS  one private copy PER WARP: the only lanes that ever collide on an address are lanes of the same
//                              warp instruction that hold the same group (a couple of CAS retries), never other warps
//   groups <= MOM_CTA_GROUPS   one copy per CTA (32 lanes spread over >= 256 addresses: collisions are rare)
//   above                      device-wide atomics, as before (addresses are many, contention is low)
// and are merged into the device-wide arrays once per CTA, warp copies in index order.
#pragma once
#include "rfb_common.cuh"

namespace moments {

constexpr int THREADS = 256, WARPS = THREADS / 32;
constexpr int MOM_WARP_GROUPS = 256, MOM_CTA_GROUPS = 2048;

template <typename V> __device__ __forceinline__ bool value_of(V x, f64 &v) {
    if (Elem<V>::is_null(x)) return false;
    v = (f64)x;
    return true;
}

// gsum / gsq / gcnt: device-wide per-group arrays (gsq only when SQ); gnull (optional): set to 1 for groups that saw a null
template <typename V, bool SQ, bool SHARED>
__global__ void __launch_bounds__(THREADS, 4)
k_group_moments(const V *__restrict__ val, const i64 *__restrict__ filter, const i64 *__restrict__ gid, i64 n, int groups, int copies,
                f64 *gsum, f64 *gsq, unsigned long long *gcnt, u32 *gnull) {
    extern __shared__ f64 s_mom[];
    f64 *ssum = nullptr, *ssq = nullptr;
    u32 *scnt = nullptr;
    if constexpr (SHARED) {
        const int slots = groups * copies;
        ssum = s_mom;
        ssq = s_mom + slots;                                   // (unused region when !SQ: not allocated, see moment_smem_bytes)
        scnt = (u32 *)(s_mom + (SQ ? 2 : 1) * slots);
        for (int s = threadIdx.x; s < slots; s += THREADS) { ssum[s] = 0.0; if (SQ) ssq[s] = 0.0; scnt[s] = 0; }
        __syncthreads();
    }
    const int mine = SHARED ? (copies > 1 ? (threadIdx.x >> 5) * groups : 0) : 0;
    for (i64 i = (i64)blockIdx.x * THREADS + threadIdx.x; i < n; i += (i64)gridDim.x * THREADS) {
        const V x = filter ? __ldg(val + ld_stream(filter + i)) : ld_stream(val + i);
        const i64 g = ld_stream(gid + i);
        f64 v;
        if (!value_of<V>(x, v)) { if (gnull) gnull[g] = 1u; continue; }
        if constexpr (SHARED) {
            atomicAdd(ssum + mine + g, v);
            if (SQ) atomicAdd(ssq + mine + g, __dmul_rn(v, v));
            atomicAdd(scnt + mine + g, 1u);
        } else {
            atomicAdd(gsum + g, v);
            if (SQ) atomicAdd(gsq + g, __dmul_rn(v, v));
            atomicAdd(gcnt + g, 1ULL);
        }
    }
    if constexpr (SHARED) {
        __syncthreads();
        for (int g = threadIdx.x; g < groups; g += THREADS) {
            f64 s = 0.0, q = 0.0;
            u32 c = 0;
            for (int k = 0; k < copies; k++) { s = __dadd_rn(s, ssum[k * groups + g]); if (SQ) q = __dadd_rn(q, ssq[k * groups + g]); c += scnt[k * groups + g]; }
            if (!c) continue;
            atomicAdd(gsum + g, s);
            if (SQ) atomicAdd(gsq + g, q);
            atomicAdd(gcnt + g, (unsigned long long)c);
        }
    }
}

// gsum / gsq / gcnt must be zeroed by the caller.  len / grid rows per CTA must stay below 2^32 (32-bit shared counts).
template <typename V, bool SQ>
int launch(rfb_ctx_t *ctx, const void *val, const i64 *filter, const i64 *gid, i64 len, i64 groups, f64 *gsum, f64 *gsq,
           unsigned long long *gcnt, u32 *gnull) {
    if (len <= 0) return RFB_OK;
    const int grid = rfb_grid_for(ctx, len, THREADS * 4, 4);
    if (groups <= MOM_CTA_GROUPS && len >= 65536) {
        const int copies = groups <= MOM_WARP_GROUPS ? WARPS : 1;
        const size_t smem = (size_t)groups * copies * ((SQ ? 16 : 8) + 4);
        RFB_CUDA(cudaFuncSetAttribute(k_group_moments<V, SQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024));
        k_group_moments<V, SQ, true><<<grid, THREADS, smem, ctx->stream>>>((const V *)val, filter, gid, len, (int)groups, copies, gsum, gsq, gcnt, gnull);
    } else {
        k_group_moments<V, SQ, false><<<grid, THREADS, 0, ctx->stream>>>((const V *)val, filter, gid, len, (int)groups, 1, gsum, gsq, gcnt, gnull);
    }
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}

}  // namespace moments
