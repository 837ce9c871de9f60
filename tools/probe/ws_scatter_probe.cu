// ws_scatter_probe.cu — design-space probe for the group-by partition pass (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ws_scatter_probe ws_scatter_probe.cu && ./ws_scatter_probe [rows]
// Question: how fast is a WARP-PRIVATE multisplit — every warp ranks 32 x R rows with ballots (as k_ms_scatter does), reserves
// its runs with one global atomic per partition (lane = partition) and writes the packed 32-bit records from REGISTERS straight
// to their global positions — compared with k_ms_scatter's CTA tiles (shared-memory staging, two barriers, linear write-out:
// 3.83 ms per 1e9 rows of (i32 key, i64 value), 1e5 keys in 25 partitions of 4096)?  No staging, no barriers, about half the
// instructions per row; the price is scattered 4-byte stores (every store instruction touches ~18 runs) and 8x more cursor
// atomics, spread over S sub-streams per partition.  Partitions are fixed-capacity regions here (the product hands out blocks).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef int64_t i64; typedef uint64_t u64; typedef uint32_t u32;
constexpr int NP = 32, KPL = 12, VB = 20, CUR_STRIDE = 64;
__device__ __forceinline__ u64 splitmix64(u64 seed, u64 i) { u64 z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
__global__ void fill(int *k, i64 *v, i64 n) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) { k[i] = (int)(splitmix64(7, i) % 100000); v[i] = (i64)(splitmix64(43, i) % (1ull << 20)); }
}
__device__ __forceinline__ u32 ballot_bits(u32 x, u32 mask) {
    u32 r;
    asm volatile("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %1, %2;\nsetp.ne.u32 p, t, 0;\nvote.sync.ballot.b32 %0, p, 0xffffffff;\n}" : "=r"(r) : "r"(x), "r"(mask));
    return r;
}
__device__ __forceinline__ uint4 ldg16(const void *p) { uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }

// lane owns R consecutive rows of the warp tile (R = 8: two 16-byte key loads, four 16-byte value loads)
template <int T, int CTAS, int S, int PRE>
__global__ void __launch_bounds__(T, CTAS) ws_scatter(const int *__restrict__ keys, const i64 *__restrict__ val, i64 wtiles, u32 *cursor, u32 *out, u32 cap, i64 *mm) {
    constexpr int R = 8;
    const int lane = threadIdx.x & 31;
    const u32 lt_mask = (1u << lane) - 1u;
    const i64 gw = ((i64)blockIdx.x * T + threadIdx.x) >> 5, nw = ((i64)gridDim.x * T) >> 5;
    const u32 sub = blockIdx.x & (S - 1);
    u32 cb[5];
#pragma unroll
    for (int b = 0; b < 5; b++) cb[b] = ((lane >> b) & 1) ? 0u : 0xFFFFFFFFu;
    int lo = 0x7FFFFFFF, hi = (int)0x80000000;
    uint4 kq[2], vq[4];
    auto load = [&](i64 t) {
        const char *kp = (const char *)(keys + t * (32 * R) + lane * R);
        const char *vp = (const char *)(val + t * (32 * R) + lane * R);
        kq[0] = ldg16(kp); kq[1] = ldg16(kp + 16);
#pragma unroll
        for (int q = 0; q < 4; q++) vq[q] = ldg16(vp + 16 * q);
    };
    if (PRE && gw < wtiles) load(gw);
    for (i64 t = gw; t < wtiles; t += nw) {
        if (!PRE) load(t);
        int k[R]; u32 vlo[R], vhi[R];
        k[0] = kq[0].x; k[1] = kq[0].y; k[2] = kq[0].z; k[3] = kq[0].w; k[4] = kq[1].x; k[5] = kq[1].y; k[6] = kq[1].z; k[7] = kq[1].w;
#pragma unroll
        for (int q = 0; q < 4; q++) { vlo[2 * q] = vq[q].x; vhi[2 * q] = vq[q].y; vlo[2 * q + 1] = vq[q].z; vhi[2 * q + 1] = vq[q].w; }
        if (PRE && t + nw < wtiles) load(t + nw);
        u32 rec[R], pp[R], wcount = 0, allok = 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < R; j++) {
            lo = min(lo, k[j]); hi = max(hi, k[j]);
            const bool ok = vhi[j] == 0 && vlo[j] < (1u << VB);
            const u32 kb = (u32)k[j], part = (kb >> KPL) & (NP - 1);
            rec[j] = ((kb & ((1u << KPL) - 1u)) << VB) | vlo[j];
            u32 pl = __ballot_sync(0xffffffffu, ok);
            allok &= pl;
#pragma unroll
            for (int b = 0; b < 5; b++) pl &= ballot_bits(kb, 1u << (KPL + b)) ^ cb[b];
            const u32 peers = __shfl_sync(0xffffffffu, pl, part);
            const u32 before = __shfl_sync(0xffffffffu, wcount, part);
            wcount += __popc(pl);
            pp[j] = ((before + __popc(peers & lt_mask)) << 8) | (ok ? part : 32u);
        }
        // lane = partition: reserve the warp's run in the partition's sub-stream
        u32 start = 0;
        if (wcount) start = atomicAdd(&cursor[(lane * S + sub) * CUR_STRIDE], wcount);
        const u32 base = (lane * S + sub) * cap + start;      // probe: one fixed region per (partition, sub-stream)
#pragma unroll
        for (int j = 0; j < R; j++) {
            const u32 b = __shfl_sync(0xffffffffu, base, pp[j] & 31u);
            if (!(pp[j] & 32u)) out[b + (pp[j] >> 8)] = rec[j];
        }
        if (allok != 0xFFFFFFFFu) atomicAdd((unsigned long long *)&mm[2], 1ull);   // the exception path of the product (never taken here)
    }
    for (int d = 16; d > 0; d >>= 1) { lo = min(lo, __shfl_down_sync(~0u, lo, d)); hi = max(hi, __shfl_down_sync(~0u, hi, d)); }
    if (lane == 0) { atomicMin((long long *)&mm[0], (long long)lo); atomicMax((long long *)&mm[1], (long long)hi); }
}

// checksum of what was written: sum of the value fields and number of records over all streams
__global__ void check(const u32 *out, const u32 *cursor, int streams, u32 cap, u64 *res) {
    u64 s = 0, c = 0;
    for (int st = blockIdx.y; st < streams; st += gridDim.y) {
        const u32 m = cursor[st * CUR_STRIDE];
        for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) { s += out[(size_t)st * cap + i] & ((1u << VB) - 1u); c++; }
    }
    for (int d = 16; d > 0; d >>= 1) { s += __shfl_down_sync(~0u, s, d); c += __shfl_down_sync(~0u, c, d); }
    if ((threadIdx.x & 31) == 0) { atomicAdd((unsigned long long *)res, (unsigned long long)s); atomicAdd((unsigned long long *)res + 1, (unsigned long long)c); }
}
__global__ void vsum(const i64 *v, i64 n, u64 *res) {
    u64 s = 0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) s += (u64)v[i];
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(~0u, s, d);
    if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)res + 2, (unsigned long long)s);
}

template <int T, int CTAS, int S, int PRE>
void run(const char *name, const int *k, const i64 *v, i64 n, u32 *cursor, u32 *out, u32 cap, i64 *mm, u64 *res, int sms, u64 want) {
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    float best = 1e9f;
    const i64 wtiles = n / 256;
    for (int r = 0; r < 5; r++) {
        cudaMemset(cursor, 0, NP * 8 * CUR_STRIDE * 4);
        cudaEventRecord(s);
        ws_scatter<T, CTAS, S, PRE><<<sms * CTAS, T>>>(k, v, wtiles, cursor, out, cap, mm);
        cudaEventRecord(e);
        cudaEventSynchronize(e);
        float ms; cudaEventElapsedTime(&ms, s, e);
        if (ms < best) best = ms;
    }
    cudaMemset(res, 0, 16);
    check<<<dim3(64, NP * S), 256>>>(out, cursor, NP * S, cap, res);
    u64 h[3]; cudaMemcpy(h, res, 24, cudaMemcpyDeviceToHost);
    cudaError_t err = cudaGetLastError();
    printf("%-34s %7.3f ms  %6.0f GB/s (16 B/row)  records %llu value sum %s  %s\n", name, best, 16.0 * wtiles * 256 / best / 1e6, (unsigned long long)h[1],
           h[0] == want && h[1] == (u64)wtiles * 256 ? "ok" : "MISMATCH", err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main(int argc, char **argv) {
    const i64 n = argc > 1 ? atoll(argv[1]) : 1000000000ll;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int *k; i64 *v; u32 *cursor, *out; i64 *mm; u64 *res;
    const int nparts = (99999 >> KPL) + 1;                   // 25 partitions of 4096 keys
    const u32 cap = (u32)(n / nparts + (1 << 20));           // per-partition region, uniform keys
    cudaMalloc(&k, n * 4); cudaMalloc(&v, n * 8); cudaMalloc(&cursor, NP * 8 * CUR_STRIDE * 4); cudaMalloc(&mm, 64); cudaMalloc(&res, 64);
    if (cudaMalloc(&out, ((size_t)nparts * cap + (size_t)nparts * 8 * (1 << 18) + (1 << 20)) * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    fill<<<sms * 8, 256>>>(k, v, n);
    cudaMemset(res, 0, 64);
    const i64 nfull = n / 256 * 256;
    vsum<<<sms * 8, 256>>>(v, nfull, res);
    u64 h[3]; cudaMemcpy(h, res, 24, cudaMemcpyDeviceToHost);
    const u64 want = h[2];
    // streams are laid out [partition * S + sub] * cap; with S sub-streams a region needs cap / S: shrink cap per variant
#define RUN(T, C, S, P) run<T, C, S, P>("T=" #T " CTAS=" #C " S=" #S " PRE=" #P, k, v, n, cursor, out, cap / S + (1 << 18), mm, res, sms, want)
    RUN(256, 3, 1, 0); RUN(256, 3, 4, 0); RUN(256, 3, 8, 0);
    RUN(256, 4, 4, 0); RUN(256, 4, 8, 0); RUN(128, 8, 4, 0); RUN(512, 2, 4, 0);
    RUN(256, 3, 4, 1); RUN(256, 4, 4, 1);
    return 0;
}
