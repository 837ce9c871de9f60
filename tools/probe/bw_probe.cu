// bw_probe.cu — design-space probe for the streaming filter+sum kernel (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o bw_probe bw_probe.cu && ./bw_probe [rows]
// Times variants of `sum of x where x < k` over an int64 column: threads/CTA, 16-byte loads in flight per thread,
// CTAs per SM, tile order (interleaved tiles vs one contiguous span per CTA), lean per-row arithmetic.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef int64_t i64; typedef uint64_t u64; typedef uint32_t u32;
struct __align__(16) v16 { u64 lo, hi; };
__device__ __forceinline__ v16 ld16(const void *p) { v16 r; asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "l"(p)); return r; }
__device__ __forceinline__ v16 ld16_plain(const void *p) { v16 r; asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "l"(p)); return r; }
__device__ __forceinline__ u64 splitmix64(u64 seed, u64 i) { u64 z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
__global__ void fill(i64 *x, i64 n) { for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) x[i] = (i64)(splitmix64(42, i) % (1ull << 40)); }

// MODE 0: tiles interleaved across CTAs (tile t -> CTA t % grid); MODE 1: CTA owns a contiguous span
// LEAN 0: biased-range predicate + null check + 64-bit counters (product arithmetic); 1: signed compare, 32-bit counter
template <int THREADS, int UNROLL, int BPS, int MODE, int LEAN, int PLAIN>
__global__ void __launch_bounds__(THREADS, BPS) k(const i64 *__restrict__ x, i64 chunks, u64 lo, u64 span, i64 kconst, u64 *out) {
    constexpr int TILE = THREADS * UNROLL;
    const i64 tiles = chunks / TILE;  // probe: full tiles only
    u64 sum = 0, rows = 0, nn = 0; u32 rows32 = 0;
    i64 t0, t1, ts;
    if (MODE == 0) { t0 = blockIdx.x; t1 = tiles; ts = gridDim.x; }
    else { i64 per = (tiles + gridDim.x - 1) / gridDim.x; t0 = blockIdx.x * per; t1 = t0 + per < tiles ? t0 + per : tiles; ts = 1; }
    for (i64 t = t0; t < t1; t += ts) {
        const char *base = (const char *)x + (t * TILE + threadIdx.x) * 16;
        v16 v[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; j++) v[j] = PLAIN ? ld16_plain(base + (i64)j * THREADS * 16) : ld16(base + (i64)j * THREADS * 16);
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
            if (LEAN) {
                const bool a = (i64)v[j].lo < kconst, b = (i64)v[j].hi < kconst;
                sum += (a ? v[j].lo : 0) + (b ? v[j].hi : 0);
                rows32 += (u32)a + (u32)b;
            } else {
                const u64 ka = v[j].lo ^ 0x8000000000000000ULL, kb = v[j].hi ^ 0x8000000000000000ULL;
                const bool a = (ka - lo) <= span, b = (kb - lo) <= span;
                const bool oa = a && v[j].lo != 0x8000000000000000ULL, ob = b && v[j].hi != 0x8000000000000000ULL;
                rows += (u64)a + (u64)b; nn += (u64)oa + (u64)ob;
                sum += (oa ? v[j].lo : 0) + (ob ? v[j].hi : 0);
            }
        }
    }
    rows += rows32;
    for (int d = 16; d > 0; d >>= 1) { sum += __shfl_down_sync(~0u, sum, d); rows += __shfl_down_sync(~0u, rows, d); nn += __shfl_down_sync(~0u, nn, d); }
    if ((threadIdx.x & 31) == 0) { atomicAdd((unsigned long long *)out, (unsigned long long)sum); atomicAdd((unsigned long long *)out + 1, (unsigned long long)rows); atomicAdd((unsigned long long *)out + 2, (unsigned long long)nn); }
}

template <int THREADS, int UNROLL, int BPS, int MODE, int LEAN, int PLAIN>
void run(const char *name, const i64 *x, i64 n, u64 *out, int sms) {
    const i64 chunks = n / 2;
    const u64 kk = (1ull << 39) ^ 0x8000000000000000ULL;
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    float best = 1e9f; u64 h[3];
    for (int r = 0; r < 6; r++) {
        cudaMemset(out, 0, 24);
        cudaEventRecord(s);
        k<THREADS, UNROLL, BPS, MODE, LEAN, PLAIN><<<sms * BPS, THREADS>>>(x, chunks, 0, kk - 1, (i64)(1ull << 39), out);
        cudaEventRecord(e); cudaEventSynchronize(e);
        float ms; cudaEventElapsedTime(&ms, s, e);
        if (r && ms < best) best = ms;
    }
    cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
    printf("%-44s thr=%4d unroll=%2d cta/sm=%d mode=%d lean=%d plain=%d : %7.3f ms  %7.1f GB/s   sum=%llu rows=%llu %s\n", name, THREADS, UNROLL, BPS, MODE, LEAN, PLAIN, best,
           8.0 * n / best / 1e6, (unsigned long long)h[0], (unsigned long long)h[1], cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv) {
    i64 n = argc > 1 ? atoll(argv[1]) : 1000000000ll;
    n = n / (1 << 16) * (1 << 16);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    i64 *x; u64 *out; cudaMalloc(&x, n * 8); cudaMalloc(&out, 64);
    fill<<<p.multiProcessorCount * 8, 256>>>(x, n); cudaDeviceSynchronize();
    const int sms = p.multiProcessorCount;
    run<256, 8, 4, 0, 0, 0>("product shape", x, n, out, sms);
    run<256, 8, 4, 0, 1, 0>("lean arithmetic", x, n, out, sms);
    run<256, 8, 4, 1, 0, 0>("contiguous span per CTA", x, n, out, sms);
    run<256, 4, 8, 0, 0, 0>("unroll 4, 8 CTAs/SM", x, n, out, sms);
    run<256, 4, 4, 0, 0, 0>("unroll 4, 4 CTAs/SM", x, n, out, sms);
    run<256, 16, 2, 0, 0, 0>("unroll 16, 2 CTAs/SM", x, n, out, sms);
    run<512, 8, 2, 0, 0, 0>("512 threads, unroll 8, 2 CTAs/SM", x, n, out, sms);
    run<512, 4, 4, 0, 0, 0>("512 threads, unroll 4, 4 CTAs/SM", x, n, out, sms);
    run<1024, 4, 2, 0, 0, 0>("1024 threads, unroll 4, 2 CTAs/SM", x, n, out, sms);
    run<1024, 8, 1, 0, 0, 0>("1024 threads, unroll 8, 1 CTA/SM", x, n, out, sms);
    run<128, 8, 8, 0, 0, 0>("128 threads, unroll 8, 8 CTAs/SM", x, n, out, sms);
    run<256, 8, 4, 0, 0, 1>("plain ld.global (L1 allocate)", x, n, out, sms);
    run<256, 8, 6, 0, 1, 0>("lean, 6 CTAs/SM", x, n, out, sms);
    run<256, 8, 8, 0, 1, 0>("lean, 8 CTAs/SM (32 regs)", x, n, out, sms);
    run<256, 4, 8, 0, 1, 0>("lean, unroll 4, 8 CTAs/SM", x, n, out, sms);
    run<256, 2, 8, 0, 1, 0>("lean, unroll 2, 8 CTAs/SM", x, n, out, sms);
    run<512, 4, 4, 1, 1, 0>("lean, contiguous, 512 thr", x, n, out, sms);
    return 0;
}
