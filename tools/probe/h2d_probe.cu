// h2d_probe.cu — how should pageable host columns be shipped? (not part of the product)
//   pageable cudaMemcpy vs cudaHostRegister + DMA vs N copier threads filling pinned staging buffers + async DMA
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }
struct Job { const char *src; char *dst; size_t bytes; };
static void *copier(void *p) { Job *j = (Job *)p; memcpy(j->dst, j->src, j->bytes); return nullptr; }

int main() {
    const size_t bytes = 4ull << 30;
    char *h = (char *)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    memset(h, 1, bytes);
    char *d; cudaMalloc(&d, bytes);
    cudaStream_t s; cudaStreamCreate(&s);
    for (int r = 0; r < 2; r++) { double t0 = now(); cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice); double t1 = now(); printf("pageable cudaMemcpy            : %6.1f GB/s\n", bytes / (t1 - t0) / 1e9); }
    { double t0 = now(); cudaError_t e = cudaHostRegister(h, bytes, cudaHostRegisterDefault); double t1 = now();
      printf("cudaHostRegister 4 GiB         : %6.3f s (%s) = %5.1f GB/s\n", t1 - t0, cudaGetErrorString(e), bytes / (t1 - t0) / 1e9);
      double t2 = now(); cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s); cudaStreamSynchronize(s); double t3 = now();
      printf("registered DMA                 : %6.1f GB/s\n", bytes / (t3 - t2) / 1e9);
      double t4 = now(); cudaHostUnregister(h); double t5 = now(); printf("cudaHostUnregister             : %6.3f s\n", t5 - t4); }
    const size_t chunk = 16ull << 20; const int NB = 4;
    char *stage[NB]; cudaEvent_t done[NB];
    for (int i = 0; i < NB; i++) { cudaHostAlloc(&stage[i], chunk, cudaHostAllocDefault); cudaEventCreate(&done[i]); }
    for (int threads : {1, 2, 4, 8, 12}) {
        double t0 = now();
        size_t nchunks = bytes / chunk;
        for (size_t c = 0; c < nchunks; c++) {
            int b = c % NB;
            if (c >= NB) cudaEventSynchronize(done[b]);
            pthread_t th[16]; Job jobs[16]; size_t per = chunk / threads;
            for (int t = 0; t < threads; t++) { jobs[t] = {h + c * chunk + t * per, stage[b] + t * per, per}; if (t) pthread_create(&th[t], nullptr, copier, &jobs[t]); }
            copier(&jobs[0]);
            for (int t = 1; t < threads; t++) pthread_join(th[t], nullptr);
            cudaMemcpyAsync(d + c * chunk, stage[b], chunk, cudaMemcpyHostToDevice, s);
            cudaEventRecord(done[b], s);
        }
        cudaStreamSynchronize(s);
        double t1 = now();
        printf("staged, %2d copier thread(s)    : %6.1f GB/s\n", threads, bytes / (t1 - t0) / 1e9);
    }
    return 0;
}
