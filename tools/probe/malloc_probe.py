#!/usr/bin/env python
"""How long does device memory allocation take on this box?  cudaMalloc / cudaFree and cudaMallocAsync (default pool with an
unlimited release threshold) for a few sizes, through the library's own allocator entry points and raw cudart."""
import ctypes as C
import json
import time

import torch

torch.cuda.init()
torch.zeros(1, device="cuda")
rt = C.CDLL("libcudart.so.12")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaFree.argtypes = [C.c_void_p]
rt.cudaMallocAsync.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_void_p]
rt.cudaFreeAsync.argtypes = [C.c_void_p, C.c_void_p]
rt.cudaDeviceGetDefaultMemPool.argtypes = [C.POINTER(C.c_void_p), C.c_int]
rt.cudaMemPoolSetAttribute.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
pool = C.c_void_p()
rt.cudaDeviceGetDefaultMemPool(C.byref(pool), 0)
thr = C.c_uint64(2 ** 64 - 1)
print("set threshold rc", rt.cudaMemPoolSetAttribute(pool, 4, C.byref(thr)))     # cudaMemPoolAttrReleaseThreshold = 4

for mb in (100, 800, 4000):
    n = mb << 20
    row = {"MB": mb}
    for rep in range(3):
        p = C.c_void_p()
        t = time.perf_counter(); rc = rt.cudaMalloc(C.byref(p), n); rt.cudaDeviceSynchronize(); a = (time.perf_counter() - t) * 1e3
        t = time.perf_counter(); rt.cudaFree(p); f = (time.perf_counter() - t) * 1e3
        row["cudaMalloc_%d" % rep] = round(a, 2); row["cudaFree_%d" % rep] = round(f, 2)
    for rep in range(3):
        p = C.c_void_p()
        t = time.perf_counter(); rc = rt.cudaMallocAsync(C.byref(p), n, None); rt.cudaDeviceSynchronize(); a = (time.perf_counter() - t) * 1e3
        t = time.perf_counter(); rt.cudaFreeAsync(p, None); rt.cudaDeviceSynchronize(); f = (time.perf_counter() - t) * 1e3
        row["mallocAsync_%d" % rep] = round(a, 2); row["freeAsync_%d" % rep] = round(f, 2)
    print(json.dumps(row), flush=True)
