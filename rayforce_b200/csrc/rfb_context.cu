// rfb_context.cu — context, error reporting, device/pinned memory and column shipping for librfb200.so.
#include <stdarg.h>

#include "rfb_common.cuh"

static thread_local char g_err[512] = "";

void rfb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int rfb_cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    rfb_set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    cudaGetLastError();  // clear the sticky-less error state
    return (e == cudaErrorMemoryAllocation) ? RFB_ERR_NOMEM : RFB_ERR_CUDA;
}

int rfb_ensure_work(rfb_ctx_t *ctx, size_t bytes, void **out) {
    if (bytes > ctx->work_bytes) {
        if (ctx->d_work) {
            RFB_CUDA(cudaStreamSynchronize(ctx->stream));
            RFB_CUDA(cudaFree(ctx->d_work));
            ctx->d_work = nullptr;
            ctx->work_bytes = 0;
        }
        size_t want = bytes + (bytes >> 3) + (1 << 20);
        RFB_CUDA(cudaMalloc(&ctx->d_work, want));
        ctx->work_bytes = want;
    }
    *out = ctx->d_work;
    return RFB_OK;
}

int rfb_ensure_aux(rfb_ctx_t *ctx, size_t bytes, void **out) {
    if (bytes > ctx->aux_bytes) {
        if (ctx->d_aux) {
            RFB_CUDA(cudaStreamSynchronize(ctx->stream));
            RFB_CUDA(cudaFree(ctx->d_aux));
            ctx->d_aux = nullptr;
            ctx->aux_bytes = 0;
        }
        size_t want = bytes + (bytes >> 3) + (1 << 20);
        RFB_CUDA(cudaMalloc(&ctx->d_aux, want));
        ctx->aux_bytes = want;
    }
    *out = ctx->d_aux;
    return RFB_OK;
}

int rfb_ensure_aux2(rfb_ctx_t *ctx, size_t bytes, void **out) {
    if (bytes > ctx->aux2_bytes) {
        if (ctx->d_aux2) {
            RFB_CUDA(cudaStreamSynchronize(ctx->stream));
            RFB_CUDA(cudaFree(ctx->d_aux2));
            ctx->d_aux2 = nullptr;
            ctx->aux2_bytes = 0;
        }
        size_t want = bytes + (bytes >> 3) + (1 << 20);
        RFB_CUDA(cudaMalloc(&ctx->d_aux2, want));
        ctx->aux2_bytes = want;
    }
    *out = ctx->d_aux2;
    return RFB_OK;
}

static rfb_options_t g_options;
extern "C" void rfb_options_reload(void) {
    rfb_options_t o;
    const char *s = getenv("RFB_GROUP_STRATEGY");
    o.group_strategy = !s ? 0 : (!strcmp(s, "smem") ? 1 : (!strcmp(s, "part") ? 2 : (!strcmp(s, "l2") ? 3 : (!strcmp(s, "narrow") ? 4 : (!strcmp(s, "hash") ? 5 : 0)))));
    s = getenv("RFB_PART_MIN_ROWS");
    o.part_min_rows = s ? atoll(s) : (1ll << 21);
    s = getenv("RFB_ACCUM_TMA");
    o.accum_tma = !(s && s[0] == '0');
    s = getenv("RFB_SORT_ALGO");
    o.sort_algo = (s && !strcmp(s, "lsd")) ? 1 : 0;       // 0: single-sweep passes (default), 1: histogram + scatter passes
    o.loaded = 1;
    g_options = o;
}
const rfb_options_t *rfb_options() {
    if (!g_options.loaded) rfb_options_reload();
    return &g_options;
}

extern "C" {

int rfb_abi_version(void) { return RFB_ABI_VERSION; }
const char *rfb_last_error(void) { return g_err; }

int rfb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int rfb_ctx_create(int device, rfb_ctx_t **out) {
    RFB_ARG(out, "rfb_ctx_create: out");
    *out = nullptr;
    int n = rfb_device_count();
    if (n <= 0) {
        rfb_set_error("no CUDA device available: librfb200 has no CPU fallback");
        return RFB_ERR_CUDA;
    }
    RFB_ARG(device >= 0 && device < n, "rfb_ctx_create: device index");
    RFB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RFB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        rfb_set_error("device %d is sm_%d%d; librfb200 is built for sm_100a only", device, prop.major, prop.minor);
        return RFB_ERR_CUDA;
    }
    rfb_ctx_t *ctx = (rfb_ctx_t *)calloc(1, sizeof(rfb_ctx_t));
    if (!ctx) return RFB_ERR_NOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    RFB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
    RFB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    // d_scratch layout (64 KB): [0, 256) fold ticket | [256, 32768) per-CTA fold partials, 48 B each, grids of at most 4 CTAs per SM
    // (k_fold.cu) | [32768, 40960) grouping scope / limit / census words (k_group.cu, k_fused_group.cu, k_hash_group.cu) |
    // [40960, 65536) k_stats.cu partials + ticket + result.  The partials region bounds the SM count this build accepts.
    if (256 + (size_t)prop.multiProcessorCount * 4 * 48 > 32768) {
        rfb_set_error("device %d has %d SMs: the per-CTA partials of the fold kernels would overrun their scratch region", device, prop.multiProcessorCount);
        free(ctx);
        return RFB_ERR_CUDA;
    }
    ctx->scratch_bytes = 1 << 16;
    RFB_CUDA(cudaMalloc(&ctx->d_scratch, ctx->scratch_bytes));
    RFB_CUDA(cudaMemset(ctx->d_scratch, 0, ctx->scratch_bytes));
    RFB_CUDA(cudaHostAlloc(&ctx->h_result, RFB_RESULT_SLOTS * sizeof(rfb_fold_t) + 4096, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(ctx->h_result, 0, RFB_RESULT_SLOTS * sizeof(rfb_fold_t) + 4096);
    ctx->h_count = (i64 *)((char *)ctx->h_result + RFB_RESULT_SLOTS * sizeof(rfb_fold_t));
    for (int i = 0; i < RFB_STAGE_BUFS; i++) {
        RFB_CUDA(cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming));
        RFB_CUDA(cudaEventCreateWithFlags(&ctx->ev_kernel[i], cudaEventDisableTiming));
    }
    RFB_CUDA(cudaDeviceSynchronize());
    *out = ctx;
    return RFB_OK;
}

void rfb_ctx_destroy(rfb_ctx_t *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < RFB_STAGE_BUFS; i++) {
        for (int c = 0; c < 2; c++)
            if (ctx->d_stage[c][i]) cudaFree(ctx->d_stage[c][i]);
        cudaEventDestroy(ctx->ev_copy[i]);
        cudaEventDestroy(ctx->ev_kernel[i]);
    }
    rfb_copy_shutdown(ctx);
    rfb_peer_mailbox_release(ctx);
    if (ctx->d_work) cudaFree(ctx->d_work);
    if (ctx->d_aux) cudaFree(ctx->d_aux);
    if (ctx->d_aux2) cudaFree(ctx->d_aux2);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_stream);
    free(ctx);
}

int rfb_ctx_set_stream(rfb_ctx_t *ctx, void *cuda_stream) {
    RFB_ARG(ctx, "rfb_ctx_set_stream");
    RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) {
        cudaStreamDestroy(ctx->stream);
        ctx->own_stream = false;
    }
    if (cuda_stream == nullptr) {
        RFB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    } else {
        ctx->stream = (cudaStream_t)cuda_stream;
    }
    return RFB_OK;
}

int rfb_ctx_set_result_ptr(rfb_ctx_t *ctx, void *device_visible) {
    RFB_ARG(ctx, "rfb_ctx_set_result_ptr");
    ctx->result_override = device_visible;
    return RFB_OK;
}

void *rfb_ctx_stream(rfb_ctx_t *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int rfb_ctx_sm_count(rfb_ctx_t *ctx) { return ctx ? ctx->sm_count : 0; }
int64_t rfb_launch_count(rfb_ctx_t *ctx) { return ctx ? ctx->launches : 0; }

int rfb_sync(rfb_ctx_t *ctx) {
    RFB_ARG(ctx, "rfb_sync");
    RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return RFB_OK;
}

int rfb_dev_alloc(rfb_ctx_t *ctx, size_t bytes, void **dptr) {
    RFB_ARG(ctx && dptr, "rfb_dev_alloc");
    *dptr = nullptr;
    if (bytes == 0) bytes = 16;
    RFB_CUDA(cudaSetDevice(ctx->device));
    RFB_CUDA(cudaMalloc(dptr, bytes));
    return RFB_OK;
}
int rfb_dev_free(rfb_ctx_t *ctx, void *dptr) {
    RFB_ARG(ctx, "rfb_dev_free");
    if (dptr) {
        RFB_CUDA(cudaSetDevice(ctx->device));   /* may be called from a host thread that never touched CUDA */
        RFB_CUDA(cudaFree(dptr));
    }
    return RFB_OK;
}
int rfb_dev_mem_info(rfb_ctx_t *ctx, size_t *free_bytes, size_t *total_bytes) {
    RFB_ARG(ctx && free_bytes && total_bytes, "rfb_dev_mem_info");
    RFB_CUDA(cudaSetDevice(ctx->device));
    RFB_CUDA(cudaMemGetInfo(free_bytes, total_bytes));
    return RFB_OK;
}
int rfb_dev_memset(rfb_ctx_t *ctx, void *dptr, int byte, size_t bytes) {
    RFB_ARG(ctx && (dptr || !bytes), "rfb_dev_memset");
    if (bytes) RFB_CUDA(cudaMemsetAsync(dptr, byte, bytes, ctx->stream));
    return RFB_OK;
}
int rfb_host_pin(void *p, size_t bytes) {
    RFB_ARG(p && bytes, "rfb_host_pin");
    RFB_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return RFB_OK;
}
int rfb_host_unpin(void *p) {
    RFB_ARG(p, "rfb_host_unpin");
    RFB_CUDA(cudaHostUnregister(p));
    return RFB_OK;
}
int rfb_host_alloc_pinned(size_t bytes, void **p) {
    RFB_ARG(p, "rfb_host_alloc_pinned");
    RFB_CUDA(cudaHostAlloc(p, bytes ? bytes : 16, cudaHostAllocPortable));
    return RFB_OK;
}
int rfb_host_free_pinned(void *p) {
    if (p) RFB_CUDA(cudaFreeHost(p));
    return RFB_OK;
}
int rfb_h2d(rfb_ctx_t *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    RFB_ARG(ctx && (bytes == 0 || (dst_dev && src_host)), "rfb_h2d");
    if (bytes) return rfb_copy_h2d(ctx, dst_dev, src_host, bytes, ctx->stream);
    return RFB_OK;
}
int rfb_d2h(rfb_ctx_t *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    RFB_ARG(ctx && (bytes == 0 || (dst_host && src_dev)), "rfb_d2h");
    if (bytes) return rfb_copy_d2h(ctx, dst_host, src_dev, bytes, ctx->stream);
    return RFB_OK;
}

int rfb_d2h_sync_plain(rfb_ctx_t *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    RFB_ARG(ctx && (bytes == 0 || (dst_host && src_dev)), "rfb_d2h_sync_plain");
    if (bytes) {
        RFB_CUDA(cudaStreamSynchronize(ctx->stream));
        RFB_CUDA(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
    }
    return RFB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ synthetic columns generated in HBM

template <typename T>
__global__ void k_fill_splitmix(T *x, i64 n, u64 seed, u64 modulus, i64 offset, i64 null_every, f64 scale) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        u64 r = splitmix64(seed, (u64)i);
        if (modulus) r %= modulus;
        T v;
        if (Elem<T>::kind == K_F64) v = (T)((f64)(i64)r / scale + (f64)offset);
        else v = (T)((i64)r + offset);
        if (null_every > 0 && (i % null_every) == null_every - 1) v = Elem<T>::null();
        x[i] = v;
    }
}

extern "C" int rfb_fill_splitmix_dev(rfb_ctx_t *ctx, int type, void *x, int64_t n, uint64_t seed, uint64_t modulus,
                                     int64_t offset, int64_t null_every, double f64_scale) {
    RFB_ARG(ctx && n >= 0 && (x || n == 0), "rfb_fill_splitmix_dev");
    if (n == 0) return RFB_OK;
    const int grid = rfb_grid_for(ctx, n, 256, 8);
    if (f64_scale == 0.0) f64_scale = 1.0;
    switch (rfb_kind_of(type)) {
        case K_I32: k_fill_splitmix<i32><<<grid, 256, 0, ctx->stream>>>((i32 *)x, n, seed, modulus, offset, null_every, f64_scale); break;
        case K_I64: k_fill_splitmix<i64><<<grid, 256, 0, ctx->stream>>>((i64 *)x, n, seed, modulus, offset, null_every, f64_scale); break;
        case K_F64: k_fill_splitmix<f64><<<grid, 256, 0, ctx->stream>>>((f64 *)x, n, seed, modulus, offset, null_every, f64_scale); break;
        default: rfb_set_error("fill: unsupported type %d", type); return RFB_ERR_TYPE;
    }
    RFB_CHECK_LAUNCH(ctx);
    return RFB_OK;
}
