// rfb_host.cu — host layer of the C ABI: HOST column payloads in, host results out (include/rfb200.h, "host layer").
//
// This is the path the reference-facing operator layer and bench.py's `e2e` number use: the column is shipped to HBM
// once, in chunks, with cudaMemcpyAsync on a copy stream through a ring of RFB_STAGE_BUFS device staging buffers; the
// scan+fold kernel of chunk i runs on the compute stream while chunk i+1 is in flight over PCIe.  Each chunk's kernel
// reports into its own slot of the mapped pinned result area; the slots are combined on the host in chunk order
// (integer sums wrap, min/max skip empty chunks, fp64 partial sums are added in a fixed order => deterministic).
//
// Pinned (cudaHostAlloc / rfb_host_pin) payloads are DMA'd directly; pageable payloads still work (the driver stages
// them) but at a fraction of the PCIe rate — INTEGRATION.md tells the reference side to pin its column blocks.
#include <fcntl.h>
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "rfb_common.cuh"

// ------------------------------------------------------------------ copies that suit the memory they are given
//
// Measured on the B200 host (profiles/r01_h2d_probe.txt): DMA from pinned memory 55.6 GB/s; cudaMemcpy from pageable
// memory 11 GB/s; cudaHostRegister itself only 12 GB/s (so registering per query buys nothing); 8 copier threads filling
// a ring of pinned 16 MiB buffers that the DMA engine drains: 31 GB/s.  The reference's columns live in pageable heap
// blocks, so that last route is what its integration gets unless it pins its blocks once at allocation.

namespace {

struct CopyPool {
    int nthreads;
    pthread_t th[16];
    pthread_mutex_t mu;
    pthread_cond_t cv_work, cv_done;
    const char *src;
    char *dst;
    size_t bytes;
    unsigned long gen;
    int pending;
    bool stop;
};

struct WorkerArg { CopyPool *p; int id; };

void *copy_worker(void *a) {
    WorkerArg *wa = (WorkerArg *)a;
    CopyPool *p = wa->p;
    const int id = wa->id;
    free(wa);
    // A host that pins its calling thread to one core (the reference pins every executor, core/pool.c:168-219) hands that
    // one-core affinity to the threads it creates: undo it, or all copiers would share a single core.
    {
        cpu_set_t all;
        CPU_ZERO(&all);
        const long cores = sysconf(_SC_NPROCESSORS_CONF);
        for (long c = 0; c < cores && c < CPU_SETSIZE; c++) CPU_SET((int)c, &all);
        pthread_setaffinity_np(pthread_self(), sizeof(all), &all);
    }
    unsigned long seen = 0;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        while (p->gen == seen && !p->stop) pthread_cond_wait(&p->cv_work, &p->mu);
        if (p->stop) { pthread_mutex_unlock(&p->mu); return nullptr; }
        seen = p->gen;
        const char *src = p->src; char *dst = p->dst; const size_t bytes = p->bytes;
        pthread_mutex_unlock(&p->mu);
        const size_t per = (bytes / (size_t)(p->nthreads + 1) + 63) & ~(size_t)63, lo = per * (size_t)(id + 1);
        if (lo < bytes) memcpy(dst + lo, src + lo, (lo + per < bytes) ? per : bytes - lo);
        pthread_mutex_lock(&p->mu);
        if (--p->pending == 0) pthread_cond_signal(&p->cv_done);
        pthread_mutex_unlock(&p->mu);
    }
}

CopyPool *pool_of(rfb_ctx_t *ctx) {
    if (ctx->copy_pool) return (CopyPool *)ctx->copy_pool;
    CopyPool *p = (CopyPool *)calloc(1, sizeof(CopyPool));
    if (!p) return nullptr;
    long cores = sysconf(_SC_NPROCESSORS_ONLN);
    const char *e = getenv("RFB200_COPY_THREADS");
    int want = e ? atoi(e) : 8;
    if (want > cores - 1) want = (int)cores - 1;
    if (want > 16) want = 16;
    if (want < 0) want = 0;
    pthread_mutex_init(&p->mu, nullptr);
    pthread_cond_init(&p->cv_work, nullptr);
    pthread_cond_init(&p->cv_done, nullptr);
    for (int i = 0; i < want; i++) {   // the calling thread copies the first slice itself
        WorkerArg *wa = (WorkerArg *)malloc(sizeof(WorkerArg));
        wa->p = p; wa->id = p->nthreads;
        if (pthread_create(&p->th[p->nthreads], nullptr, copy_worker, wa) == 0) p->nthreads++; else free(wa);
    }
    ctx->copy_pool = p;
    return p;
}

// memcpy split over the pool (the caller takes slice 0)
void parallel_memcpy(CopyPool *p, void *dst, const void *src, size_t bytes) {
    if (!p || p->nthreads == 0 || bytes < (1u << 20)) { memcpy(dst, src, bytes); return; }
    pthread_mutex_lock(&p->mu);
    p->src = (const char *)src; p->dst = (char *)dst; p->bytes = bytes;
    p->pending = p->nthreads;
    p->gen++;
    pthread_cond_broadcast(&p->cv_work);
    pthread_mutex_unlock(&p->mu);
    const size_t per = (bytes / (size_t)(p->nthreads + 1) + 63) & ~(size_t)63;
    memcpy(dst, src, per < bytes ? per : bytes);
    pthread_mutex_lock(&p->mu);
    while (p->pending) pthread_cond_wait(&p->cv_done, &p->mu);
    pthread_mutex_unlock(&p->mu);
}

bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int ensure_ring(rfb_ctx_t *ctx) {
    for (int i = 0; i < RFB_HOST_RING; i++)
        if (!ctx->h_ring[i]) {
            RFB_CUDA(cudaHostAlloc(&ctx->h_ring[i], RFB_HOST_RING_BYTES, cudaHostAllocDefault));
            RFB_CUDA(cudaEventCreateWithFlags(&ctx->ev_ring[i], cudaEventDisableTiming));
        }
    return RFB_OK;
}

constexpr size_t STAGED_MIN = 4u << 20;   // below this the plain call is as good

bool staged_enabled() {   // RFB200_STAGED_COPY=0 falls back to plain cudaMemcpyAsync from/to pageable memory
    static int on = -1;
    if (on < 0) { const char *e = getenv("RFB200_STAGED_COPY"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

}  // namespace

int rfb_copy_h2d(rfb_ctx_t *ctx, void *dst_dev, const void *src_host, size_t bytes, cudaStream_t stream) {
    if (bytes < STAGED_MIN || !staged_enabled() || is_pinned(src_host)) {
        RFB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, stream));
        return RFB_OK;
    }
    int rc = ensure_ring(ctx);
    if (rc) return rc;
    CopyPool *pool = pool_of(ctx);
    for (size_t off = 0; off < bytes; off += RFB_HOST_RING_BYTES) {
        const size_t len = bytes - off < RFB_HOST_RING_BYTES ? bytes - off : RFB_HOST_RING_BYTES;
        const int b = ctx->ring_next;
        ctx->ring_next = (b + 1) % RFB_HOST_RING;
        RFB_CUDA(cudaEventSynchronize(ctx->ev_ring[b]));   // the DMA that last used this pinned buffer is done
        parallel_memcpy(pool, ctx->h_ring[b], (const char *)src_host + off, len);
        RFB_CUDA(cudaMemcpyAsync((char *)dst_dev + off, ctx->h_ring[b], len, cudaMemcpyHostToDevice, stream));
        RFB_CUDA(cudaEventRecord(ctx->ev_ring[b], stream));
    }
    return RFB_OK;
}

// device -> pageable host: DMA into the pinned ring, copier threads move it out.  Returns with the data in place.
int rfb_copy_d2h(rfb_ctx_t *ctx, void *dst_host, const void *src_dev, size_t bytes, cudaStream_t stream) {
    if (bytes < STAGED_MIN || !staged_enabled() || is_pinned(dst_host)) {
        RFB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, stream));
        return RFB_OK;
    }
    int rc = ensure_ring(ctx);
    if (rc) return rc;
    CopyPool *pool = pool_of(ctx);
    const size_t nchunks = (bytes + RFB_HOST_RING_BYTES - 1) / RFB_HOST_RING_BYTES;
    // software pipeline: chunk c's DMA is in flight while chunk c-1 is copied out of its pinned buffer
    for (size_t c = 0; c <= nchunks; c++) {
        if (c < nchunks) {
            const size_t off = c * RFB_HOST_RING_BYTES, len = bytes - off < RFB_HOST_RING_BYTES ? bytes - off : RFB_HOST_RING_BYTES;
            const int b = (int)(c % RFB_HOST_RING);
            if (c >= RFB_HOST_RING - 1) { /* buffer b was drained by the host at step c - RING + 1 <= c - 1: free */ }
            RFB_CUDA(cudaMemcpyAsync(ctx->h_ring[b], (const char *)src_dev + off, len, cudaMemcpyDeviceToHost, stream));
            RFB_CUDA(cudaEventRecord(ctx->ev_ring[b], stream));
        }
        if (c >= 1) {
            const size_t pc = c - 1, off = pc * RFB_HOST_RING_BYTES, len = bytes - off < RFB_HOST_RING_BYTES ? bytes - off : RFB_HOST_RING_BYTES;
            const int b = (int)(pc % RFB_HOST_RING);
            RFB_CUDA(cudaEventSynchronize(ctx->ev_ring[b]));
            parallel_memcpy(pool, (char *)dst_host + off, ctx->h_ring[b], len);
        }
    }
    ctx->ring_next = 0;
    return RFB_OK;
}

void rfb_copy_shutdown(rfb_ctx_t *ctx) {
    CopyPool *p = (CopyPool *)ctx->copy_pool;
    if (p) {
        pthread_mutex_lock(&p->mu);
        p->stop = true;
        pthread_cond_broadcast(&p->cv_work);
        pthread_mutex_unlock(&p->mu);
        for (int i = 0; i < p->nthreads; i++) pthread_join(p->th[i], nullptr);
        pthread_mutex_destroy(&p->mu);
        pthread_cond_destroy(&p->cv_work);
        pthread_cond_destroy(&p->cv_done);
        free(p);
        ctx->copy_pool = nullptr;
    }
    for (int i = 0; i < RFB_HOST_RING; i++)
        if (ctx->h_ring[i]) { cudaFreeHost(ctx->h_ring[i]); cudaEventDestroy(ctx->ev_ring[i]); ctx->h_ring[i] = nullptr; }
}

// ------------------------------------------------------------------ column files (splayed / parted tables on disk)
//
// The reference persists a column as its 16-byte object header (mmod = 0xfd MMOD_EXTERNAL_SIMPLE, type, attrs, len)
// followed by the raw payload (set: core/binary.c:264-307; get mmaps it back: core/unary.c:60-133); a splayed table is a
// directory of such files, a parted table one directory per partition.  The payload of a mapped file is a valid host
// column for every *_host entry point: it is pageable memory, so it travels through the copier-thread ring.


extern "C" int rfb_column_file_open(const char *path, rfb_column_file_t *out) {
    RFB_ARG(path && out, "rfb_column_file_open");
    memset(out, 0, sizeof(*out));
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { rfb_set_error("column file %s: cannot open", path); return RFB_ERR_ARG; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 16) { close(fd); rfb_set_error("column file %s: too short for an object header", path); return RFB_ERR_ARG; }
    void *base = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (base == MAP_FAILED) { rfb_set_error("column file %s: mmap failed", path); return RFB_ERR_NOMEM; }
    const unsigned char *h = (const unsigned char *)base;
    const int type = (signed char)h[2];
    int64_t len;
    memcpy(&len, h + 8, 8);
    const int w = rfb_type_size(type);
    if (h[0] != 0xfd || w == 0 || len < 0 || (uint64_t)len * (uint64_t)w + 16 > (uint64_t)st.st_size) {
        rfb_set_error("column file %s: not a simple numeric column (mmod 0x%02x, type %d, len %lld, %lld bytes)", path, h[0], type, (long long)len, (long long)st.st_size);
        munmap(base, (size_t)st.st_size);
        return RFB_ERR_TYPE;
    }
    out->type = type;
    out->attrs = h[3];
    out->len = len;
    out->payload = h + 16;
    out->map_base = base;
    out->map_bytes = (size_t)st.st_size;
    return RFB_OK;
}

extern "C" int rfb_column_file_close(rfb_column_file_t *f) {
    RFB_ARG(f, "rfb_column_file_close");
    if (f->map_base) munmap(f->map_base, f->map_bytes);
    memset(f, 0, sizeof(*f));
    return RFB_OK;
}

extern "C" int rfb_column_file_write(const char *path, int type, int attrs, const void *payload, int64_t len) {
    RFB_ARG(path && len >= 0 && (payload || len == 0), "rfb_column_file_write");
    const int w = rfb_type_size(type);
    if (!w) { rfb_set_error("column file: unsupported element type %d", type); return RFB_ERR_TYPE; }
    unsigned char h[16];
    memset(h, 0, sizeof(h));
    h[0] = 0xfd;                 // MMOD_EXTERNAL_SIMPLE
    h[2] = (unsigned char)type;  // order 0, rc 0
    h[3] = (unsigned char)attrs; // e.g. the reference's ATTR_ASC / ATTR_DISTINCT flags of a sorted key column
    memcpy(h + 8, &len, 8);
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) { rfb_set_error("column file %s: cannot create", path); return RFB_ERR_ARG; }
    bool ok = write(fd, h, 16) == 16;
    const char *p = (const char *)payload;
    size_t left = (size_t)len * (size_t)w;
    while (ok && left) {
        const ssize_t c = write(fd, p, left > (1u << 30) ? (1u << 30) : left);
        if (c <= 0) ok = false; else { p += c; left -= (size_t)c; }
    }
    close(fd);
    if (!ok) { rfb_set_error("column file %s: short write", path); return RFB_ERR_ARG; }
    return RFB_OK;
}

namespace {

constexpr i64 DEFAULT_CHUNK_BYTES = 64ll << 20;  // 64 MiB per column per chunk: ~1.2 ms of PCIe Gen5 x16

int ensure_stage(rfb_ctx_t *ctx, size_t bytes, int ncols) {
    if (bytes > ctx->stage_bytes) {
        RFB_CUDA(cudaStreamSynchronize(ctx->stream));
        RFB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        for (int c = 0; c < 2; c++)
            for (int i = 0; i < RFB_STAGE_BUFS; i++)
                if (ctx->d_stage[c][i]) { RFB_CUDA(cudaFree(ctx->d_stage[c][i])); ctx->d_stage[c][i] = nullptr; }
        ctx->stage_bytes = 0;
    }
    for (int c = 0; c < ncols; c++)
        for (int i = 0; i < RFB_STAGE_BUFS; i++)
            if (!ctx->d_stage[c][i]) RFB_CUDA(cudaMalloc(&ctx->d_stage[c][i], bytes > ctx->stage_bytes ? bytes : ctx->stage_bytes));
    if (bytes > ctx->stage_bytes) ctx->stage_bytes = bytes;
    return RFB_OK;
}

// fold `part` (one chunk) into `acc`; vkind = element kind of the value column
void combine(rfb_fold_t *acc, const rfb_fold_t *part, int vkind, bool first) {
    if (first) { *acc = *part; return; }
    const bool acc_empty = acc->nonnull == 0, part_empty = part->nonnull == 0;
    acc->rows += part->rows;
    acc->nonnull += part->nonnull;
    if (vkind == K_F64) {
        // error-free merge of the chunk sums: (hi, lo) pairs added with TwoSum, rounded once at the end (finalise_f64)
        const double a = acc->sum_f64, b = part->sum_f64, t = a + b, bp = t - a;
        acc->sum_f64_err = acc->sum_f64_err + part->sum_f64_err + ((a - (t - bp)) + (b - bp));
        acc->sum_f64 = t;
        if (!part_empty) {
            acc->min_f64 = acc_empty ? part->min_f64 : (part->min_f64 < acc->min_f64 ? part->min_f64 : acc->min_f64);
            acc->max_f64 = acc_empty ? part->max_f64 : (part->max_f64 > acc->max_f64 ? part->max_f64 : acc->max_f64);
        }
    } else {
        i64 s = (i64)((u64)acc->sum_i64 + (u64)part->sum_i64);
        if (vkind == K_I32) s = (i64)(i32)(u32)(u64)s;  // I32/TIME sums live in 32 bits (reference core/math.c:1865)
        acc->sum_i64 = s;
        if (!part_empty) {
            acc->min_i64 = acc_empty ? part->min_i64 : (part->min_i64 < acc->min_i64 ? part->min_i64 : acc->min_i64);
            acc->max_i64 = acc_empty ? part->max_i64 : (part->max_i64 > acc->max_i64 ? part->max_i64 : acc->max_i64);
        }
    }
}

// has_pred: filter+fold (pred/k meaningful); otherwise plain fold over val
int pipeline(rfb_ctx_t *ctx, bool has_pred, int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k, int folds,
             int val_type, const void *val, i64 n, i64 chunk_rows, rfb_fold_t *out, i64 *h2d_bytes) {
    const int vsz = rfb_type_size(val_type), psz = has_pred ? rfb_type_size(pred_type) : 0;
    if (!vsz || (has_pred && !psz)) { rfb_set_error("host fold: unsupported element type"); return RFB_ERR_TYPE; }
    const bool same = has_pred && pred == val && pred_type == val_type;
    const int ncols = (has_pred && !same) ? 2 : 1;
    const int maxsz = vsz > psz ? vsz : psz;
    if (chunk_rows <= 0) chunk_rows = DEFAULT_CHUNK_BYTES / maxsz;
    if ((n + chunk_rows - 1) / chunk_rows > RFB_RESULT_SLOTS) chunk_rows = (n + RFB_RESULT_SLOTS - 1) / RFB_RESULT_SLOTS;
    chunk_rows = (chunk_rows + 15) & ~15ll;  // keep every chunk's base 16-byte aligned for the 128-bit loads
    const i64 nchunks = n == 0 ? 1 : (n + chunk_rows - 1) / chunk_rows;
    int rc = ensure_stage(ctx, (size_t)(chunk_rows < n ? chunk_rows : (n ? n : 1)) * maxsz, ncols);
    if (rc) return rc;
    i64 copied = 0;
    const int saved_slot = ctx->result_slot;
    for (i64 c = 0; c < nchunks; c++) {
        const int b = (int)(c % RFB_STAGE_BUFS);
        const i64 r0 = c * chunk_rows, rows = (n - r0) < chunk_rows ? (n - r0) : chunk_rows;
        // the staging slot is free once the kernel that last read it has finished
        if (c >= RFB_STAGE_BUFS) RFB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_kernel[b], 0));
        if (rows > 0) {
            rc = rfb_copy_h2d(ctx, ctx->d_stage[0][b], (const char *)val + r0 * vsz, (size_t)rows * vsz, ctx->copy_stream);
            if (rc) { ctx->result_slot = saved_slot; return rc; }
            copied += rows * vsz;
            if (ncols == 2) {
                rc = rfb_copy_h2d(ctx, ctx->d_stage[1][b], (const char *)pred + r0 * psz, (size_t)rows * psz, ctx->copy_stream);
                if (rc) { ctx->result_slot = saved_slot; return rc; }
                copied += rows * psz;
            }
        }
        RFB_CUDA(cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream));
        RFB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[b], 0));
        ctx->result_slot = (int)c;
        if (has_pred)
            rc = rfb_filter_fold_launch(ctx, cmp_op, pred_type, ncols == 2 ? ctx->d_stage[1][b] : ctx->d_stage[0][b], k, folds,
                                        val_type, ctx->d_stage[0][b], rows);
        else
            rc = rfb_fold_launch(ctx, folds, val_type, ctx->d_stage[0][b], rows);
        if (rc) { ctx->result_slot = saved_slot; return rc; }
        RFB_CUDA(cudaEventRecord(ctx->ev_kernel[b], ctx->stream));
    }
    ctx->result_slot = saved_slot;
    RFB_CUDA(cudaStreamSynchronize(ctx->stream));
    const rfb_fold_t *slots = (const rfb_fold_t *)ctx->h_result;
    const int vk = rfb_kind_of(val_type);
    for (i64 c = 0; c < nchunks; c++) combine(out, &slots[c], vk, c == 0);
    if (vk == K_F64) {  // round the merged pair once
        const double hi = out->sum_f64, lo = out->sum_f64_err, s = hi + lo;
        out->sum_f64_err = (hi - s) + lo;
        out->sum_f64 = s;
    }
    if (h2d_bytes) *h2d_bytes = copied;
    return RFB_OK;
}

}  // namespace

extern "C" int rfb_filter_fold_host(rfb_ctx_t *ctx, int cmp_op, int pred_type, const void *pred, const rfb_scalar_t *k,
                                    int folds, int val_type, const void *val, int64_t n, int64_t chunk_rows,
                                    rfb_fold_t *out, int64_t *h2d_bytes) {
    RFB_ARG(ctx && out && k && n >= 0 && ((pred && val) || n == 0), "rfb_filter_fold_host");
    return pipeline(ctx, true, cmp_op, pred_type, pred, k, folds, val_type, val, n, chunk_rows, out, h2d_bytes);
}

extern "C" int rfb_fold_host(rfb_ctx_t *ctx, int folds, int type, const void *x, int64_t n, int64_t chunk_rows,
                             rfb_fold_t *out, int64_t *h2d_bytes) {
    RFB_ARG(ctx && out && n >= 0 && (x || n == 0), "rfb_fold_host");
    return pipeline(ctx, false, 0, 0, nullptr, nullptr, folds, type, x, n, chunk_rows, out, h2d_bytes);
}
