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
#include "rfb_common.cuh"

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
            RFB_CUDA(cudaMemcpyAsync(ctx->d_stage[0][b], (const char *)val + r0 * vsz, (size_t)rows * vsz,
                                     cudaMemcpyHostToDevice, ctx->copy_stream));
            copied += rows * vsz;
            if (ncols == 2) {
                RFB_CUDA(cudaMemcpyAsync(ctx->d_stage[1][b], (const char *)pred + r0 * psz, (size_t)rows * psz,
                                         cudaMemcpyHostToDevice, ctx->copy_stream));
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
