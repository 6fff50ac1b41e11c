// C ABI of libmemci_b200.so (see include/memci_b200.h): context, slots, copies, dispatch.
#include <stdarg.h>
#include <stdlib.h>

#include "memci_internal.cuh"

static char g_err[512] = "";

int memci_fail(memci_ctx *ctx, int code, const char *fmt, ...)
{
    char *dst = ctx ? ctx->err : g_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

int memci_reserve(memci_ctx *ctx, void **ptr, size_t *cap, size_t need)
{
    if (*cap >= need && *ptr) return MEMCI_OK;
    if (*ptr) { CK(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(*ptr); *ptr = nullptr; *cap = 0; }
    size_t sz = need < 256 ? 256 : need;
    CK(ctx, cudaMalloc(ptr, sz));
    *cap = sz;
    return MEMCI_OK;
}

int memci_mv_reserve(memci_ctx *ctx, MVField *f, int mv_cell, int mv_rows, int mv_cols,
                     int sad_cell, int sad_rows, int sad_cols)
{
    size_t nv = (size_t)mv_rows * mv_cols, ns = (size_t)sad_rows * sad_cols;
    size_t capv = f->vec_cap * sizeof(float2), caps = f->sad_cap * sizeof(float);
    int rc;
    if ((rc = memci_reserve(ctx, (void **)&f->vec, &capv, nv * sizeof(float2)))) return rc;
    if ((rc = memci_reserve(ctx, (void **)&f->sad, &caps, ns * sizeof(float)))) return rc;
    f->vec_cap = capv / sizeof(float2);
    f->sad_cap = caps / sizeof(float);
    f->mv_cell = mv_cell; f->mv_rows = mv_rows; f->mv_cols = mv_cols;
    f->sad_cell = sad_cell; f->sad_rows = sad_rows; f->sad_cols = sad_cols;
    return MEMCI_OK;
}

#define REQ_CTX(ctx) do { if (!(ctx)) return memci_fail(nullptr, MEMCI_ERR_ARG, "null context"); \
                          if ((ctx)->sticky) return MEMCI_ERR_CUDA; \
                          cudaSetDevice((ctx)->device); } while (0)
#define REQ_FRAME(ctx, s, need_valid) do { \
        if ((s) < 0 || (s) >= MEMCI_NUM_FRAME_SLOTS) return memci_fail(ctx, MEMCI_ERR_ARG, "frame slot %d out of range", (s)); \
        if ((need_valid) && !(ctx)->frames[s].valid) return memci_fail(ctx, MEMCI_ERR_STATE, "frame slot %d is empty", (s)); } while (0)
#define REQ_MV(ctx, s, need_valid) do { \
        if ((s) < 0 || (s) >= MEMCI_NUM_MV_SLOTS) return memci_fail(ctx, MEMCI_ERR_ARG, "mv slot %d out of range", (s)); \
        if ((need_valid) && !(ctx)->mvs[s].valid) return memci_fail(ctx, MEMCI_ERR_STATE, "mv slot %d is empty", (s)); } while (0)

int memci_prof_begin(memci_ctx *ctx, int kind, double work)
{
    if (!ctx->pool || ctx->pool_n >= ctx->pool_cap) return -1;
    int i = ctx->pool_n++;
    ctx->pool_kind[i] = (unsigned char)kind;
    ctx->pool_work[i] = work;
    cudaEventRecord(ctx->pool[2 * i], ctx->stream);
    return i;
}

void memci_prof_end(memci_ctx *ctx, int idx)
{
    if (idx >= 0) cudaEventRecord(ctx->pool[2 * idx + 1], ctx->stream);
}

extern "C" {

int memci_version(void) { return 100; }

int memci_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { memci_fail(nullptr, MEMCI_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); return MEMCI_ERR_CUDA; }
    return n;
}

int memci_ctx_create(int device, int H, int W, void *stream, memci_ctx **out)
{
    if (!out) return memci_fail(nullptr, MEMCI_ERR_ARG, "null out pointer");
    *out = nullptr;
    if (H <= 0 || W <= 0 || H > 8192 || W > 8192) return memci_fail(nullptr, MEMCI_ERR_ARG, "frame size %dx%d out of range", H, W);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return memci_fail(nullptr, MEMCI_ERR_CUDA, "cudaSetDevice(%d): %s (no CPU fallback exists)", device, cudaGetErrorString(e));
    memci_ctx *ctx = (memci_ctx *)calloc(1, sizeof(memci_ctx));
    if (!ctx) return memci_fail(nullptr, MEMCI_ERR_NOMEM, "out of host memory");
    ctx->device = device; ctx->H = H; ctx->W = W; ctx->Wp = (W + 3) & ~3;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { free(ctx); return memci_fail(nullptr, MEMCI_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    if (prop.major != 10) { free(ctx); return memci_fail(nullptr, MEMCI_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); }
    ctx->num_sms = prop.multiProcessorCount;
    if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = 0; }
    else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { free(ctx); return memci_fail(nullptr, MEMCI_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
        ctx->own_stream = 1;
    }
    size_t fb = (size_t)H * W * 3, lb = (size_t)H * ctx->Wp * 4;
    for (int i = 0; i < MEMCI_NUM_FRAME_SLOTS; ++i) {
        if (cudaMalloc(&ctx->frames[i].rgb, fb + 16) != cudaSuccess   /* +16: the splat kernel reads pixels as aligned word pairs */ || cudaMalloc(&ctx->frames[i].luma, lb) != cudaSuccess) {
            memci_fail(nullptr, MEMCI_ERR_NOMEM, "cudaMalloc of frame slots failed");
            memci_ctx_destroy(ctx);
            return MEMCI_ERR_NOMEM;
        }
    }
    if (cudaMalloc(&ctx->d_counters, 16 * sizeof(int)) != cudaSuccess ||
        cudaMemset(ctx->d_counters, 0, 16 * sizeof(int)) != cudaSuccess) {
        memci_fail(nullptr, MEMCI_ERR_NOMEM, "cudaMalloc of counters failed");
        memci_ctx_destroy(ctx);
        return MEMCI_ERR_NOMEM;
    }
    for (int i = 0; i < 4; ++i) cudaEventCreate(&ctx->ev[i]);
    *out = ctx;
    return MEMCI_OK;
}

int memci_ctx_destroy(memci_ctx *ctx)
{
    if (!ctx) return MEMCI_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < MEMCI_NUM_FRAME_SLOTS; ++i) {
        FrameSlot &f = ctx->frames[i];
        if (f.rgb) cudaFree(f.rgb);
        if (f.luma) cudaFree(f.luma);
        if (f.luma8) cudaFree(f.luma8);
        for (int l = 0; l <= MEMCI_MAX_PYR; ++l) {
            if (f.pyr_rgb[l]) cudaFree(f.pyr_rgb[l]);
            if (f.pyr_luma[l]) cudaFree(f.pyr_luma[l]);
        }
    }
    for (int i = 0; i < MEMCI_NUM_MV_SLOTS; ++i) {
        if (ctx->mvs[i].vec) cudaFree(ctx->mvs[i].vec);
        if (ctx->mvs[i].sad) cudaFree(ctx->mvs[i].sad);
    }
    for (int i = 0; i < 2; ++i) {
        if (ctx->tmp_mv[i].vec) cudaFree(ctx->tmp_mv[i].vec);
        if (ctx->tmp_mv[i].sad) cudaFree(ctx->tmp_mv[i].sad);
        if (ctx->d_disp[i]) cudaFree(ctx->d_disp[i]);
        if (ctx->d_b2blk[i]) cudaFree(ctx->d_b2blk[i]);
    }
    if (ctx->d_quality) cudaFree(ctx->d_quality);
    if (ctx->h_fs8_feedback) { cudaFreeHost(ctx->h_fs8_feedback); cudaEventDestroy(ctx->ev_fs8_feedback); }
    if (ctx->d_fs8_list) cudaFree(ctx->d_fs8_list);
    if (ctx->d_f8_tiles) cudaFree(ctx->d_f8_tiles);
    if (ctx->d_fs8_best) cudaFree(ctx->d_fs8_best);
    if (ctx->d_fs8_codes) cudaFree(ctx->d_fs8_codes);
    for (int i = 0; i < 4; ++i) {
        if (ctx->fs_plan[i].d_tiles) cudaFree(ctx->fs_plan[i].d_tiles);
        free(ctx->fs_plan[i].row_tile_start); free(ctx->fs_plan[i].row_cand);
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    }
    for (int i = 0; i < MEMCI_NUM_FRAME_SLOTS; ++i) {
        if (ctx->ev_out_ready[i]) cudaEventDestroy(ctx->ev_out_ready[i]);
        if (ctx->ev_out_copied[i]) cudaEventDestroy(ctx->ev_out_copied[i]);
    }
    if (ctx->d_redo_list) cudaFree(ctx->d_redo_list);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_fill) cudaFree(ctx->d_fill);
    if (ctx->d_hole_list) cudaFree(ctx->d_hole_list);
    if (ctx->d_dense) cudaFree(ctx->d_dense);
    for (int i = 0; i < 2 * ctx->pool_cap; ++i) cudaEventDestroy(ctx->pool[i]);
    free(ctx->pool); free(ctx->pool_kind); free(ctx->pool_work);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    free(ctx);
    return MEMCI_OK;
}

const char *memci_last_error(memci_ctx *ctx) { return ctx ? ctx->err : g_err; }

int memci_sync(memci_ctx *ctx)
{
    REQ_CTX(ctx);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return MEMCI_OK;
}

void *memci_stream(memci_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int memci_host_alloc(size_t bytes, void **out)
{
    if (!out) return MEMCI_ERR_ARG;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) return memci_fail(nullptr, MEMCI_ERR_CUDA, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    return MEMCI_OK;
}

int memci_host_free(void *p)
{
    if (p && cudaFreeHost(p) != cudaSuccess) return MEMCI_ERR_CUDA;
    return MEMCI_OK;
}

// ---- frames ----------------------------------------------------------------------------
static void frame_touched(FrameSlot &f) { f.valid = 1; f.luma_ok = 0; f.luma8_ok = 0; f.pyr_levels = 0; }

// output slot still being copied out by a stager -> the context's stream waits for that copy
static int memci_wait_slot_free(memci_ctx *ctx, int slot)
{
    if (slot >= 0 && slot < MEMCI_NUM_FRAME_SLOTS && ctx->out_pending[slot]) {
        CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out_copied[slot], 0));
        ctx->out_pending[slot] = 0;
    }
    return MEMCI_OK;
}

int memci_upload_frame(memci_ctx *ctx, int slot, const uint8_t *host)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot, 0);
    if (!host) return memci_fail(ctx, MEMCI_ERR_ARG, "null host frame");
    { int rc_ = memci_wait_slot_free(ctx, slot); if (rc_) return rc_; }
    CK(ctx, cudaMemcpyAsync(ctx->frames[slot].rgb, host, (size_t)ctx->H * ctx->W * 3, cudaMemcpyHostToDevice, ctx->stream));
    frame_touched(ctx->frames[slot]);
    return MEMCI_OK;
}

int memci_set_frame_device(memci_ctx *ctx, int slot, const void *dev)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot, 0);
    if (!dev) return memci_fail(ctx, MEMCI_ERR_ARG, "null device frame");
    { int rc_ = memci_wait_slot_free(ctx, slot); if (rc_) return rc_; }
    CK(ctx, cudaMemcpyAsync(ctx->frames[slot].rgb, dev, (size_t)ctx->H * ctx->W * 3, cudaMemcpyDeviceToDevice, ctx->stream));
    frame_touched(ctx->frames[slot]);
    return MEMCI_OK;
}

int memci_download_frame(memci_ctx *ctx, int slot, uint8_t *host)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot, 1);
    if (!host) return memci_fail(ctx, MEMCI_ERR_ARG, "null host frame");
    CK(ctx, cudaMemcpyAsync(host, ctx->frames[slot].rgb, (size_t)ctx->H * ctx->W * 3, cudaMemcpyDeviceToHost, ctx->stream));
    return MEMCI_OK;
}

// ---- frame staging: host frames cross PCIe once, on copy streams of their own -------------------------
// A stager owns a ring of device frame buffers, an H2D stream and a D2H stream.  Source frames are uploaded
// once (memci_stager_upload) and handed to any number of contexts with a device-to-device copy that waits
// on the upload's event (memci_frame_from_stager); output frames leave through the D2H stream
// (memci_download_frame_via) while the context's own stream goes on with the next pair.  A ring slot is
// overwritten only after every context that took the frame has copied it, an output slot only after its
// D2H copy has finished (memci_wait_slot_free, called by the MCI entry points).
#define STAGER_MAX_CONSUMERS 4
struct memci_stager {
    int device, H, W, n;
    cudaStream_t h2d, d2h;
    uint8_t **buf;
    cudaEvent_t *ready;
    cudaEvent_t *consumed;      // [n][STAGER_MAX_CONSUMERS]
    int *n_consumed;
    unsigned char *has_data;
};

int memci_stager_create(int device, int H, int W, int n_slots, memci_stager **out)
{
    if (!out || H <= 0 || W <= 0 || n_slots <= 0) return MEMCI_ERR_ARG;
    *out = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return memci_fail(nullptr, MEMCI_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    memci_stager *st = (memci_stager *)calloc(1, sizeof(memci_stager));
    if (!st) return MEMCI_ERR_NOMEM;
    st->device = device; st->H = H; st->W = W; st->n = n_slots;
    st->buf = (uint8_t **)calloc(n_slots, sizeof(uint8_t *));
    st->ready = (cudaEvent_t *)calloc(n_slots, sizeof(cudaEvent_t));
    st->consumed = (cudaEvent_t *)calloc((size_t)n_slots * STAGER_MAX_CONSUMERS, sizeof(cudaEvent_t));
    st->n_consumed = (int *)calloc(n_slots, sizeof(int));
    st->has_data = (unsigned char *)calloc(n_slots, 1);
    bool ok = st->buf && st->ready && st->consumed && st->n_consumed && st->has_data &&
              cudaStreamCreateWithFlags(&st->h2d, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&st->d2h, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < n_slots; ++i) {
        ok = cudaMalloc(&st->buf[i], (size_t)H * W * 3 + 16) == cudaSuccess &&
             cudaEventCreateWithFlags(&st->ready[i], cudaEventDisableTiming) == cudaSuccess;
        for (int c = 0; ok && c < STAGER_MAX_CONSUMERS; ++c)
            ok = cudaEventCreateWithFlags(&st->consumed[i * STAGER_MAX_CONSUMERS + c], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) { memci_stager_destroy(st); return memci_fail(nullptr, MEMCI_ERR_NOMEM, "stager allocation failed"); }
    *out = st;
    return MEMCI_OK;
}

int memci_stager_destroy(memci_stager *st)
{
    if (!st) return MEMCI_OK;
    cudaSetDevice(st->device);
    if (st->h2d) { cudaStreamSynchronize(st->h2d); cudaStreamDestroy(st->h2d); }
    if (st->d2h) { cudaStreamSynchronize(st->d2h); cudaStreamDestroy(st->d2h); }
    for (int i = 0; i < st->n; ++i) {
        if (st->buf && st->buf[i]) cudaFree(st->buf[i]);
        if (st->ready && st->ready[i]) cudaEventDestroy(st->ready[i]);
        for (int c = 0; st->consumed && c < STAGER_MAX_CONSUMERS; ++c)
            if (st->consumed[i * STAGER_MAX_CONSUMERS + c]) cudaEventDestroy(st->consumed[i * STAGER_MAX_CONSUMERS + c]);
    }
    free(st->buf); free(st->ready); free(st->consumed); free(st->n_consumed); free(st->has_data); free(st);
    return MEMCI_OK;
}

int memci_stager_upload(memci_stager *st, int slot, const uint8_t *host)
{
    if (!st || !host || slot < 0 || slot >= st->n) return MEMCI_ERR_ARG;
    cudaSetDevice(st->device);
    // the previous tenant of this ring slot must have been copied out by everyone who asked for it
    for (int c = 0; c < st->n_consumed[slot]; ++c)
        if (cudaStreamWaitEvent(st->h2d, st->consumed[slot * STAGER_MAX_CONSUMERS + c], 0) != cudaSuccess) return MEMCI_ERR_CUDA;
    st->n_consumed[slot] = 0;
    if (cudaMemcpyAsync(st->buf[slot], host, (size_t)st->H * st->W * 3, cudaMemcpyHostToDevice, st->h2d) != cudaSuccess) return MEMCI_ERR_CUDA;
    if (cudaEventRecord(st->ready[slot], st->h2d) != cudaSuccess) return MEMCI_ERR_CUDA;
    st->has_data[slot] = 1;
    return MEMCI_OK;
}

int memci_frame_from_stager(memci_ctx *ctx, int frame_slot, memci_stager *st, int slot)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, frame_slot, 0);
    if (!st || slot < 0 || slot >= st->n) return memci_fail(ctx, MEMCI_ERR_ARG, "bad stager slot %d", slot);
    if (st->H != ctx->H || st->W != ctx->W || st->device != ctx->device) return memci_fail(ctx, MEMCI_ERR_ARG, "stager and context differ in device or frame size");
    if (st->n_consumed[slot] >= STAGER_MAX_CONSUMERS) return memci_fail(ctx, MEMCI_ERR_STATE, "stager slot %d handed out more than %d times", slot, STAGER_MAX_CONSUMERS);
    if (!st->has_data[slot]) return memci_fail(ctx, MEMCI_ERR_STATE, "stager slot %d was never uploaded", slot);
    { int rc_ = memci_wait_slot_free(ctx, frame_slot); if (rc_) return rc_; }
    CK(ctx, cudaStreamWaitEvent(ctx->stream, st->ready[slot], 0));
    CK(ctx, cudaMemcpyAsync(ctx->frames[frame_slot].rgb, st->buf[slot], (size_t)ctx->H * ctx->W * 3, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(ctx, cudaEventRecord(st->consumed[slot * STAGER_MAX_CONSUMERS + st->n_consumed[slot]++], ctx->stream));
    frame_touched(ctx->frames[frame_slot]);
    return MEMCI_OK;
}

int memci_download_frame_via(memci_ctx *ctx, int slot, uint8_t *host, memci_stager *st)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot, 1);
    if (!host || !st) return memci_fail(ctx, MEMCI_ERR_ARG, "null host frame / stager");
    if (!ctx->ev_out_ready[slot]) {
        CK(ctx, cudaEventCreateWithFlags(&ctx->ev_out_ready[slot], cudaEventDisableTiming));
        CK(ctx, cudaEventCreateWithFlags(&ctx->ev_out_copied[slot], cudaEventDisableTiming));
    }
    CK(ctx, cudaEventRecord(ctx->ev_out_ready[slot], ctx->stream));
    CK(ctx, cudaStreamWaitEvent(st->d2h, ctx->ev_out_ready[slot], 0));
    CK(ctx, cudaMemcpyAsync(host, ctx->frames[slot].rgb, (size_t)ctx->H * ctx->W * 3, cudaMemcpyDeviceToHost, st->d2h));
    CK(ctx, cudaEventRecord(ctx->ev_out_copied[slot], st->d2h));
    ctx->out_pending[slot] = 1;
    return MEMCI_OK;
}

int memci_stager_sync(memci_stager *st)
{
    if (!st) return MEMCI_ERR_ARG;
    cudaSetDevice(st->device);
    if (cudaStreamSynchronize(st->h2d) != cudaSuccess || cudaStreamSynchronize(st->d2h) != cudaSuccess) return MEMCI_ERR_CUDA;
    return MEMCI_OK;
}

void *memci_frame_device_ptr(memci_ctx *ctx, int slot)
{
    if (!ctx || slot < 0 || slot >= MEMCI_NUM_FRAME_SLOTS) return nullptr;
    return ctx->frames[slot].rgb;
}

// ---- ME ----------------------------------------------------------------------------------
int memci_me_fs(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int R, int mv_slot)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot_src, 1); REQ_FRAME(ctx, slot_tgt, 1); REQ_MV(ctx, mv_slot, 0);
    return memci_fs_run(ctx, slot_src, slot_tgt, bs, R, &ctx->mvs[mv_slot], 0, -1);
}

int memci_me_fs_pair(memci_ctx *ctx, int slot_a, int slot_b, int bs, int R, int mv_fwd, int mv_bwd)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot_a, 1); REQ_FRAME(ctx, slot_b, 1); REQ_MV(ctx, mv_fwd, 0); REQ_MV(ctx, mv_bwd, 0);
    if (mv_fwd == mv_bwd) return memci_fail(ctx, MEMCI_ERR_ARG, "pair search needs two distinct MV slots");
    return memci_fs_run(ctx, slot_a, slot_b, bs, R, &ctx->mvs[mv_fwd], 0, -1, &ctx->mvs[mv_bwd]);
}

int memci_me_fs_rows(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int R, int mv_slot, int block_row0, int block_row1)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot_src, 1); REQ_FRAME(ctx, slot_tgt, 1); REQ_MV(ctx, mv_slot, 0);
    return memci_fs_run(ctx, slot_src, slot_tgt, bs, R, &ctx->mvs[mv_slot], block_row0, block_row1);
}

int memci_me_tss(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int steps, int mv_slot)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot_src, 1); REQ_FRAME(ctx, slot_tgt, 1); REQ_MV(ctx, mv_slot, 0);
    return memci_tss_run(ctx, slot_src, slot_tgt, bs, steps, &ctx->mvs[mv_slot]);
}

int memci_me_hbma(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int win, int sub, int steps,
                  int min_bs, int cost_key, int mv_slot)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot_src, 1); REQ_FRAME(ctx, slot_tgt, 1); REQ_MV(ctx, mv_slot, 0);
    return memci_hbma_run(ctx, slot_src, slot_tgt, bs, win, sub, steps, min_bs, cost_key, &ctx->mvs[mv_slot]);
}

int memci_download_pyramid_level(memci_ctx *ctx, int slot, int level, uint8_t *host)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot, 1);
    if (level < 1 || level > MEMCI_MAX_PYR || !host) return memci_fail(ctx, MEMCI_ERR_ARG, "bad pyramid level %d", level);
    int rc = memci_ensure_pyramid(ctx, slot, level);
    if (rc) return rc;
    int h = ctx->H, w = ctx->W;
    for (int l = 1; l <= level; ++l) { h = (h + 1) / 2; w = (w + 1) / 2; }
    CK(ctx, cudaMemcpyAsync(host, ctx->frames[slot].pyr_rgb[level], (size_t)h * w * 3, cudaMemcpyDeviceToHost, ctx->stream));
    return MEMCI_OK;
}

int memci_smooth(memci_ctx *ctx, int mv_in, int mode, int fsz, int mv_out)
{
    REQ_CTX(ctx); REQ_MV(ctx, mv_in, 1); REQ_MV(ctx, mv_out, 0);
    if (mode < MEMCI_FILTER_NONE || mode > MEMCI_FILTER_WEIGHTED) return memci_fail(ctx, MEMCI_ERR_ARG, "unknown filter mode %d", mode);
    MVField *in = &ctx->mvs[mv_in], *out = &ctx->mvs[mv_out];
    if (mode == MEMCI_FILTER_NONE) {                    // smooth(None, ...) returns the field unchanged (:15-16)
        if (in == out) return MEMCI_OK;
        int rc = memci_mv_reserve(ctx, out, in->mv_cell, in->mv_rows, in->mv_cols, in->sad_cell, in->sad_rows, in->sad_cols);
        if (rc) return rc;
        CK(ctx, cudaMemcpyAsync(out->vec, in->vec, sizeof(float2) * (size_t)in->mv_rows * in->mv_cols, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(ctx, cudaMemcpyAsync(out->sad, in->sad, sizeof(float) * (size_t)in->sad_rows * in->sad_cols, cudaMemcpyDeviceToDevice, ctx->stream));
        out->valid = 1;
        return MEMCI_OK;
    }
    MVField *tmp = &ctx->tmp_mv[0];
    int rc = memci_smooth_run(ctx, in, mode, fsz, tmp);
    if (rc) return rc;
    MVField t = *out; *out = *tmp; *tmp = t;            // hand the buffers over (stream-ordered reuse)
    tmp->valid = 0;
    return MEMCI_OK;
}

int memci_mv_info(memci_ctx *ctx, int s, int *mv_cell, int *mv_rows, int *mv_cols, int *sad_cell, int *sad_rows, int *sad_cols)
{
    REQ_CTX(ctx); REQ_MV(ctx, s, 1);
    MVField &f = ctx->mvs[s];
    if (mv_cell) *mv_cell = f.mv_cell;
    if (mv_rows) *mv_rows = f.mv_rows;
    if (mv_cols) *mv_cols = f.mv_cols;
    if (sad_cell) *sad_cell = f.sad_cell;
    if (sad_rows) *sad_rows = f.sad_rows;
    if (sad_cols) *sad_cols = f.sad_cols;
    return MEMCI_OK;
}

int memci_download_mv_blocks(memci_ctx *ctx, int s, float *vec, float *sad)
{
    REQ_CTX(ctx); REQ_MV(ctx, s, 1);
    MVField &f = ctx->mvs[s];
    if (vec) CK(ctx, cudaMemcpyAsync(vec, f.vec, sizeof(float2) * (size_t)f.mv_rows * f.mv_cols, cudaMemcpyDeviceToHost, ctx->stream));
    if (sad) CK(ctx, cudaMemcpyAsync(sad, f.sad, sizeof(float) * (size_t)f.sad_rows * f.sad_cols, cudaMemcpyDeviceToHost, ctx->stream));
    return MEMCI_OK;
}

int memci_download_mv_dense(memci_ctx *ctx, int s, float *host)
{
    REQ_CTX(ctx); REQ_MV(ctx, s, 1);
    if (!host) return memci_fail(ctx, MEMCI_ERR_ARG, "null host buffer");
    size_t bytes = (size_t)ctx->H * ctx->W * 3 * sizeof(float), cap = ctx->dense_cap;
    int rc = memci_reserve(ctx, (void **)&ctx->d_dense, &cap, bytes);
    ctx->dense_cap = cap;
    if (rc) return rc;
    if ((rc = memci_mv_to_dense(ctx, &ctx->mvs[s], ctx->d_dense))) return rc;
    CK(ctx, cudaMemcpyAsync(host, ctx->d_dense, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return MEMCI_OK;
}

int memci_upload_mv_dense(memci_ctx *ctx, int s, const float *host)
{
    REQ_CTX(ctx); REQ_MV(ctx, s, 0);
    if (!host) return memci_fail(ctx, MEMCI_ERR_ARG, "null host buffer");
    size_t bytes = (size_t)ctx->H * ctx->W * 3 * sizeof(float), cap = ctx->dense_cap;
    int rc = memci_reserve(ctx, (void **)&ctx->d_dense, &cap, bytes);
    ctx->dense_cap = cap;
    if (rc) return rc;
    CK(ctx, cudaMemcpyAsync(ctx->d_dense, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return memci_dense_to_mv(ctx, ctx->d_dense, &ctx->mvs[s]);
}

int memci_upload_mv_blocks(memci_ctx *ctx, int s, int mv_cell, const float *vec, int sad_cell, const float *sad)
{
    REQ_CTX(ctx); REQ_MV(ctx, s, 0);
    if (!vec || !sad || mv_cell <= 0 || sad_cell <= 0) return memci_fail(ctx, MEMCI_ERR_ARG, "bad block-level MV upload arguments");
    MVField &f = ctx->mvs[s];
    int mr = ceil_div_i(ctx->H, mv_cell), mc = ceil_div_i(ctx->W, mv_cell);
    int sr = ceil_div_i(ctx->H, sad_cell), sc = ceil_div_i(ctx->W, sad_cell);
    int rc = memci_mv_reserve(ctx, &f, mv_cell, mr, mc, sad_cell, sr, sc);
    if (rc) return rc;
    CK(ctx, cudaMemcpyAsync(f.vec, vec, sizeof(float2) * (size_t)mr * mc, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(f.sad, sad, sizeof(float) * (size_t)sr * sc, cudaMemcpyHostToDevice, ctx->stream));
    f.valid = 1;
    return MEMCI_OK;
}

// ---- MCI ---------------------------------------------------------------------------------
int memci_mci_unidir(memci_ctx *ctx, int slot_src, int slot_tgt, int mv_slot, float dist, int out_slot)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot_src, 1); REQ_FRAME(ctx, slot_tgt, 1); REQ_FRAME(ctx, out_slot, 0); REQ_MV(ctx, mv_slot, 1);
    if (out_slot == slot_src || out_slot == slot_tgt) return memci_fail(ctx, MEMCI_ERR_ARG, "output slot aliases an input slot");
    { int rc_ = memci_wait_slot_free(ctx, out_slot); if (rc_) return rc_; }
    return memci_mci_run(ctx, slot_src, slot_tgt, &ctx->mvs[mv_slot], nullptr, dist, out_slot, 0, 0, -1);
}

int memci_mci_bidir(memci_ctx *ctx, int slot_src, int slot_tgt, int mv_fwd, int mv_bwd, float dist, int out_slot)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot_src, 1); REQ_FRAME(ctx, slot_tgt, 1); REQ_FRAME(ctx, out_slot, 0);
    REQ_MV(ctx, mv_fwd, 1); REQ_MV(ctx, mv_bwd, 1);
    if (out_slot == slot_src || out_slot == slot_tgt) return memci_fail(ctx, MEMCI_ERR_ARG, "output slot aliases an input slot");
    { int rc_ = memci_wait_slot_free(ctx, out_slot); if (rc_) return rc_; }
    return memci_mci_run(ctx, slot_src, slot_tgt, &ctx->mvs[mv_fwd], &ctx->mvs[mv_bwd], dist, out_slot, 1, 0, -1);
}

int memci_mci_bidir2(memci_ctx *ctx, int slot_src, int slot_tgt, int mv_fwd, int mv_bwd, float dist_fwd, float dist_bwd,
                     int min_block_size, int pad_size, const float *weights, int out_slot)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot_src, 1); REQ_FRAME(ctx, slot_tgt, 1); REQ_FRAME(ctx, out_slot, 0);
    REQ_MV(ctx, mv_fwd, 1); REQ_MV(ctx, mv_bwd, 1);
    if (!weights) return memci_fail(ctx, MEMCI_ERR_ARG, "null weight kernel");
    if (out_slot == slot_src || out_slot == slot_tgt) return memci_fail(ctx, MEMCI_ERR_ARG, "output slot aliases an input slot");
    { int rc_ = memci_wait_slot_free(ctx, out_slot); if (rc_) return rc_; }
    return memci_bidir2_run(ctx, slot_src, slot_tgt, &ctx->mvs[mv_fwd], &ctx->mvs[mv_bwd], dist_fwd, dist_bwd,
                            min_block_size, pad_size, weights, out_slot);
}

int memci_bidir2_stats(memci_ctx *ctx, int *max_shift, int *n_out_of_range)
{
    REQ_CTX(ctx);
    int h[2] = {0, 0};
    CK(ctx, cudaMemcpyAsync(h, ctx->d_counters + 5, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (max_shift) *max_shift = h[0];
    if (n_out_of_range) *n_out_of_range = h[1];
    return MEMCI_OK;
}

int memci_bidir2_out_of_range_total(memci_ctx *ctx, int *n_out_of_range, int reset)
{
    REQ_CTX(ctx);
    int h = 0;
    CK(ctx, cudaMemcpyAsync(&h, ctx->d_counters + 15, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (reset) CK(ctx, cudaMemsetAsync(ctx->d_counters + 15, 0, sizeof(int), ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_out_of_range) *n_out_of_range = h;
    return MEMCI_OK;
}

int memci_quality(memci_ctx *ctx, int slot_a, int slot_b, double *mse, double *ssim)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot_a, 1); REQ_FRAME(ctx, slot_b, 1);
    const int H = ctx->H, W = ctx->W;
    if (H < 7 || W < 7) return memci_fail(ctx, MEMCI_ERR_ARG, "win_size exceeds image extent (%d x %d frame, 7 x 7 window)", H, W);
    double h[2] = {0.0, 0.0};
    int rc = memci_quality_run(ctx, slot_a, slot_b, h);
    if (rc) return rc;
    if (mse) *mse = h[0] / ((double)H * (double)W * 3.0);
    if (ssim) *ssim = h[1] / ((double)(H - 6) * (double)(W - 6) * 3.0);
    return MEMCI_OK;
}

int memci_mci_rows(memci_ctx *ctx, int slot_src, int slot_tgt, int mv_fwd, int mv_bwd, float dist, int out_slot,
                   int bidir, int row0, int row1)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot_src, 1); REQ_FRAME(ctx, slot_tgt, 1); REQ_FRAME(ctx, out_slot, 0);
    REQ_MV(ctx, mv_fwd, 1);
    if (bidir) REQ_MV(ctx, mv_bwd, 1);
    if (out_slot == slot_src || out_slot == slot_tgt) return memci_fail(ctx, MEMCI_ERR_ARG, "output slot aliases an input slot");
    { int rc_ = memci_wait_slot_free(ctx, out_slot); if (rc_) return rc_; }
    return memci_mci_run(ctx, slot_src, slot_tgt, &ctx->mvs[mv_fwd], bidir ? &ctx->mvs[mv_bwd] : nullptr, dist, out_slot,
                         bidir ? 1 : 0, row0, row1);
}

// ---- spatial (band) tiling support: rows of frames / MV fields written from outside ------------
int memci_frame_mark(memci_ctx *ctx, int slot)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot, 0);
    frame_touched(ctx->frames[slot]);
    return MEMCI_OK;
}

int memci_upload_frame_rows(memci_ctx *ctx, int slot, const uint8_t *host_hwc, int row0, int row1)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot, 0);
    if (!host_hwc || row0 < 0 || row1 > ctx->H || row0 >= row1) return memci_fail(ctx, MEMCI_ERR_ARG, "bad row range [%d, %d)", row0, row1);
    const size_t pitch = (size_t)ctx->W * 3;
    CK(ctx, cudaMemcpyAsync(ctx->frames[slot].rgb + row0 * pitch, host_hwc + row0 * pitch, (size_t)(row1 - row0) * pitch,
                            cudaMemcpyHostToDevice, ctx->stream));
    frame_touched(ctx->frames[slot]);
    return MEMCI_OK;
}

int memci_download_frame_rows(memci_ctx *ctx, int slot, uint8_t *host_hwc, int row0, int row1)
{
    REQ_CTX(ctx); REQ_FRAME(ctx, slot, 1);
    if (!host_hwc || row0 < 0 || row1 > ctx->H || row0 >= row1) return memci_fail(ctx, MEMCI_ERR_ARG, "bad row range [%d, %d)", row0, row1);
    const size_t pitch = (size_t)ctx->W * 3;
    CK(ctx, cudaMemcpyAsync(host_hwc + row0 * pitch, ctx->frames[slot].rgb + row0 * pitch, (size_t)(row1 - row0) * pitch,
                            cudaMemcpyDeviceToHost, ctx->stream));
    return MEMCI_OK;
}

int memci_mv_alloc(memci_ctx *ctx, int mv_slot, int cell)
{
    REQ_CTX(ctx); REQ_MV(ctx, mv_slot, 0);
    if (cell <= 0) return memci_fail(ctx, MEMCI_ERR_ARG, "bad cell size %d", cell);
    const int r = ceil_div_i(ctx->H, cell), c = ceil_div_i(ctx->W, cell);
    int rc = memci_mv_reserve(ctx, &ctx->mvs[mv_slot], cell, r, c, cell, r, c);
    if (rc) return rc;
    ctx->mvs[mv_slot].valid = 1;
    return MEMCI_OK;
}

int memci_mv_device_ptrs(memci_ctx *ctx, int mv_slot, void **vec, void **sad)
{
    REQ_CTX(ctx); REQ_MV(ctx, mv_slot, 1);
    if (vec) *vec = ctx->mvs[mv_slot].vec;
    if (sad) *sad = ctx->mvs[mv_slot].sad;
    return MEMCI_OK;
}

int memci_estimate_pair(memci_ctx *ctx, const memci_pair_params *p, int slot_src, int slot_tgt, int mv_fwd, int mv_bwd)
{
    REQ_CTX(ctx);
    if (!p) return memci_fail(ctx, MEMCI_ERR_ARG, "null params");
    int rc;
    const bool fused = p->me_mode == 0 && p->bidir && mv_fwd != mv_bwd;
    if (fused && (rc = memci_me_fs_pair(ctx, slot_src, slot_tgt, p->block_size, p->region, mv_fwd, mv_bwd))) return rc;
    for (int d = 0; d < (p->bidir ? 2 : 1); ++d) {
        int a = d ? slot_tgt : slot_src, b = d ? slot_src : slot_tgt, mv = d ? mv_bwd : mv_fwd;
        if (fused) rc = MEMCI_OK;
        else if (p->me_mode == 0) rc = memci_me_fs(ctx, a, b, p->block_size, p->region, mv);
        else if (p->me_mode == 1) rc = memci_me_hbma(ctx, a, b, p->block_size, p->region, p->sub_region, p->steps,
                                                     p->min_block_size, MEMCI_COST_SAD, mv);
        else if (p->me_mode == 2) rc = memci_me_tss(ctx, a, b, p->block_size, p->region, mv);
        else return memci_fail(ctx, MEMCI_ERR_ARG, "unknown me_mode %d", p->me_mode);
        if (rc) return rc;
        if (p->filter_mode != MEMCI_FILTER_NONE && (rc = memci_smooth(ctx, mv, p->filter_mode, p->filter_size, mv))) return rc;
    }
    return MEMCI_OK;
}

// ---- instrumentation ---------------------------------------------------------------------
int memci_profile_start(memci_ctx *ctx, int capacity)
{
    REQ_CTX(ctx);
    if (capacity <= 0) return memci_fail(ctx, MEMCI_ERR_ARG, "profile capacity must be positive");
    if (ctx->pool_cap < capacity) {
        for (int i = 0; i < 2 * ctx->pool_cap; ++i) cudaEventDestroy(ctx->pool[i]);
        free(ctx->pool); free(ctx->pool_kind); free(ctx->pool_work);
        ctx->pool = (cudaEvent_t *)calloc(2 * (size_t)capacity, sizeof(cudaEvent_t));
        ctx->pool_kind = (unsigned char *)calloc(capacity, 1);
        ctx->pool_work = (double *)calloc(capacity, sizeof(double));
        if (!ctx->pool || !ctx->pool_kind || !ctx->pool_work) return memci_fail(ctx, MEMCI_ERR_NOMEM, "out of host memory");
        for (int i = 0; i < 2 * capacity; ++i) CK(ctx, cudaEventCreate(&ctx->pool[i]));
        ctx->pool_cap = capacity;
    }
    ctx->pool_n = 0;
    return MEMCI_OK;
}

int memci_profile_collect(memci_ctx *ctx, double *fs_ms_sum, int *fs_n, double *mci_ms_sum, int *mci_n)
{
    REQ_CTX(ctx);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    double s[2] = {0.0, 0.0};
    int n[2] = {0, 0};
    for (int i = 0; i < ctx->pool_n; ++i) {
        float ms = 0.f;
        CK(ctx, cudaEventElapsedTime(&ms, ctx->pool[2 * i], ctx->pool[2 * i + 1]));
        if (ctx->pool_kind[i] < 2) { s[ctx->pool_kind[i]] += ms; n[ctx->pool_kind[i]]++; }
    }
    if (fs_ms_sum) *fs_ms_sum = s[0];
    if (fs_n) *fs_n = n[0];
    if (mci_ms_sum) *mci_ms_sum = s[1];
    if (mci_n) *mci_n = n[1];
    ctx->pool_n = ctx->pool_cap;          // stop recording until the next memci_profile_start
    return MEMCI_OK;
}

int memci_profile_collect_kinds(memci_ctx *ctx, double *ms_sum, int *launches, double *work)
{
    REQ_CTX(ctx);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < MEMCI_PROF_KINDS; ++k) { if (ms_sum) ms_sum[k] = 0.0; if (launches) launches[k] = 0; if (work) work[k] = 0.0; }
    for (int i = 0; i < ctx->pool_n; ++i) {
        float ms = 0.f;
        CK(ctx, cudaEventElapsedTime(&ms, ctx->pool[2 * i], ctx->pool[2 * i + 1]));
        const int k = ctx->pool_kind[i];
        if (k >= MEMCI_PROF_KINDS) continue;
        if (ms_sum) ms_sum[k] += ms;
        if (launches) launches[k]++;
        if (work) work[k] += ctx->pool_work[i];
    }
    ctx->pool_n = ctx->pool_cap;          // stop recording until the next memci_profile_start
    return MEMCI_OK;
}

long long memci_launch_count(memci_ctx *ctx) { return ctx ? ctx->launches : -1; }
int memci_reset_launch_count(memci_ctx *ctx) { if (!ctx) return MEMCI_ERR_ARG; ctx->launches = 0; return MEMCI_OK; }

int memci_fs_stats(memci_ctx *ctx, long long *n_cand, long long *px_ops, int *n_redo)
{
    REQ_CTX(ctx);
    if (n_cand) *n_cand = ctx->fs_last_cand;
    if (px_ops) *px_ops = ctx->fs_last_ops;
    if (n_redo) {                                       // float64 replays: redo-list blocks (exact kernel) + flagged survivors (prefilter path)
        int h[12];
        CK(ctx, cudaMemcpyAsync(h, ctx->d_counters, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        *n_redo = ctx->fs_used_prefilter ? h[10] + h[11] : h[0];     // float64 replays of flagged survivors + redo blocks
    }
    return MEMCI_OK;
}

int memci_fs_prefilter_stats(memci_ctx *ctx, int *n_survivors, int *overflowed)
{
    REQ_CTX(ctx);
    int h[2] = {0, 0};
    if (ctx->fs_used_prefilter) {
        CK(ctx, cudaMemcpyAsync(h, ctx->d_counters + 8, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (overflowed) *overflowed = h[0];
    if (n_survivors) *n_survivors = h[1];
    return MEMCI_OK;
}

int memci_mci_stats(memci_ctx *ctx, int *n_holes)
{
    REQ_CTX(ctx);
    if (n_holes) {
        CK(ctx, cudaMemcpyAsync(n_holes, ctx->d_counters + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MEMCI_OK;
}

int memci_enable_kernel_timing(memci_ctx *ctx, int on) { REQ_CTX(ctx); ctx->timing = on ? 1 : 0; return MEMCI_OK; }

int memci_last_kernel_ms(memci_ctx *ctx, float *fs_ms, float *mci_ms)
{
    REQ_CTX(ctx);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (fs_ms) { *fs_ms = -1.f; if (ctx->fs_timed) CK(ctx, cudaEventElapsedTime(fs_ms, ctx->ev[0], ctx->ev[1])); }
    if (mci_ms) { *mci_ms = -1.f; if (ctx->mci_timed) CK(ctx, cudaEventElapsedTime(mci_ms, ctx->ev[2], ctx->ev[3])); }
    return MEMCI_OK;
}

}  // extern "C"
