// Internal declarations shared by the CUDA translation units of libmemci_b200.so.
// Not part of the ABI (see include/memci_b200.h for that).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/memci_b200.h"

#define MEMCI_MAX_PYR 12
#define MEMCI_L8_PAD 128             // zero border (pixels / rows) around the 8-bit luma plane of the full-search prefilter

struct MVField {
    int valid;
    int mv_cell, mv_rows, mv_cols;     // (dy,dx) grid: value of pixel (r,c) = vec[(r/mv_cell)*mv_cols + c/mv_cell]
    int sad_cell, sad_rows, sad_cols;  // SAD grid (stays on ME blocks after smoothing)
    float2 *vec;  size_t vec_cap;      // elements
    float *sad;   size_t sad_cap;
};

struct FrameSlot {
    int valid;
    uint8_t *rgb;                      // [H][W][3]
    int32_t *luma;                     // [H][Wp] luma*1000 = 299R+587G+114B (exact integer)
    int luma_ok;
    uint8_t *luma8;                    // round(luma) as uint8: four planes [H + 2 PAD][pitch8], plane q shifted left by q bytes, zero border (fs8.cu); allocated on first use
    int luma8_ok;
    // HBMA pyramid cache (levels 1..pyr_levels), invalidated on upload
    int pyr_levels;
    uint8_t *pyr_rgb[MEMCI_MAX_PYR + 1];
    int32_t *pyr_luma[MEMCI_MAX_PYR + 1];
    size_t pyr_cap[MEMCI_MAX_PYR + 1];
};

struct FsTile { int by, bx0, nb, pad; };

struct FsPlan {                        // cached per (bs, R): tile table + tensor maps
    int bs, R, G, box_w, box_h, n_tiles, nb_max, valid;
    FsTile *d_tiles;
    long long n_cand;
    int *row_tile_start;               // host: first tile of each block row (nbr + 1 entries)
    long long *row_cand;               // host: candidates per block row
};

struct memci_ctx {
    int device, H, W, Wp;
    cudaStream_t stream;
    int own_stream;
    int num_sms;
    FrameSlot frames[MEMCI_NUM_FRAME_SLOTS];
    MVField mvs[MEMCI_NUM_MV_SLOTS];
    MVField tmp_mv[2];                 // HBMA ping-pong
    // full-search scratch
    int *d_redo_list; int redo_cap; int *d_counters;   // [0] redo count, [1] hole count, [2] dmax, [4] hole runs, [5] bidir2 max |T|, [6] bidir2 out-of-range blocks, [8] fs8 list overflow, [9] fs8 survivors, [10] fs8 float64 replays, [11] redo blocks of the fs8 path
    FsPlan fs_plan[4];
    unsigned *d_fs8_list; size_t fs8_list_cap;          // prefilter survivors (bytes)
    unsigned long long *d_fs8_best; size_t fs8_best_cap; // per (direction, block) minimum key of the refine pass (bytes)
    unsigned char *d_fs8_codes; size_t fs8_codes_cap, fs8_bmin_off;   // prefilter: candidate codes, column minima, block minima (bytes)
    long long fs_last_cand, fs_last_ops;
    int fs_used_prefilter;             // the last full search took the two-phase path (fs8.cu)
    // content feedback for the two-phase path: counters [8..11] of a recent search, copied back asynchronously
    int *h_fs8_feedback; cudaEvent_t ev_fs8_feedback; int fs8_feedback_pending; long long fs8_feedback_cand;
    int fs8_attr_set;                  // dynamic shared-memory attribute of the prefilter kernels set on this context's device
    void *d_f8_tiles; int f8_tiles_n; int f8_tiles_key[8];   // job table of the prefilter (fs8.cu), cached per search geometry
    int fs8_backoff;                   // > 0: that many searches skip the prefilter (the last feedback said it does not pay on this content)
    // MCI scratch
    short2 *d_disp[2]; size_t disp_cap[2];
    uint32_t *d_fill; int *d_hole_list; size_t hole_cap;
    float *d_dense; size_t dense_cap;  // dense MV staging for download/upload
    void *d_quality; size_t quality_cap;     // PSNR / SSIM partial sums
    int4 *d_b2blk[2]; size_t b2blk_cap[2];   // bidir2: per 4x4 block (Xb, Yb, TXb, TYb << 1 | occluded), per direction
    // timing
    int timing; cudaEvent_t ev[4]; int fs_timed, mci_timed;
    // event pool: per-launch device times of the two hot kernels inside a caller's timed region
    cudaEvent_t *pool; int pool_cap, pool_n; unsigned char *pool_kind; double *pool_work;   // pairs (2i, 2i+1); kinds: MEMCI_PROF_* (memci_b200.h); work = algorithmic px-ops / bytes of the launch
    // asynchronous drain of output frames through a stager's D2H stream: slot may be rewritten only after its copy
    cudaEvent_t ev_out_ready[MEMCI_NUM_FRAME_SLOTS], ev_out_copied[MEMCI_NUM_FRAME_SLOTS];
    unsigned char out_pending[MEMCI_NUM_FRAME_SLOTS];
    long long launches;
    int sticky;                        // sticky CUDA error
    char err[512];
};

// ---- error handling --------------------------------------------------------------
int memci_fail(memci_ctx *ctx, int code, const char *fmt, ...);
#define CK(ctx, call)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            if (ctx) (ctx)->sticky = 1;                                                 \
            return memci_fail(ctx, MEMCI_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__,  \
                              #call, cudaGetErrorString(e__));                          \
        }                                                                               \
    } while (0)
#define CKL(ctx) do { (ctx)->launches++; CK(ctx, cudaGetLastError()); } while (0)

int memci_reserve(memci_ctx *ctx, void **ptr, size_t *cap, size_t need_bytes);   // grow-only device buffer
int memci_mv_reserve(memci_ctx *ctx, MVField *f, int mv_cell, int mv_rows, int mv_cols,
                     int sad_cell, int sad_rows, int sad_cols);

// ---- stage entry points (host side, enqueue on ctx->stream) -------------------------
int memci_luma_plane(memci_ctx *ctx, const uint8_t *rgb, int H, int W, int Wp, int32_t *luma);
int memci_ensure_luma(memci_ctx *ctx, int slot);
int memci_ensure_pyramid(memci_ctx *ctx, int slot, int levels);
int memci_fs_run(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int R, MVField *out, int br0, int br1,
                 MVField *out2 = nullptr);   // block rows [br0, br1); br1 < 0 = all; out2: reverse direction in the same launch
int memci_tmap_encode(memci_ctx *ctx, CUtensorMap *tm, int dtype_u8, int rank, const void *base,
                      const unsigned long long *dims, const unsigned long long *strides, const unsigned *box);   // fs.cu
int memci_fs8_eligible(memci_ctx *ctx, int bs, int R);
int memci_fs8_run(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int R, MVField *out, int br0, int br1, MVField *out2);
int memci_tss_run(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int steps, MVField *out);
int memci_hbma_run(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int win, int sub, int steps,
                   int min_bs, int cost_key, MVField *out);
int memci_smooth_run(memci_ctx *ctx, const MVField *in, int mode, int fsz, MVField *out);
int memci_mci_run(memci_ctx *ctx, int slot_src, int slot_tgt, const MVField *fwd, const MVField *bwd,
                  float dist, int out_slot, int bidir, int row0, int row1);   // output rows [row0, row1); row1 < 0 = all
int memci_bidir2_run(memci_ctx *ctx, int slot_src, int slot_tgt, const MVField *fwd, const MVField *bwd,
                     float dist_fwd, float dist_bwd, int mbs, int pad, const float *w64, int out_slot);
int memci_quality_run(memci_ctx *ctx, int slot_a, int slot_b, double *h_result);   // h_result[0] = sum of squared differences, [1] = sum of SSIM map
int memci_mv_to_dense(memci_ctx *ctx, const MVField *f, float *d_dense);
// returns an event pair index (>= 0) after recording the start event, or -1 when profiling is off / full
int memci_prof_begin(memci_ctx *ctx, int kind, double work = 0.0);
void memci_prof_end(memci_ctx *ctx, int idx);
int memci_dense_to_mv(memci_ctx *ctx, const float *d_dense, MVField *f);

static inline int ceil_div_i(int a, int b) { return (a + b - 1) / b; }

// ---- device helpers ----------------------------------------------------------------
#ifdef __CUDACC__

// luma*1000 as an exact integer (<= 255000, 18 bits).  SAD_ref = trunc(sum|dL|) relates to
// S1000 = sum |dL1000| by  SAD_ref = S1000 / 1000  unless S1000 % 1000 == 0.
__device__ __forceinline__ int luma1000(int r, int g, int b) { return 299 * r + 587 * g + 114 * b; }

// Luma exactly as the reference's numba code evaluates it (f64, separate mul/add, N1):
// ME/full_search.py:15-16, ME/hbma.py:20-21.
__device__ __forceinline__ double luma_f64(int r, int g, int b)
{
    double t0 = __dmul_rn(0.299, (double)r);
    double t1 = __dmul_rn(0.587, (double)g);
    double t2 = __dmul_rn(0.114, (double)b);
    return __dadd_rn(__dadd_rn(t0, t1), t2);
}

// Exact replay of get_sad / get_cost in the reference's operation order: sequential
// row-major f64 accumulation from 0.0.  Pixels outside the frame read as 0 (np.pad).
// cost_key 0 = SAD (full_search.py:17 / hbma.py:22), 1 = SSD (hbma.py:28).
__device__ inline double exact_cost_f64(const uint8_t *__restrict__ A, const uint8_t *__restrict__ B,
                                        int H, int W, int ar, int ac, int br, int bc, int bs, int cost_key)
{
    double acc = 0.0;
    for (int i = 0; i < bs; ++i) {
        for (int j = 0; j < bs; ++j) {
            int r1 = ar + i, c1 = ac + j, r2 = br + i, c2 = bc + j;
            double l1 = 0.0, l2 = 0.0;
            if (r1 < H && c1 < W) {
                const uint8_t *p = A + ((size_t)r1 * W + c1) * 3;
                l1 = luma_f64(p[0], p[1], p[2]);
            }
            if (r2 < H && c2 < W) {
                const uint8_t *p = B + ((size_t)r2 * W + c2) * 3;
                l2 = luma_f64(p[0], p[1], p[2]);
            }
            double d = __dsub_rn(l1, l2);
            acc = cost_key == 0 ? __dadd_rn(acc, fabs(d)) : __dadd_rn(acc, __dmul_rn(d, d));
        }
    }
    return acc;
}

// Integer SAD of luma1000 planes over a bs x bs block with zero padding outside the frame.
__device__ inline unsigned int sad1000_global(const int32_t *__restrict__ LA, const int32_t *__restrict__ LB,
                                              int H, int W, int Wp, int ar, int ac, int br, int bc, int bs)
{
    unsigned int s = 0;
    for (int i = 0; i < bs; ++i) {
        int r1 = ar + i, r2 = br + i;
        for (int j = 0; j < bs; ++j) {
            int c1 = ac + j, c2 = bc + j;
            int a = (r1 < H && c1 < W) ? LA[(size_t)r1 * Wp + c1] : 0;
            int b = (r2 < H && c2 < W) ? LB[(size_t)r2 * Wp + c2] : 0;
            s = __sad(a, b, s);
        }
    }
    return s;
}

__device__ __forceinline__ unsigned long long shfl_min_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}


// ---- mbarrier / TMA (cp.async.bulk.tensor) primitives shared by fs.cu and fs8.cu ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Same, for waits that may take a while: after a failed probe the warp sleeps instead of spinning through the issue slots
// the working warps need.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra MBARS_DONE;\n"
        "MBARS_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra MBARS_DONE;\n"
        "bra MBARS_WAIT;\n"
        "MBARS_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");      // the thread is suspended for up to ~20 us per try
}
// Non-blocking probe: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int x, int y, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int x, int y, int z, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

#endif  // __CUDACC__
