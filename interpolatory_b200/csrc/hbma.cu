// Hierarchical block matching (ME/hbma.py) on sm_100a.
//
// HBMA evaluates ~27 candidates per block per level (3 candidate vectors x 3x3 refinement),
// two orders of magnitude less SAD work than full search, so the kernels here are the simple
// warp-per-block form: lanes = candidates, luma1000 planes read through L1/L2.
//
// Cost model (get_cost, hbma.py:15-31): float32( sum_f64 |L1-L2| ), compared as float32.
// (float)((double)S1000 / 1000.0) reproduces it unless the value sits within the f64
// accumulation error of a float32 rounding boundary; those candidates, and every
// SSD candidate, are replayed in float64 in the reference's operation order.
#include "memci_internal.cuh"

struct BwRes { float r, c, cost; };

__device__ __forceinline__ float hbma_cost(const int32_t *__restrict__ L1, const int32_t *__restrict__ L2,
                                           const uint8_t *__restrict__ rgb1, const uint8_t *__restrict__ rgb2,
                                           int H, int W, int Wp, int ar, int ac, int br, int bc, int bs,
                                           int cost_key, double eps)
{
    if (cost_key == 0) {
        unsigned S = sad1000_global(L1, L2, H, W, Wp, ar, ac, br, bc, bs);
        double v = (double)S / 1000.0;
        float lo = (float)(v - eps), hi = (float)(v + eps);
        if (lo == hi) return (float)v;
    }
    return (float)exact_cost_f64(rgb1, rgb2, H, W, ar, ac, br, bc, bs, cost_key);
}

// block_wise_fs (hbma.py:34-75), executed by one warp.  (ar, ac) = top-left of block1 in
// frame 1, (idx_r, idx_c) = search centre in frame 2.  All lanes return the result.
__device__ inline BwRes bwfs_warp(const int32_t *__restrict__ L1, const int32_t *__restrict__ L2,
                                  const uint8_t *__restrict__ rgb1, const uint8_t *__restrict__ rgb2,
                                  int H, int W, int Wp, int ar, int ac, int idx_r, int idx_c, int win, int bs,
                                  int cost_key, double eps, int lane)
{
    BwRes res;
    if (idx_r >= H || idx_c >= W || idx_r < 0 || idx_c < 0) {           // :36-37
        res.r = 0.f; res.c = 0.f; res.cost = __int_as_float(0x7f800000);
        return res;
    }
    int max_r = win + 1, min_r = -win, max_c = win + 1, min_c = -win;
    if (idx_r + win + bs >= H) max_r = max(1, H - idx_r - bs);           // :53-54
    if (idx_r - win < 0) min_r += win - idx_r;                           // :55-56
    if (idx_c + win + bs >= W) max_c = max(1, W - idx_c - bs);
    if (idx_c - win < 0) min_c += win - idx_c;
    const int nr = max_r - min_r, nc = max_c - min_c, n = nr * nc;
    unsigned long long best = 0xffffffffffffffffull;
    for (int c = lane; c < n; c += 32) {
        int ri = c / nc, ci = c - ri * nc;
        int r_off = min_r + ri, c_off = min_c + ci;
        float cost = hbma_cost(L1, L2, rgb1, rgb2, H, W, Wp, ar, ac, idx_r + r_off, idx_c + c_off, bs, cost_key, eps);
        unsigned d2 = (unsigned)(r_off * r_off + c_off * c_off);
        // cost >= 0: the float32 bit pattern orders like the value; ties on the f32 value fall
        // through to distance, then (r_off, c_off) scan order  (:69)
        unsigned long long k = ((unsigned long long)__float_as_uint(cost) << 32) |
                               ((unsigned long long)d2 << 16) | ((unsigned)ri << 8) | (unsigned)ci;
        best = k < best ? k : best;
    }
    best = shfl_min_u64(best);
    res.r = (float)(min_r + (int)((best >> 8) & 255));
    res.c = (float)(min_c + (int)(best & 255));
    res.cost = __uint_as_float((unsigned)(best >> 32));
    return res;
}

// full_search_jit (hbma.py:77-90): block-level field on the coarsest level.
__global__ void hbma_fs_kernel(const int32_t *__restrict__ L1, const int32_t *__restrict__ L2,
                               const uint8_t *__restrict__ rgb1, const uint8_t *__restrict__ rgb2,
                               int H, int W, int Wp, int bs, int win, int cost_key, double eps,
                               int rows, int cols, float2 *__restrict__ vec, float *__restrict__ cost)
{
    int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int blk = warp; blk < rows * cols; blk += nwarps) {
        int row = blk / cols, col = blk - row * cols;
        BwRes r = bwfs_warp(L1, L2, rgb1, rgb2, H, W, Wp, row * bs, col * bs, row * bs, col * bs, win, bs,
                            cost_key, eps, lane);
        if (lane == 0) { vec[blk] = make_float2(r.r, r.c); cost[blk] = r.cost; }
    }
}

// increase_vec_density_jit (hbma.py:100-140): one warp per child block of the sliced output.
__global__ void hbma_densify_kernel(const int32_t *__restrict__ L1, const int32_t *__restrict__ L2,
                                    const uint8_t *__restrict__ rgb1, const uint8_t *__restrict__ rgb2,
                                    int H, int W, int Wp, int bs, int sub_win, int cost_key, double eps, int vec_scale,
                                    const float2 *__restrict__ in_vec, int in_rows, int in_cols, int in_pitch,
                                    float2 *__restrict__ out_vec, float *__restrict__ out_cost,
                                    int out_rows, int out_cols, int out_pitch)
{
    int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int blk = warp; blk < out_rows * out_cols; blk += nwarps) {
        int o_r = blk / out_cols, o_c = blk - o_r * out_cols;
        int row = o_r >> 1, col = o_c >> 1;
        float2 o = make_float2(0.f, 0.f);
        float oc = 0.f;
        bool evaluated = row < in_rows && col < in_cols;
        if (evaluated) {
            // second child row/col is skipped when the parent touches the frame edge (:117-124, Q5)
            int r_max = ((row + 1) * 2 * bs >= H) ? row * 2 + 1 : (row + 1) * 2;
            int c_max = ((col + 1) * 2 * bs >= W) ? col * 2 + 1 : (col + 1) * 2;
            evaluated = o_r < r_max && o_c < c_max;
        }
        if (evaluated) {
            float2 v[3];
            v[0] = in_vec[(size_t)row * in_pitch + col];                          // parent, unscaled (:107, Q4)
            int hc = col >= 1 ? col - 1 : min(col + 1, in_cols - 1);
            int vr = row >= 1 ? row - 1 : min(row + 1, in_rows - 1);
            float2 h = in_vec[(size_t)row * in_pitch + hc];
            float2 u = in_vec[(size_t)vr * in_pitch + col];
            float sc = (float)vec_scale;
            v[1] = make_float2(__fmul_rn(h.x, sc), __fmul_rn(h.y, sc));
            v[2] = make_float2(__fmul_rn(u.x, sc), __fmul_rn(u.y, sc));
            int im_r = o_r * bs, im_c = o_c * bs;
            float lowest = __int_as_float(0x7f800000);
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {
                int ir = (int)((double)im_r + (double)v[k].x);                    // f64 sum, trunc to int32 (:134)
                int ic = (int)((double)im_c + (double)v[k].y);
                BwRes r = bwfs_warp(L1, L2, rgb1, rgb2, H, W, Wp, im_r, im_c, ir, ic, sub_win, bs, cost_key, eps, lane);
                if (r.cost < lowest) {                                            // strict: first vector wins ties (:135)
                    lowest = r.cost;
                    oc = r.cost;
                    o = make_float2(__fadd_rn(r.r, v[k].x), __fadd_rn(r.c, v[k].y));
                }
            }
        }
        if (lane == 0) {
            out_vec[(size_t)o_r * out_pitch + o_c] = o;
            out_cost[(size_t)o_r * out_pitch + o_c] = oc;
        }
    }
}

__global__ void hbma_compact_kernel(const float2 *__restrict__ in_vec, const float *__restrict__ in_cost, int pitch,
                                    int rows, int cols, float2 *__restrict__ out_vec, float *__restrict__ out_sad)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    int r = i / cols, c = i - r * cols;
    out_vec[i] = in_vec[(size_t)r * pitch + c];
    out_sad[i] = in_cost[(size_t)r * pitch + c];
}

static double hbma_eps(int bs)
{
    // bound on |f64 sequential sum - exact sum| for bs*bs terms of magnitude <= 255 (x2 margin)
    double n = (double)bs * bs;
    return 2.0 * n * (2e-13 + 2.9e-14 * n);
}

int memci_hbma_run(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int win, int sub, int steps,
                   int min_bs, int cost_key, MVField *out)
{
    if (bs <= 0 || bs > 64 || win < 0 || win > 127 || sub < 0 || sub > 127 || steps < 0 || steps > MEMCI_MAX_PYR ||
        min_bs <= 0 || (cost_key != 0 && cost_key != 1))
        return memci_fail(ctx, MEMCI_ERR_ARG, "hbma: argument out of range (bs %d win %d sub %d steps %d min_bs %d cost %d)",
                          bs, win, sub, steps, min_bs, cost_key);
    int rc;
    if ((rc = memci_ensure_luma(ctx, slot_src))) return rc;
    if ((rc = memci_ensure_luma(ctx, slot_tgt))) return rc;
    if ((rc = memci_ensure_pyramid(ctx, slot_src, steps))) return rc;
    if ((rc = memci_ensure_pyramid(ctx, slot_tgt, steps))) return rc;
    FrameSlot &f1 = ctx->frames[slot_src], &f2 = ctx->frames[slot_tgt];
    int hs[MEMCI_MAX_PYR + 1], ws[MEMCI_MAX_PYR + 1];
    hs[0] = ctx->H; ws[0] = ctx->W;
    for (int l = 1; l <= steps; ++l) { hs[l] = (hs[l - 1] + 1) / 2; ws[l] = (ws[l - 1] + 1) / 2; }
    auto LUMA = [&](FrameSlot &f, int l) { return l == 0 ? f.luma : f.pyr_luma[l]; };
    auto RGB = [&](FrameSlot &f, int l) { return l == 0 ? f.rgb : f.pyr_rgb[l]; };
    auto WP = [&](int l) { return l == 0 ? ctx->Wp : ((ws[l] + 3) & ~3); };

    // worst-case intermediate size: the last densification output at the final block size
    int fbs = bs;
    while (fbs > min_bs) fbs >>= 1;
    if (fbs <= 0) return memci_fail(ctx, MEMCI_ERR_ARG, "hbma: min_block_size %d unreachable from %d", min_bs, bs);
    size_t max_cells = 0;
    {
        int r = ceil_div_i(hs[steps], bs), c = ceil_div_i(ws[steps], bs);
        max_cells = (size_t)r * c;
        for (int l = steps - 1; l >= 0; --l) {
            size_t n = (size_t)(2 * r) * (2 * c);
            if (n > max_cells) max_cells = n;
            r = ceil_div_i(hs[l], bs); c = ceil_div_i(ws[l], bs);
        }
        int b = bs;
        while (b > min_bs) {
            b >>= 1;
            size_t n = (size_t)(2 * r) * (2 * c);
            if (n > max_cells) max_cells = n;
            r = ceil_div_i(ctx->H, b); c = ceil_div_i(ctx->W, b);
        }
    }
    for (int i = 0; i < 2; ++i)
        if ((rc = memci_mv_reserve(ctx, &ctx->tmp_mv[i], 1, 1, (int)max_cells, 1, 1, (int)max_cells))) return rc;

    const int grid = ctx->num_sms * 8, nt = 256;
    int cur = 0;
    int rows = ceil_div_i(hs[steps], bs), cols = ceil_div_i(ws[steps], bs), pitch = cols;
    hbma_fs_kernel<<<grid, nt, 0, ctx->stream>>>(LUMA(f1, steps), LUMA(f2, steps), RGB(f1, steps), RGB(f2, steps),
                                                 hs[steps], ws[steps], WP(steps), bs, win, cost_key, hbma_eps(bs),
                                                 rows, cols, ctx->tmp_mv[cur].vec, ctx->tmp_mv[cur].sad);
    CKL(ctx);
    auto densify = [&](int lvl, int b, int vec_scale) -> int {
        int orows = ceil_div_i(hs[lvl], b), ocols = ceil_div_i(ws[lvl], b);   // the slice kept by the reference
        int opitch = cols * 2;
        hbma_densify_kernel<<<grid, nt, 0, ctx->stream>>>(LUMA(f1, lvl), LUMA(f2, lvl), RGB(f1, lvl), RGB(f2, lvl),
                                                          hs[lvl], ws[lvl], WP(lvl), b, sub, cost_key, hbma_eps(b), vec_scale,
                                                          ctx->tmp_mv[cur].vec, rows, cols, pitch,
                                                          ctx->tmp_mv[cur ^ 1].vec, ctx->tmp_mv[cur ^ 1].sad,
                                                          orows, ocols, opitch);
        CKL(ctx);
        cur ^= 1; rows = orows; cols = ocols; pitch = opitch;
        return MEMCI_OK;
    };
    for (int lvl = steps - 1; lvl >= 0; --lvl)
        if ((rc = densify(lvl, bs, 2))) return rc;                               // hbma.py:181-182
    int b = bs;
    while (b > min_bs) {                                                         // hbma.py:184-186
        b >>= 1;
        if ((rc = densify(0, b, 1))) return rc;
    }
    if ((rc = memci_mv_reserve(ctx, out, b, rows, cols, b, rows, cols))) return rc;
    hbma_compact_kernel<<<ceil_div_i(rows * cols, 256), 256, 0, ctx->stream>>>(ctx->tmp_mv[cur].vec, ctx->tmp_mv[cur].sad,
                                                                               pitch, rows, cols, out->vec, out->sad);
    CKL(ctx);
    out->valid = 1;
    return MEMCI_OK;
}
