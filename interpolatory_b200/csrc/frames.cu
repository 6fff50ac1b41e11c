// Frame staging: uint8 RGB -> luma*1000 planes, and the HBMA image pyramid.
#include "memci_internal.cuh"

// ---------------------------------------------------------------------------------------
// luma1000 plane: L[y][x] = 299 R + 587 G + 114 B  (int32, row pitch Wp, pad columns = 0).
// The reference recomputes the luma of both blocks inside every get_sad call
// (ME/full_search.py:13-16); we compute it once per frame.
// ---------------------------------------------------------------------------------------
__global__ void luma_kernel(const uint8_t *__restrict__ rgb, int H, int W, int Wp, int32_t *__restrict__ L)
{
    int quads = Wp >> 2;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * quads) return;
    int y = idx / quads, x = (idx - y * quads) << 2;
    const uint8_t *row = rgb + (size_t)y * W * 3;
    int4 o = make_int4(0, 0, 0, 0);
    if (((W & 3) == 0)) {
        // 4 pixels = 12 bytes = three aligned 32-bit words
        const uint32_t *p = reinterpret_cast<const uint32_t *>(row + (size_t)x * 3);
        uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        o.x = luma1000(w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255);
        o.y = luma1000(w0 >> 24, w1 & 255, (w1 >> 8) & 255);
        o.z = luma1000((w1 >> 16) & 255, w1 >> 24, w2 & 255);
        o.w = luma1000((w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24);
    } else {
        int v[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (x + k < W) {
                const uint8_t *p = row + (size_t)(x + k) * 3;
                v[k] = luma1000(p[0], p[1], p[2]);
            }
        o = make_int4(v[0], v[1], v[2], v[3]);
    }
    *reinterpret_cast<int4 *>(L + (size_t)y * Wp + x) = o;
}

int memci_luma_plane(memci_ctx *ctx, const uint8_t *rgb, int H, int W, int Wp, int32_t *luma)
{
    int n = H * (Wp >> 2);
    luma_kernel<<<ceil_div_i(n, 256), 256, 0, ctx->stream>>>(rgb, H, W, Wp, luma);
    CKL(ctx);
    return MEMCI_OK;
}

int memci_ensure_luma(memci_ctx *ctx, int slot)
{
    FrameSlot &f = ctx->frames[slot];
    if (f.luma_ok) return MEMCI_OK;
    int rc = memci_luma_plane(ctx, f.rgb, ctx->H, ctx->W, ctx->Wp, f.luma);
    if (rc) return rc;
    f.luma_ok = 1;
    return MEMCI_OK;
}

// ---------------------------------------------------------------------------------------
// One pyramid level, bit-exact with ME/hbma.py:172-176:
//   convolve(im / 255.0, w[:, :, None], mode='constant')[::2, ::2] * 255.0 -> astype(uint8)
// f64 divide, taps accumulated from 0.0 in raster order of the image offsets, zero border,
// truncation on the uint8 store.  Products with the power-of-two weights are
// exact, so the only roundings are the divide, the eight adds and the final multiply.
// ---------------------------------------------------------------------------------------
__global__ void pyr_down_kernel(const uint8_t *__restrict__ in, int H, int W, uint8_t *__restrict__ out, int Ho, int Wo)
{
    // v / 255.0 for the 256 byte values, one correctly rounded f64 divide each (blockDim.x == 256)
    __shared__ double lut[256];
    lut[threadIdx.x] = __ddiv_rn((double)threadIdx.x, 255.0);
    __syncthreads();
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Ho * Wo) return;
    int oy = idx / Wo, ox = idx - oy * Wo;
    int y = 2 * oy, x = 2 * ox;
    const double wt[3] = {0.0625, 0.125, 0.0625};
    double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            int yy = y + dy, xx = x + dx;
            double w = (dy == 0 ? 2.0 : 1.0) * wt[dx + 1];
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const uint8_t *p = in + ((size_t)yy * W + xx) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    acc[c] = __dadd_rn(acc[c], __dmul_rn(lut[p[c]], w));
                }
            }
            // out-of-frame taps add 0.0 * w = +0.0, which never changes acc
        }
    }
    uint8_t *o = out + ((size_t)oy * Wo + ox) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = (uint8_t)(int)__dmul_rn(acc[c], 255.0);
}

int memci_ensure_pyramid(memci_ctx *ctx, int slot, int levels)
{
    if (levels > MEMCI_MAX_PYR) return memci_fail(ctx, MEMCI_ERR_ARG, "hbma steps %d > %d", levels, MEMCI_MAX_PYR);
    FrameSlot &f = ctx->frames[slot];
    int h = ctx->H, w = ctx->W;
    const uint8_t *prev = f.rgb;
    for (int l = 1; l <= levels; ++l) {
        int ho = (h + 1) / 2, wo = (w + 1) / 2;
        int wop = (wo + 3) & ~3;
        if (l > f.pyr_levels) {
            size_t need_rgb = (size_t)ho * wo * 3, need_l = (size_t)ho * wop * 4;
            if (f.pyr_cap[l] < need_rgb) {
                if (f.pyr_rgb[l]) cudaFree(f.pyr_rgb[l]);
                if (f.pyr_luma[l]) cudaFree(f.pyr_luma[l]);
                f.pyr_rgb[l] = nullptr; f.pyr_luma[l] = nullptr; f.pyr_cap[l] = 0;
                CK(ctx, cudaMalloc(&f.pyr_rgb[l], need_rgb));
                CK(ctx, cudaMalloc(&f.pyr_luma[l], need_l));
                f.pyr_cap[l] = need_rgb;
            }
            int n = ho * wo;
            // algorithmic bytes of one level: the finer level read once (3 B/px), the coarser level written as RGB (3 B/px)
            // and read back + written as luma (3 + 4 B/px)
            const int prof = memci_prof_begin(ctx, MEMCI_PROF_HBM_SMALL, 3.0 * h * w + 10.0 * ho * wo);
            pyr_down_kernel<<<ceil_div_i(n, 256), 256, 0, ctx->stream>>>(prev, h, w, f.pyr_rgb[l], ho, wo);
            CKL(ctx);
            int rc = memci_luma_plane(ctx, f.pyr_rgb[l], ho, wo, wop, f.pyr_luma[l]);
            memci_prof_end(ctx, prof);
            if (rc) return rc;
            f.pyr_levels = l;
        }
        prev = f.pyr_rgb[l];
        h = ho; w = wo;
    }
    return MEMCI_OK;
}
