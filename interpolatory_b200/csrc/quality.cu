// Frame-quality metrics of the Middlebury harness on the device (reference: src/Benchmark.py:46-47 / :105-106,
// skimage.metrics.peak_signal_noise_ratio and structural_similarity with data_range=255, multichannel=True,
// default win_size 7, uniform window, sample covariance, K1 0.01, K2 0.03).
//
// Per channel and per pixel at least 3 px inside the frame (skimage crops (win_size-1)/2 before the mean, so
// the filter's border mode never matters):
//     ux = mean7x7(X)  uy = mean7x7(Y)  vx = 49/48 (mean(XX) - ux ux)  vy likewise  vxy = 49/48 (mean(XY) - ux uy)
//     S  = (2 ux uy + C1)(2 vxy + C2) / ((ux^2 + uy^2 + C1)(vx + vy + C2)),   C1 = (0.01*255)^2, C2 = (0.03*255)^2
// and SSIM = mean of S over those pixels and the three channels.  The 49-tap sums of uint8 data are exact
// integers (sum XX <= 49 * 65025), so only the final f64 expression rounds; skimage's running-sum filter
// rounds a little differently (relative 1e-15), which is why this metric is compared with a tolerance.
// MSE = sum (X - Y)^2 / (3 H W) is an exact integer sum.
//
// One CTA per 32 x 8 tile: both frames' (8+6) x (32+6) halo tile in shared memory, horizontal 7-tap sums of
// (X, Y, XX, YY, XY) per row, then the vertical 7-tap sums per pixel.  Per-CTA partial sums go to a buffer that
// a second single-CTA kernel adds in a fixed order (bit-reproducible from run to run, no floating atomics).
#include "memci_internal.cuh"

#define Q_TW 32
#define Q_TH 8
#define Q_HALO 3
#define Q_SW (Q_TW + 2 * Q_HALO)
#define Q_SH (Q_TH + 2 * Q_HALO)

__global__ void __launch_bounds__(Q_TW * Q_TH)
quality_tile_kernel(const uint8_t *__restrict__ A, const uint8_t *__restrict__ B, int H, int W,
                    double *__restrict__ part_ssim, unsigned long long *__restrict__ part_sse)
{
    __shared__ uint8_t sa[Q_SH][Q_SW * 3], sb[Q_SH][Q_SW * 3];
    __shared__ int hs[5][Q_SH][Q_TW * 3];                  // horizontal sums: X, Y, XX, YY, XY
    __shared__ double red_s[Q_TW * Q_TH / 32];
    __shared__ unsigned long long red_e[Q_TW * Q_TH / 32];
    const int tid = threadIdx.y * Q_TW + threadIdx.x;
    const int tx0 = blockIdx.x * Q_TW, ty0 = blockIdx.y * Q_TH;
    for (int e = tid; e < Q_SH * Q_SW * 3; e += Q_TW * Q_TH) {
        const int r = e / (Q_SW * 3), cb = e - r * (Q_SW * 3);
        const int gr = ty0 - Q_HALO + r, gcb = (tx0 - Q_HALO) * 3 + cb;
        const bool in = gr >= 0 && gr < H && gcb >= 0 && gcb < W * 3;
        sa[r][cb] = in ? A[(size_t)gr * W * 3 + gcb] : (uint8_t)0;
        sb[r][cb] = in ? B[(size_t)gr * W * 3 + gcb] : (uint8_t)0;
    }
    __syncthreads();
    for (int e = tid; e < Q_SH * Q_TW * 3; e += Q_TW * Q_TH) {
        const int r = e / (Q_TW * 3), cb = e - r * (Q_TW * 3);   // cb = 3 * column + channel
        int x = 0, y = 0, xx = 0, yy = 0, xy = 0;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int a = sa[r][cb + 3 * k], b = sb[r][cb + 3 * k];
            x += a; y += b; xx += a * a; yy += b * b; xy += a * b;
        }
        hs[0][r][cb] = x; hs[1][r][cb] = y; hs[2][r][cb] = xx; hs[3][r][cb] = yy; hs[4][r][cb] = xy;
    }
    __syncthreads();
    const int q_r = ty0 + threadIdx.y, q_c = tx0 + threadIdx.x;
    double s_sum = 0.0;
    unsigned long long sse = 0ull;
    if (q_r < H && q_c < W) {
        const bool interior = q_r >= Q_HALO && q_r < H - Q_HALO && q_c >= Q_HALO && q_c < W - Q_HALO;
        const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
        const double cov_norm = 49.0 / 48.0;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const int cb = threadIdx.x * 3 + ch;
            const int d = (int)sa[threadIdx.y + Q_HALO][cb + 3 * Q_HALO] - (int)sb[threadIdx.y + Q_HALO][cb + 3 * Q_HALO];
            sse += (unsigned long long)(d * d);
            if (interior) {
                int x = 0, y = 0, xx = 0, yy = 0, xy = 0;
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    x += hs[0][threadIdx.y + k][cb]; y += hs[1][threadIdx.y + k][cb];
                    xx += hs[2][threadIdx.y + k][cb]; yy += hs[3][threadIdx.y + k][cb]; xy += hs[4][threadIdx.y + k][cb];
                }
                const double ux = (double)x / 49.0, uy = (double)y / 49.0;
                const double uxx = (double)xx / 49.0, uyy = (double)yy / 49.0, uxy = (double)xy / 49.0;
                const double vx = cov_norm * (uxx - ux * ux), vy = cov_norm * (uyy - uy * uy), vxy = cov_norm * (uxy - ux * uy);
                const double A1 = 2.0 * ux * uy + C1, A2 = 2.0 * vxy + C2;
                const double B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
                s_sum += (A1 * A2) / (B1 * B2);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_sum += __shfl_down_sync(0xffffffffu, s_sum, o);
        sse += __shfl_down_sync(0xffffffffu, sse, o);
    }
    if ((tid & 31) == 0) { red_s[tid >> 5] = s_sum; red_e[tid >> 5] = sse; }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0; unsigned long long e = 0ull;
        for (int i = 0; i < Q_TW * Q_TH / 32; ++i) { s += red_s[i]; e += red_e[i]; }
        const int bi = blockIdx.y * gridDim.x + blockIdx.x;
        part_ssim[bi] = s; part_sse[bi] = e;
    }
}

// fixed-order sum of the per-tile partials: thread t adds tiles t, t+1024, ...; then a shared-memory tree
__global__ void __launch_bounds__(1024)
quality_sum_kernel(const double *__restrict__ part_ssim, const unsigned long long *__restrict__ part_sse, int n, double *__restrict__ result)
{
    __shared__ double s_s[1024];
    __shared__ unsigned long long s_e[1024];
    double s = 0.0; unsigned long long e = 0ull;
    for (int i = threadIdx.x; i < n; i += 1024) { s += part_ssim[i]; e += part_sse[i]; }
    s_s[threadIdx.x] = s; s_e[threadIdx.x] = e;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_s[threadIdx.x] += s_s[threadIdx.x + o]; s_e[threadIdx.x] += s_e[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { result[0] = (double)s_e[0]; result[1] = s_s[0]; }   // SSE < 2^53: exact as a double
}

int memci_quality_run(memci_ctx *ctx, int slot_a, int slot_b, double *h_result)
{
    const int H = ctx->H, W = ctx->W;
    const dim3 grid(ceil_div_i(W, Q_TW), ceil_div_i(H, Q_TH));
    const size_t n = (size_t)grid.x * grid.y;
    size_t cap = ctx->quality_cap;
    int rc = memci_reserve(ctx, (void **)&ctx->d_quality, &cap, n * 16 + 16);
    ctx->quality_cap = cap;
    if (rc) return rc;
    double *result = reinterpret_cast<double *>(ctx->d_quality);
    double *part_ssim = result + 2;
    unsigned long long *part_sse = reinterpret_cast<unsigned long long *>(part_ssim + n);
    quality_tile_kernel<<<grid, dim3(Q_TW, Q_TH), 0, ctx->stream>>>(ctx->frames[slot_a].rgb, ctx->frames[slot_b].rgb, H, W, part_ssim, part_sse);
    CKL(ctx);
    quality_sum_kernel<<<1, 1024, 0, ctx->stream>>>(part_ssim, part_sse, (int)n, result);
    CKL(ctx);
    CK(ctx, cudaMemcpyAsync(h_result, result, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return MEMCI_OK;
}
