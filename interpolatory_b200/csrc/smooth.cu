// 3x3 motion-vector smoothing (smoothing/threeXthree_mv_smoothing.py:14-27) and the
// block <-> dense conversions of the MV field.
#include "memci_internal.cuh"

// One thread per filter_size x filter_size output cell.  The reference samples the DENSE field
// at stride filter_size (mv_field[low:high:fs]); on the block-level field that is a lookup of
// the block containing each sampled pixel.  Only (dy,dx) are rewritten; SAD keeps its grid.
__global__ void smooth_kernel(const float2 *__restrict__ in, int in_cell, int in_cols, int H, int W, int mode, int fs,
                              float2 *__restrict__ out, int out_rows, int out_cols)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= out_rows * out_cols) return;
    int i = idx / out_cols, j = idx - i * out_cols;
    int row = i * fs, col = j * fs;
    int low_r = max(0, row - fs), high_r = min(min(row + fs, H) + 1, H);      // :20-23, slice clamps at H
    int low_c = max(0, col - fs), high_c = min(min(col + fs, W) + 1, W);
    float vr[9], vc[9];
    int n = 0;
    for (int r = low_r; r < high_r; r += fs)
        for (int c = low_c; c < high_c; c += fs) {
            float2 v = in[(size_t)(r / in_cell) * in_cols + c / in_cell];
            vr[n] = v.x; vc[n] = v.y; ++n;
        }
    float o_r, o_c;
    if (mode == MEMCI_FILTER_MEAN) {
        // mean_filter.py:2-4: np.mean in float32 (sums of integer-valued MVs are exact), one f32 divide
        float sr = 0.f, sc = 0.f;
        for (int k = 0; k < n; ++k) { sr = __fadd_rn(sr, vr[k]); sc = __fadd_rn(sc, vc[k]); }
        o_r = __fdiv_rn(sr, (float)n); o_c = __fdiv_rn(sc, (float)n);
    } else if (mode == MEMCI_FILTER_MEDIAN) {
        // median_filter.py:3-11: element at sorted position n//2, per component
        for (int a = 1; a < n; ++a) {
            float x = vr[a], y = vc[a];
            int b = a - 1;
            while (b >= 0 && vr[b] > x) { vr[b + 1] = vr[b]; --b; }
            vr[b + 1] = x;
            b = a - 1;
            while (b >= 0 && vc[b] > y) { vc[b + 1] = vc[b]; --b; }
            vc[b + 1] = y;
        }
        o_r = vr[n >> 1]; o_c = vc[n >> 1];
    } else {
        // weighted_mean_filter.py:4-15: weights by flat index, f32 product, f64 sums, floor-div, int32
        const float wts[9] = {1.f, 2.f, 1.f, 2.f, 6.f, 2.f, 1.f, 2.f, 1.f};
        double tr = 0.0, tc = 0.0, tw = 0.0;
        for (int k = 0; k < n; ++k) {
            tr = __dadd_rn(tr, (double)__fmul_rn(vr[k], wts[k]));
            tc = __dadd_rn(tc, (double)__fmul_rn(vc[k], wts[k]));
            tw = __dadd_rn(tw, (double)wts[k]);
        }
        o_r = (float)(int)floor(__ddiv_rn(tr, tw));
        o_c = (float)(int)floor(__ddiv_rn(tc, tw));
    }
    out[idx] = make_float2(o_r, o_c);
}

int memci_smooth_run(memci_ctx *ctx, const MVField *in, int mode, int fsz, MVField *out)
{
    if (fsz <= 0) return memci_fail(ctx, MEMCI_ERR_ARG, "filter_size must be positive");
    int rows = ceil_div_i(ctx->H, fsz), cols = ceil_div_i(ctx->W, fsz);
    int rc = memci_mv_reserve(ctx, out, fsz, rows, cols, in->sad_cell, in->sad_rows, in->sad_cols);
    if (rc) return rc;
    // algorithmic bytes: the vector grid read once (8 B per input cell), one filtered vector written per filter cell
    const int prof = memci_prof_begin(ctx, MEMCI_PROF_HBM_SMALL, 8.0 * in->mv_rows * in->mv_cols + 8.0 * rows * cols);
    smooth_kernel<<<ceil_div_i(rows * cols, 128), 128, 0, ctx->stream>>>(in->vec, in->mv_cell, in->mv_cols, ctx->H, ctx->W,
                                                                         mode, fsz, out->vec, rows, cols);
    memci_prof_end(ctx, prof);
    CKL(ctx);
    CK(ctx, cudaMemcpyAsync(out->sad, in->sad, sizeof(float) * (size_t)in->sad_rows * in->sad_cols,
                            cudaMemcpyDeviceToDevice, ctx->stream));
    out->valid = 1;
    return MEMCI_OK;
}

// ---- block-level field -> dense float32 [H][W][3]  (full_search.py:51-53, hbma.py:151-159) ----
__global__ void mv_to_dense_kernel(const float2 *__restrict__ vec, int cell, int cols, const float *__restrict__ sad,
                                   int scell, int scols, int H, int W, float *__restrict__ dense)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * W) return;
    int r = idx / W, c = idx - r * W;
    float2 v = vec[(size_t)(r / cell) * cols + c / cell];
    float s = sad[(size_t)(r / scell) * scols + c / scell];
    float *o = dense + (size_t)idx * 3;
    o[0] = v.x; o[1] = v.y; o[2] = s;
}

__global__ void dense_to_mv_kernel(const float *__restrict__ dense, int n, float2 *__restrict__ vec, float *__restrict__ sad)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const float *d = dense + (size_t)idx * 3;
    vec[idx] = make_float2(d[0], d[1]);
    sad[idx] = d[2];
}

int memci_mv_to_dense(memci_ctx *ctx, const MVField *f, float *d_dense)
{
    int n = ctx->H * ctx->W;
    mv_to_dense_kernel<<<ceil_div_i(n, 256), 256, 0, ctx->stream>>>(f->vec, f->mv_cell, f->mv_cols, f->sad, f->sad_cell,
                                                                    f->sad_cols, ctx->H, ctx->W, d_dense);
    CKL(ctx);
    return MEMCI_OK;
}

int memci_dense_to_mv(memci_ctx *ctx, const float *d_dense, MVField *f)
{
    int rc = memci_mv_reserve(ctx, f, 1, ctx->H, ctx->W, 1, ctx->H, ctx->W);
    if (rc) return rc;
    int n = ctx->H * ctx->W;
    dense_to_mv_kernel<<<ceil_div_i(n, 256), 256, 0, ctx->stream>>>(d_dense, n, f->vec, f->sad);
    CKL(ctx);
    f->valid = 1;
    return MEMCI_OK;
}
