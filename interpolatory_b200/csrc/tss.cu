// Three-step search (ME/tss.py:18-61) on sm_100a.
//
// 9 candidates per step and 3 steps per block: ~27 SADs per block, three orders of magnitude less
// work than full search, so the kernel is the simple warp-per-block form.  Lanes 0..8 evaluate the
// 3x3 pattern of one step (same luma SAD as full search, incl. the f64 truncation fix-up, N2); the
// stateful update rule -- strict SAD improvement, ties resolved by the precedence table against the
// CURRENT best's pattern index, which may come from an earlier step (:51) -- is then replayed in
// pattern order by every lane.
#include "memci_internal.cuh"

__global__ void tss_kernel(const int32_t *__restrict__ Ls, const int32_t *__restrict__ Lt,
                           const uint8_t *__restrict__ rgb_s, const uint8_t *__restrict__ rgb_t,
                           int H, int W, int Wp, int bs, int steps, int nbr, int nbc,
                           float2 *__restrict__ out_vec, float *__restrict__ out_sad)
{
    const int prec[9] = {1, 2, 1, 2, 3, 2, 1, 2, 1};                       // :21
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int blk = warp; blk < nbr * nbc; blk += nwarps) {
        const int br = blk / nbc, bc = blk - br * nbc;
        const int s_row = br * bs, s_col = bc * bs;
        const int tmr = (s_row + bs >= H) ? s_row + 1 : H - bs;             // :36-39
        const int tmc = (s_col + bs >= H) ? s_col + 1 : W - bs;             // :40-43, sic: H
        bool have_center = false;
        bool have_best = false;                                             // lowest_sad = inf
        unsigned lowest_sad = 0u;
        int li_r = 0, li_c = 0, lowest_i = 0;
        int curr_row = s_row, curr_col = s_col;
        for (int step = steps; step > 0; --step) {
            const int S = 1 << (step - 1);
            const int i = lane;                                             // pattern index, row-major
            const int t_row = curr_row + (i / 3 - 1) * S, t_col = curr_col + (i % 3 - 1) * S;
            bool valid = i < 9 && (t_row != curr_row || t_col != curr_col || !have_center) &&
                         (t_row >= 0 && t_row <= tmr && t_col >= 0 && t_col <= tmc);   // :48 (inclusive bounds)
            unsigned sad = 0u;
            if (valid) {
                const unsigned Sm = sad1000_global(Ls, Lt, H, W, Wp, s_row, s_col, t_row, t_col, bs);
                sad = Sm / 1000u;
                if (Sm != 0 && sad * 1000u == Sm)
                    sad = (unsigned)exact_cost_f64(rgb_s, rgb_t, H, W, s_row, s_col, t_row, t_col, bs, 0);
            }
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const unsigned sj = __shfl_sync(0xffffffffu, sad, j);
                if ((vmask >> j) & 1u) {
                    if (!have_best || sj < lowest_sad || (sj == lowest_sad && prec[j] > prec[lowest_i])) {   // :51
                        have_best = true; lowest_sad = sj; lowest_i = j;
                        li_r = curr_row + (j / 3 - 1) * S; li_c = curr_col + (j % 3 - 1) * S;
                    }
                }
            }
            curr_row = li_r; curr_col = li_c;                               // :55-57
            have_center = true;
        }
        if (lane == 0) {
            out_vec[blk] = make_float2((float)(li_r - s_row), (float)(li_c - s_col));
            out_sad[blk] = have_best ? (float)lowest_sad : __int_as_float(0x7f800000);
        }
    }
}

int memci_tss_run(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int steps, MVField *out)
{
    if (bs <= 0 || bs > 64 || steps < 0 || steps > 20)
        return memci_fail(ctx, MEMCI_ERR_ARG, "tss: block_size %d / steps %d out of range", bs, steps);
    int rc;
    if ((rc = memci_ensure_luma(ctx, slot_src))) return rc;
    if ((rc = memci_ensure_luma(ctx, slot_tgt))) return rc;
    const int nbr = ceil_div_i(ctx->H, bs), nbc = ceil_div_i(ctx->W, bs);
    if ((rc = memci_mv_reserve(ctx, out, bs, nbr, nbc, bs, nbr, nbc))) return rc;
    FrameSlot &fs = ctx->frames[slot_src], &ft = ctx->frames[slot_tgt];
    tss_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(fs.luma, ft.luma, fs.rgb, ft.rgb, ctx->H, ctx->W, ctx->Wp, bs, steps,
                                                          nbr, nbc, out->vec, out->sad);
    CKL(ctx);
    out->valid = 1;
    return MEMCI_OK;
}
