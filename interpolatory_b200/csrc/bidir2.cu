// bidir2: MCI/Bidirectional_2.py -- irregular-grid expanded-block weighted motion compensation (IEWMC,
// :49-144), error-adaptive combination (:147-166), block-wise directional hole interpolation (BDHI,
// :168-286) and unpad/downconvert (:288-290), after Wang et al., "Motion-Compensated Frame Rate
// Up-Conversion -- Part II".
//
// The reference SCATTERS every 4x4 block's 8x8 expanded block (both frames blended, weighted by a fixed
// 8x8 kernel) to a position displaced by int(dist * mv), accumulating three float32 planes in
// block-raster order.  float32 addition is not associative, so we GATHER: each target pixel walks the
// blocks that can reach it in the same raster order and adds their contributions in float32, one by
// one -- bit-identical sums, no atomics.  Per-block quantities (truncated vector, truncated displacement,
// occlusion flag) are computed once by b2_blocks_kernel; the frame kernel stages the tile's neighbourhood
// of them in shared memory, gathers both directions, normalises, combines (EAC), and -- because a 32x8
// tile is exactly four 8x8 hole-fill blocks -- runs BDHI on the tile in shared memory and writes uint8.
//
// numba's typing is reproduced term by term; every such line below carries a comment saying which
// promotion it follows (DESIGN.md section 2 lists how each was probed against the reference).
#include "memci_internal.cuh"

#define B2_MBS 4                 // min_block_size (fixed by MEMCIBaseInterpolator.py:44)
#define B2_EB 8                  // expanded block = 2 * min_block_size
#define B2_TW 32
#define B2_TH 8
#define B2_MAXB 640              // staged blocks per direction

struct B2Params {
    int H, W, Hr, Wr;            // frame, and the region BDHI reads: multiples of 8 covering the frame
    int nbr, nbc;                // 4x4 block grid
    int pad;
    float dist[2];               // forward: dist; backward: float32(1 - dist)
    const uint8_t *f1[2], *f2[2];          // per direction: F1 (blended with weight 1-dist, origin of the expanded block), F2
    const int4 *blk[2];
    const int *counters;         // [5] max |T|, [6] blocks whose shifted slices leave the padded arrays
    uint8_t *out;
    float w[64];
};

// ---- per-block quantities: MV_i -> (Xb, Yb, TXb, TYb, occluded)   (:62-74, :87) -----------------
__global__ void b2_blocks_kernel(const float2 *__restrict__ vec, int mv_cell, int mv_cols, const float *__restrict__ sad,
                                 int sad_cell, int sad_cols, float dist, int H, int W, int nbr, int nbc, int pad,
                                 int4 *__restrict__ out, int *__restrict__ counters)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int tmax = 0, bad = 0;
    if (i < nbr * nbc) {
        const int br = i / nbc, bc = i - br * nbc;
        const int r = br * B2_MBS, c = bc * B2_MBS;
        const float2 v = vec[(size_t)(r / mv_cell) * mv_cols + c / mv_cell];
        const float s = sad[(size_t)(r / sad_cell) * sad_cols + c / sad_cell];
        const int Xb = (int)v.x, Yb = (int)v.y;                                  // int(MV_i[0]): truncation
        const int TXb = (int)__dmul_rn((double)dist, (double)Xb);                 // float32 * int64 -> float64
        const int TYb = (int)__dmul_rn((double)dist, (double)Yb);
        const double sn = __ddiv_rn((double)s, (double)(B2_MBS * B2_MBS * 3));    // float32 / int64 -> float64
        const int occluded = !(sn < 250.0);
        out[i] = make_int4(Xb, Yb, TXb, (TYb << 1) | occluded);
        tmax = max(abs(TXb), abs(TYb));
        // the reference's slices [r0+X : r0+X+8] must stay inside the padded arrays (pad = 16)
        const int r0 = r + pad - B2_MBS / 2, c0 = c + pad - B2_MBS / 2, Hp = H + 2 * pad, Wp = W + 2 * pad;
        bad = (r0 + Xb < 0 || r0 + Xb + B2_EB > Hp || c0 + Yb < 0 || c0 + Yb + B2_EB > Wp ||
               r0 + TXb < 0 || r0 + TXb + B2_EB > Hp || c0 + TYb < 0 || c0 + TYb + B2_EB > Wp);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    const unsigned nbad = __popc(__ballot_sync(0xffffffffu, bad));
    if ((threadIdx.x & 31) == 0) {
        if (tmax > 0) atomicMax(&counters[5], tmax);
        if (nbad) { atomicAdd(&counters[6], (int)nbad); atomicAdd(&counters[15], (int)nbad); }   // [15]: running total (memci_bidir2_out_of_range_total)
    }
}

__device__ __forceinline__ int floor_div4(int a) { return a >> 2; }        // arithmetic shift = floor for negatives

// ---- fill_holes on one 8x8 block in shared memory (:168-262) -----------------------------------------
// a_array: the four directional mean absolute differences.  The reference accumulates them as float32
// running sums in (i, j) raster order; the 16 chains (4 directions x {3 channels, count}) are independent,
// so 16 threads of the block's first two rows each run ONE chain in that order.  sums[d][k] then holds the
// raw sum (k < 3) or the count (k == 3); the quotient is taken by the reader.
__device__ void b2_direction_chain(const float (*fc)[B2_TW][3], int c0, int d, int k, float *sum_out)
{
    const int di = d == 3 ? -1 : (d == 1 ? 0 : 1), dj = d == 0 ? 0 : 1;      // (1,0) (0,1) (1,1) (-1,1)
    float a = 0.1f;
    for (int i = 0; i < B2_EB - 1; ++i) {
        if (d == 3 && i < 1) continue;                                      // (i-1 > -1), :192
        for (int j = 0; j < B2_EB - 1; ++j) {
            const float *pi = fc[i][c0 + j];
            if (pi[0] == -1.f) continue;
            const float *q = fc[i + di][c0 + j + dj];
            if (q[0] == -1.f) continue;
            if (k < 3) a = __fadd_rn(a, fabsf(__fsub_rn(pi[k], q[k])));
            else if (d != 3) a = __fadd_rn(a, 1.f);                          // the anti-diagonal counter is never incremented (:193-194)
        }
    }
    *sum_out = a;
}

// one undetermined pixel (i, j) of the block starting at column c0: four directional estimates, truncated
// to int64 by the reference's int array (:236-244), combined with weights 1 / a (:246-256)
__device__ void b2_fill_pixel(const float (*fc)[B2_TW][3], int c0, const float (*sums)[4], int i, int j, float *res)
{
    const int dir[4][2] = {{1, 0}, {0, 1}, {1, 1}, {-1, 1}};
    float A[3] = {0.f, 0.f, 0.f}, pix[3] = {0.f, 0.f, 0.f};
    float aa[4][3];
#pragma unroll
    for (int d = 0; d < 4; ++d)
#pragma unroll
        for (int k = 0; k < 3; ++k) aa[d][k] = __fdiv_rn(sums[d][k], sums[d][3]);      // ah = ah[0:3] / ah[3]  (:196-199)
    for (int d = 0; d < 4; ++d) {
        const float *p1 = nullptr, *p2 = nullptr;
        double d1 = 0.0, d2 = 0.0;
        int n1 = i, m1 = j, n2 = i, m2 = j;
        while (!p1 && n1 >= 0 && n1 < B2_EB && m1 >= 0 && m1 < B2_EB) {
            const float *q = fc[n1][c0 + m1];
            if (q[0] != -1.f) { p1 = q; d1 = __dsqrt_rn((double)((i - n1) * (i - n1) + (j - m1) * (j - m1))); }   // np.hypot of small integers
            n1 -= dir[d][0]; m1 -= dir[d][1];
        }
        while (!p2 && n2 >= 0 && n2 < B2_EB && m2 >= 0 && m2 < B2_EB) {
            const float *q = fc[n2][c0 + m2];
            if (q[0] != -1.f) { p2 = q; d2 = __dsqrt_rn((double)((i - n2) * (i - n2) + (j - m2) * (j - m2))); }
            n2 += dir[d][0]; m2 += dir[d][1];
        }
        if (!p1 && !p2) continue;
        for (int k = 0; k < 3; ++k) {
            long long pv;
            if (p1 && p2) pv = (long long)__ddiv_rn(__dadd_rn(__dmul_rn(d2, (double)p1[k]), __dmul_rn(d1, (double)p2[k])), __dadd_rn(d1, d2));
            else pv = (long long)(p1 ? p1[k] : p2[k]);
            // in-place adds on float32 arrays: right-hand side rounded to float32, then a float32 add
            A[k] = __fadd_rn(A[k], __fdiv_rn(1.f, aa[d][k]));
            pix[k] = __fadd_rn(pix[k], __double2float_rn(__ddiv_rn((double)pv, (double)aa[d][k])));
        }
    }
    for (int k = 0; k < 3; ++k) res[k] = __fdiv_rn(pix[k], A[k]);
}

// Contribution of one block to one target pixel (x, y): local offset (i, j) inside the block's 8x8 expanded
// block, one float32 add per accumulator -- IEWMC_helper :87-108.
__device__ __forceinline__ void b2_contribute(const B2Params &P, const int4 b, int i, int j, int y, int x,
                                              const uint8_t *__restrict__ F1, const uint8_t *__restrict__ F2, float dist,
                                              double one_minus, float (&acc)[3], float (&wm)[3], float &ws)
{
    const int H = P.H, W = P.W;
    const float wk = P.w[i * B2_EB + j];
    const int sy = y - b.z, sx = x - (b.w >> 1);                 // position inside the (zero-padded) frames
    const int ty = sy + b.x, tx = sx + b.y;
    const bool in1 = (unsigned)sy < (unsigned)H && (unsigned)sx < (unsigned)W;
    const bool in2 = (unsigned)ty < (unsigned)H && (unsigned)tx < (unsigned)W;
    const uint8_t *p1 = F1 + ((size_t)sy * W + sx) * 3, *p2 = F2 + ((size_t)ty * W + tx) * 3;
    const bool occluded = b.w & 1;
    const double wkd = (double)wk;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int a = in1 ? p1[k] : 0, bb = in2 ? p2[k] : 0;
        float av;
        if (!occluded) {
            // (double)a without the conversion unit: 2^52 + a has a's bits in its low mantissa word
            const double ad = __dsub_rn(__hiloint2double(0x43300000, a), 4503599627370496.0);
            const double ta = __dmul_rn(one_minus, ad);                        // float64 * uint8
            const float tb = __fmul_rn(dist, (float)bb);                       // float32 * uint8
            av = __double2float_rn(__dmul_rn(wkd, __dadd_rn(ta, (double)tb)));
        } else {
            av = __fmul_rn(wk, (float)a);                                      // float32 * uint8
        }
        acc[k] = __fadd_rn(acc[k], av);
        // float32(float64(w) * |d|): the float64 product of a 24-bit and an 8-bit number is exact, so the single
        // rounding to float32 equals the float32 product
        wm[k] = __fadd_rn(wm[k], __fmul_rn(wk, (float)abs(a - bb)));
    }
    ws = __fadd_rn(ws, wk);
}

__global__ void __launch_bounds__(B2_TW * B2_TH)
b2_frame_kernel(const __grid_constant__ B2Params P)
{
    __shared__ int4 sblk[2][B2_MAXB];
    __shared__ unsigned stl[2][B2_MAXB];          // top row | left column << 16 of each staged block's shifted expanded block (signed 16-bit each)
    __shared__ float fc[B2_TH][B2_TW][3];
    __shared__ float dsum[B2_TW / B2_EB][4][4];
    __shared__ int blk_hole[B2_TW / B2_EB];
    __shared__ int tile_tmax[2][2];               // largest |T| per direction and axis among the staged blocks: tighter per-pixel loops than the frame-wide maximum

    const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;
    const int tx0 = blockIdx.x * B2_TW, ty0 = blockIdx.y * B2_TH;
    const int x = tx0 + lx, y = ty0 + ly;
    const int H = P.H, W = P.W;
    const int tmax = P.counters[5];
    // blocks whose shifted expanded block (rows 4br-2+T .. 4br+5+T) can reach the tile
    const int br_lo = max(0, floor_div4(ty0 - 5 - tmax)), br_hi = min(P.nbr - 1, floor_div4(ty0 + B2_TH - 1 + 2 + tmax));
    const int bc_lo = max(0, floor_div4(tx0 - 5 - tmax)), bc_hi = min(P.nbc - 1, floor_div4(tx0 + B2_TW - 1 + 2 + tmax));
    const int nrb = br_hi - br_lo + 1, ncb = bc_hi - bc_lo + 1;
    const bool staged = nrb > 0 && ncb > 0 && nrb * ncb <= B2_MAXB;
    if (tid < B2_TW / B2_EB) blk_hole[tid] = 0;
    if (tid < 4) (&tile_tmax[0][0])[tid] = 0;
    __syncthreads();
    if (staged) {
        int tm[2][2] = {{0, 0}, {0, 0}};
        for (int i = tid; i < nrb * ncb; i += B2_TW * B2_TH) {
            const int r = i / ncb, c = i - r * ncb;
            const size_t g = (size_t)(br_lo + r) * P.nbc + bc_lo + c;
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const int4 b = P.blk[d][g];
                sblk[d][i] = b;
                stl[d][i] = (unsigned)((4 * (br_lo + r) - 2 + b.z) & 0xffff) | ((unsigned)(4 * (bc_lo + c) - 2 + (b.w >> 1)) << 16);
                tm[d][0] = max(tm[d][0], abs(b.z)); tm[d][1] = max(tm[d][1], abs(b.w >> 1));
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int m = __reduce_max_sync(0xffffffffu, tm[q >> 1][q & 1]);
            if (lx == 0 && m > 0) atomicMax(&tile_tmax[q >> 1][q & 1], m);
        }
    }
    __syncthreads();

    const bool in_region = x < P.Wr;            // y < Hr always (grid)
    float out_c[3] = {-1.f, -1.f, -1.f};
    if (nrb > 0 && ncb > 0) {
        float Fd[2][3], Ed[2][3];
        bool has[2];
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const uint8_t *F1 = P.f1[d], *F2 = P.f2[d];
            const float dist = P.dist[d];
            const double one_minus = __dsub_rn(1.0, (double)dist);          // (1 - dist): int64 - float32 -> float64
            float acc[3] = {0.f, 0.f, 0.f}, wm[3] = {0.f, 0.f, 0.f}, ws = 0.f;
            // every block that can reach the tile is staged, so the staged maxima bound this tile's loops exactly
            const int tr = staged ? tile_tmax[d][0] : tmax, tc = staged ? tile_tmax[d][1] : tmax;
            const int brp_lo_w = max(br_lo, floor_div4(y - 5 - tr)), brp_hi_w = min(br_hi, floor_div4(y + 2 + tr));    // warp-uniform
            const int bcp_lo = max(bc_lo, floor_div4(x - 5 - tc)), bcp_hi = min(bc_hi, floor_div4(x + 2 + tc));
            // Per-lane loops over the blocks that can reach (x, y), in block raster order = the reference's accumulation
            // order.  Lanes of a warp share y and walk bc ranges shifted with x, so neighbouring lanes test neighbouring
            // blocks in the same iteration and hit together.  (A warp-compacted list of the blocks passing the row test
            // was tried: fewer tests, but every entry then hits only 8 of 32 lanes -- 40 % slower.)
            if (in_region) {
                if (staged) {
                    for (int br = brp_lo_w; br <= brp_hi_w; ++br) {
                        const int row0 = (br - br_lo) * ncb - bc_lo;
                        for (int bc = bcp_lo; bc <= bcp_hi; ++bc) {
                            const unsigned tl = stl[d][row0 + bc];
                            const int i = y - (short)(tl & 0xffffu), j = x - ((int)tl >> 16);
                            if ((i | j) & ~(B2_EB - 1)) continue;                  // both in [0, 8)
                            b2_contribute(P, sblk[d][row0 + bc], i, j, y, x, F1, F2, dist, one_minus, acc, wm, ws);
                        }
                    }
                } else {
                    for (int br = brp_lo_w; br <= brp_hi_w; ++br) {
                        for (int bc = bcp_lo; bc <= bcp_hi; ++bc) {
                            const int4 b = P.blk[d][(size_t)br * P.nbc + bc];
                            const int i = y - (4 * br - 2 + b.z), j = x - (4 * bc - 2 + (b.w >> 1));
                            if ((i | j) & ~(B2_EB - 1)) continue;
                            b2_contribute(P, b, i, j, y, x, F1, F2, dist, one_minus, acc, wm, ws);
                        }
                    }
                }
            }
            has[d] = !(ws <= 0.f);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                Fd[d][k] = has[d] ? __fdiv_rn(acc[k], ws) : -1.f;
                Ed[d][k] = has[d] ? __fdiv_rn(wm[k], ws) : -1.f;
            }
        }
        // Error_adaptive_combination (:147-166): the "!= -1" tests look at channel 0 of the normalised frames
        const bool hf = Fd[0][0] != -1.f, hb = Fd[1][0] != -1.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (hf && hb) {
                const double num = __dadd_rn(__dmul_rn(__dadd_rn((double)Ed[1][k], 1.0), (double)Fd[0][k]),
                                             __dmul_rn(__dadd_rn((double)Ed[0][k], 1.0), (double)Fd[1][k]));
                const float es = __fadd_rn(Ed[0][k], Ed[1][k]);
                out_c[k] = __double2float_rn(__ddiv_rn(num, __dadd_rn((double)es, 2.0)));
            } else if (hf) out_c[k] = Fd[0][k];
            else if (hb) out_c[k] = Fd[1][k];
        }
    }
    fc[ly][lx][0] = out_c[0]; fc[ly][lx][1] = out_c[1]; fc[ly][lx][2] = out_c[2];
    const bool hole = in_region && out_c[0] == -1.f;
    if (hole) blk_hole[lx >> 3] = 1;
    __syncthreads();
    // BDHI (:265-286): a block with a hole gets fill_holes(block); statistics by one thread per block
    const int myblk = lx >> 3;
    if (ly < 2 && blk_hole[myblk] && tx0 + myblk * B2_EB < P.Wr) {
        const int chain = ly * 8 + (lx & 7);                                   // 16 threads per block: chain = direction * 4 + channel
        b2_direction_chain(fc, myblk * B2_EB, chain >> 2, chain & 3, &dsum[myblk][chain >> 2][chain & 3]);
    }
    __syncthreads();
    if (hole) b2_fill_pixel(fc, myblk * B2_EB, dsum[myblk], ly, lx & 7, out_c);
    if (y < H && x < W) {
        uint8_t *o = P.out + ((size_t)y * W + x) * 3;                          // unpad_and_downconvert: astype(uint8)
        o[0] = (uint8_t)__float2int_rz(out_c[0]); o[1] = (uint8_t)__float2int_rz(out_c[1]); o[2] = (uint8_t)__float2int_rz(out_c[2]);
    }
}

int memci_bidir2_run(memci_ctx *ctx, int slot_src, int slot_tgt, const MVField *fwd, const MVField *bwd,
                     float dist_fwd, float dist_bwd, int mbs, int pad, const float *w64, int out_slot)
{
    const int H = ctx->H, W = ctx->W;
    if (mbs != B2_MBS) return memci_fail(ctx, MEMCI_ERR_ARG, "bidir2: min_block_size %d not supported (the reference fixes 4)", mbs);
    if (pad < B2_EB) return memci_fail(ctx, MEMCI_ERR_ARG, "bidir2: pad_size %d too small (need >= 8)", pad);
    const int nbr = ceil_div_i(H, B2_MBS), nbc = ceil_div_i(W, B2_MBS);
    const size_t nblk = (size_t)nbr * nbc;
    int rc;
    for (int d = 0; d < 2; ++d)
        if ((rc = memci_reserve(ctx, (void **)&ctx->d_b2blk[d], &ctx->b2blk_cap[d], nblk * sizeof(int4)))) return rc;
    CK(ctx, cudaMemsetAsync(ctx->d_counters + 5, 0, 2 * sizeof(int), ctx->stream));
    const MVField *fl[2] = {fwd, bwd};
    const float dists[2] = {dist_fwd, dist_bwd};
    for (int d = 0; d < 2; ++d) {
        const MVField *f = fl[d];
        b2_blocks_kernel<<<ceil_div_i((int)nblk, 256), 256, 0, ctx->stream>>>(f->vec, f->mv_cell, f->mv_cols, f->sad, f->sad_cell, f->sad_cols,
                                                                             dists[d], H, W, nbr, nbc, pad, ctx->d_b2blk[d], ctx->d_counters);
        CKL(ctx);
    }
    B2Params P;
    memset(&P, 0, sizeof(P));
    P.H = H; P.W = W; P.Hr = ceil_div_i(H, B2_EB) * B2_EB; P.Wr = ceil_div_i(W, B2_EB) * B2_EB;
    P.nbr = nbr; P.nbc = nbc; P.pad = pad;
    P.dist[0] = dist_fwd; P.dist[1] = dist_bwd;
    P.f1[0] = ctx->frames[slot_src].rgb; P.f2[0] = ctx->frames[slot_tgt].rgb;      // IEWMC(fwr, source, target, dist)
    P.f1[1] = ctx->frames[slot_tgt].rgb; P.f2[1] = ctx->frames[slot_src].rgb;      // IEWMC(bwr, target, source, 1 - dist)
    P.blk[0] = ctx->d_b2blk[0]; P.blk[1] = ctx->d_b2blk[1];
    P.counters = ctx->d_counters;
    P.out = ctx->frames[out_slot].rgb;
    memcpy(P.w, w64, sizeof(P.w));
    dim3 grid(ceil_div_i(P.Wr, B2_TW), P.Hr / B2_TH);
    b2_frame_kernel<<<grid, B2_TW * B2_TH, 0, ctx->stream>>>(P);
    CKL(ctx);
    FrameSlot &o = ctx->frames[out_slot];
    o.valid = 1; o.luma_ok = 0; o.luma8_ok = 0; o.pyr_levels = 0;
    return MEMCI_OK;
}
