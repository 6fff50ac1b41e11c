// Full-search block matching (ME/full_search.py) on sm_100a.
//
//   fs_tiled_kernel   persistent, TMA-staged, register-tiled luma-SAD + argmin (block sizes 8k)
//   fs_generic_kernel warp-per-block exact kernel: any block size, and the "redo" pass that
//                     resolves blocks whose winner may hinge on the float64 truncation (N2)
//
// Cost model of the reference (get_sad, full_search.py:10-17): SAD = trunc(sum_f64 |L1 - L2|)
// with L = 0.299R + 0.587G + 0.114B in float64.  With L1000 = 299R + 587G + 114B (exact int)
// and S = sum |L1000_1 - L1000_2|:  SAD = S / 1000 unless S % 1000 == 0 && S > 0, where the
// f64 sum may land just below the integer and SAD = S/1000 - 1 ("flagged" candidates).
#include <stdlib.h>

#include "memci_internal.cuh"

#define FS_NT 512
#define FS_BOXW 256
#define FS_NBMAX 32
#define KEY_NONE 0xffffffffffffffffull

// ---- candidate window of one block: full_search.py:33-43 incl. the H-for-W bug at :37 ----
struct FsBounds { int r0, r1, c0, c1; };
__host__ __device__ inline FsBounds fs_bounds(int H, int W, int bs, int R, int s_row, int s_col)
{
    FsBounds b;
    int tmr = (s_row + bs >= H) ? s_row + 1 : H - bs;
    int tmc = (s_col + bs >= H) ? s_col + 1 : W - bs;   // sic: compares against H
    b.r0 = s_row - R > 0 ? s_row - R : 0;
    b.r1 = tmr < s_row + R + 1 ? tmr : s_row + R + 1;
    b.c0 = s_col - R > 0 ? s_col - R : 0;
    b.c1 = tmc < s_col + R + 1 ? tmc : s_col + R + 1;
    return b;
}

// total order of candidates = the reference's update rule (:47): lower SAD, then smaller
// distance, then earlier in (t_row, t_col) scan order.
__device__ __forceinline__ unsigned long long fs_key(unsigned sad, unsigned d2, unsigned ri, unsigned ci)
{
    return ((unsigned long long)sad << 32) | ((unsigned long long)d2 << 16) | (ri << 8) | ci;
}

// =========================================================================================
// Generic exact kernel: one warp per block, lanes stride over candidates.
// =========================================================================================
__global__ void fs_generic_kernel(const int32_t *__restrict__ L0, const int32_t *__restrict__ L1,
                                  const uint8_t *__restrict__ rgb0, const uint8_t *__restrict__ rgb1,
                                  int H, int W, int Wp, int bs, int R, int blk_begin, int blk_end, int nbc,
                                  const int *__restrict__ list, const int *__restrict__ list_count,
                                  float2 *__restrict__ out_vec0, float *__restrict__ out_sad0,
                                  float2 *__restrict__ out_vec1, float *__restrict__ out_sad1)
{
    int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    int total = list ? *list_count : blk_end - blk_begin;
    for (int it = warp; it < total; it += nwarps) {
        const int entry = list ? list[it] : blk_begin + it;
        const int dir = (entry & 0x40000000) ? 1 : 0, blk = entry & 0x3fffffff;     // FS_DIR_BIT (redo list of a pair launch)
        const int32_t *Ls = dir ? L1 : L0, *Lt = dir ? L0 : L1;
        const uint8_t *rgb_s = dir ? rgb1 : rgb0, *rgb_t = dir ? rgb0 : rgb1;
        float2 *out_vec = dir ? out_vec1 : out_vec0;
        float *out_sad = dir ? out_sad1 : out_sad0;
        int br = blk / nbc, bc = blk - br * nbc;
        int s_row = br * bs, s_col = bc * bs;
        FsBounds b = fs_bounds(H, W, bs, R, s_row, s_col);
        int nr = b.r1 - b.r0, nc = b.c1 - b.c0;
        unsigned long long best = KEY_NONE;
        if (nr > 0 && nc > 0) {
            int ncand = nr * nc;
            for (int c = lane; c < ncand; c += 32) {
                int ri = c / nc, ci = c - ri * nc;
                int t_row = b.r0 + ri, t_col = b.c0 + ci;
                unsigned S = sad1000_global(Ls, Lt, H, W, Wp, s_row, s_col, t_row, t_col, bs);
                unsigned sad = S / 1000u;
                if (S != 0 && sad * 1000u == S)   // flagged: replay in f64, reference op order
                    sad = (unsigned)exact_cost_f64(rgb_s, rgb_t, H, W, s_row, s_col, t_row, t_col, bs, 0);
                int dr = s_row - t_row, dc = s_col - t_col;
                unsigned long long k = fs_key(sad, (unsigned)(dr * dr + dc * dc), ri, ci);
                best = k < best ? k : best;
            }
        }
        best = shfl_min_u64(best);
        if (lane == 0) {
            float dy, dx, sad;
            if (best == KEY_NONE) {          // empty window: target_index stays (0,0), sad = inf
                dy = (float)(-s_row); dx = (float)(-s_col); sad = __int_as_float(0x7f800000);
            } else {
                int ri = (int)((best >> 8) & 255), ci = (int)(best & 255);
                dy = (float)(b.r0 + ri - s_row); dx = (float)(b.c0 + ci - s_col);
                sad = (float)(unsigned)(best >> 32);
            }
            out_vec[blk] = make_float2(dy, dx);
            out_sad[blk] = sad;
        }
    }
}

// =========================================================================================
// Redo pass: one CTA per listed block, same exact arithmetic as fs_generic_kernel.  Block and search window are staged
// in shared memory as luma1000 ints (zero padded like np.pad).  A thread takes (column offset, FS_REDO_G consecutive row
// offsets): it keeps an 8 x 8 source sub-block in 64 registers and walks the FS_REDO_G + 7 window rows those candidates
// share, 8 LDS per row, neighbouring threads = neighbouring columns (conflict-free) -- a third of a shared-memory load per
// pixel pair (the first version read both operands from shared memory for every pair, staged a float64 copy of the
// whole window for the rare flagged candidate, and spent 20 us on a thousand 8 x 8 blocks, 350 us on as many 16 x 16 +-32
// blocks).  Flagged candidates (sum a positive multiple of 1000) are marked in a bitmap; only if there are any, the
// float64 luma of block and window is staged as well and every marked candidate is replayed by one thread in the
// reference's order.  (Flat grey content flags EVERY candidate of a block -- all sums are multiples of 1000 there -- so the
// replays must run thread-parallel out of shared memory; textured blocks have about one flagged candidate in a thousand,
// and up to FS_REDO_FLAGS of them are replayed straight from the RGB bytes, a warp each, without staging anything.)
// =========================================================================================
// List entries carry the search direction in bit 30 (pair launches: direction 1 swaps the frames).
#define FS_DIR_BIT 0x40000000
#define FS_REDO_NT 256
#define FS_REDO_G 4
#define FS_REDO_FLAGS 16           // up to this many flagged candidates per block are replayed straight from the RGB bytes, a warp each
static inline int fs_redo_smem_bytes(int bs, int R)
{
    const int ww = bs + 2 * R;
    return (bs * bs + (ww + FS_REDO_G) * ww) * 4 + 8 + (bs * bs + ww * ww) * 8 + ((ww * ww + 31) / 32) * 4;
}
__global__ void __launch_bounds__(FS_REDO_NT, 2)
fs_redo_kernel(const int32_t *__restrict__ L0, const int32_t *__restrict__ L1,
               const uint8_t *__restrict__ rgb0, const uint8_t *__restrict__ rgb1,
               int H, int W, int Wp, int bs, int R, int nbc,
               const int *__restrict__ list, const int *__restrict__ list_count,
               float2 *__restrict__ out_vec0, float *__restrict__ out_sad0,
               float2 *__restrict__ out_vec1, float *__restrict__ out_sad1)
{
    extern __shared__ __align__(16) int redo_sm[];
    __shared__ unsigned long long wbest[FS_REDO_NT / 32];
    __shared__ int n_flag_s;
    __shared__ unsigned short flag_s[FS_REDO_FLAGS];
    const int ww = bs + 2 * R, m = bs >> 3;
    int *ref = redo_sm, *win = ref + bs * bs;
    double *dref = reinterpret_cast<double *>(redo_sm + (((bs * bs + (ww + FS_REDO_G) * ww) + 1) & ~1)), *dwin = dref + bs * bs;
    unsigned *bits = reinterpret_cast<unsigned *>(dwin + ww * ww);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = *list_count;
    for (int it = blockIdx.x; it < total; it += gridDim.x) {
        const int entry = list[it];
        const int dir = (entry & FS_DIR_BIT) ? 1 : 0, blk = entry & (FS_DIR_BIT - 1);
        const int32_t *Ls = dir ? L1 : L0, *Lt = dir ? L0 : L1;
        const uint8_t *rgb_s = dir ? rgb1 : rgb0, *rgb_t = dir ? rgb0 : rgb1;
        float2 *out_vec = dir ? out_vec1 : out_vec0;
        float *out_sad = dir ? out_sad1 : out_sad0;
        const int br = blk / nbc, bc = blk - br * nbc;
        const int s_row = br * bs, s_col = bc * bs;
        const FsBounds b = fs_bounds(H, W, bs, R, s_row, s_col);
        const int nr = b.r1 - b.r0, nc = b.c1 - b.c0;
        for (int i = threadIdx.x; i < bs * bs; i += blockDim.x) {
            const int r = s_row + i / bs, c = s_col + i % bs;
            ref[i] = (r < H && c < W) ? Ls[(size_t)r * Wp + c] : 0;
        }
        for (int i = threadIdx.x; i < (ww + FS_REDO_G) * ww; i += blockDim.x) {
            const int r = s_row - R + i / ww, c = s_col - R + i % ww;
            win[i] = (r >= 0 && r < H && c >= 0 && c < W) ? Lt[(size_t)r * Wp + c] : 0;
        }
        const int ncand = nr > 0 && nc > 0 ? nr * nc : 0;
        for (int i = threadIdx.x; i < (ncand + 31) / 32; i += blockDim.x) bits[i] = 0u;
        if (threadIdx.x == 0) n_flag_s = 0;
        __syncthreads();
        unsigned long long best = KEY_NONE;
        const int n_grp = (nr + FS_REDO_G - 1) / FS_REDO_G;
        const int n_items = nr > 0 && nc > 0 ? n_grp * nc : 0;
        for (int item = threadIdx.x; item < n_items; item += blockDim.x) {
            const int gi = item / nc, ci = item - gi * nc, ri0 = gi * FS_REDO_G;
            const int *w0 = win + (b.r0 - (s_row - R) + ri0) * ww + (b.c0 - (s_col - R) + ci);
            unsigned S[FS_REDO_G];
#pragma unroll
            for (int g = 0; g < FS_REDO_G; ++g) S[g] = 0u;
            for (int sub = 0; sub < m * m; ++sub) {
                const int si = sub / m, sj = sub - si * m;
                int rf[64];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int4 a0 = *reinterpret_cast<const int4 *>(ref + (si * 8 + i) * bs + sj * 8);
                    const int4 a1 = *reinterpret_cast<const int4 *>(ref + (si * 8 + i) * bs + sj * 8 + 4);
                    rf[i * 8 + 0] = a0.x; rf[i * 8 + 1] = a0.y; rf[i * 8 + 2] = a0.z; rf[i * 8 + 3] = a0.w;
                    rf[i * 8 + 4] = a1.x; rf[i * 8 + 5] = a1.y; rf[i * 8 + 6] = a1.z; rf[i * 8 + 7] = a1.w;
                }
                const int *wp = w0 + (si * 8) * ww + sj * 8;
#pragma unroll
                for (int wr = 0; wr < FS_REDO_G + 7; ++wr) {
                    int wv[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) wv[q] = wp[wr * ww + q];
#pragma unroll
                    for (int g = 0; g < FS_REDO_G; ++g) {
                        const int i = wr - g;
                        if (i >= 0 && i < 8) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) S[g] = __sad(rf[i * 8 + q], wv[q], S[g]);
                        }
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < FS_REDO_G; ++g) {
                const int ri = ri0 + g;
                if (ri >= nr) continue;
                const int t_row = b.r0 + ri, t_col = b.c0 + ci;
                unsigned sad = S[g] / 1000u;
                if (S[g] != 0 && sad * 1000u == S[g]) {
                    const int c = ri * nc + ci;                      // flagged: decided by the float64 replay below
                    atomicOr(&bits[c >> 5], 1u << (c & 31));
                    const int fp = atomicAdd(&n_flag_s, 1);
                    if (fp < FS_REDO_FLAGS) flag_s[fp] = (unsigned short)((ri << 8) | ci);
                    continue;
                }
                const int dr = s_row - t_row, dc = s_col - t_col;
                const unsigned long long k = fs_key(sad, (unsigned)(dr * dr + dc * dc), ri, ci);
                best = k < best ? k : best;
            }
        }
        __syncthreads();
        const int nf = n_flag_s;                          // CTA-uniform
        if (nf > 0 && nf <= FS_REDO_FLAGS && (FS_REDO_NT / 32) * bs * bs <= ww * ww) {
            // a few flagged candidates: a warp each -- lanes form the |dL| terms side by side, lane 0 adds them up in the
            // reference's order (full_search.py:17); the term buffers live where the float64 window would
            for (int f = warp; f < nf; f += FS_REDO_NT / 32) {
                const int ri = flag_s[f] >> 8, ci = flag_s[f] & 255;
                const int t_row = b.r0 + ri, t_col = b.c0 + ci;
                double *tb = dwin + (size_t)warp * bs * bs;
                for (int px = lane; px < bs * bs; px += 32) {
                    const int i = px / bs, j = px - i * bs;
                    const int r1 = s_row + i, c1 = s_col + j, r2 = t_row + i, c2 = t_col + j;
                    double l1 = 0.0, l2 = 0.0;
                    if (r1 < H && c1 < W) { const uint8_t *q = rgb_s + ((size_t)r1 * W + c1) * 3; l1 = luma_f64(q[0], q[1], q[2]); }
                    if (r2 < H && c2 < W) { const uint8_t *q = rgb_t + ((size_t)r2 * W + c2) * 3; l2 = luma_f64(q[0], q[1], q[2]); }
                    tb[px] = fabs(__dsub_rn(l1, l2));
                }
                __syncwarp();
                if (lane == 0) {
                    double acc = 0.0;
                    for (int px = 0; px < bs * bs; ++px) acc = __dadd_rn(acc, tb[px]);
                    const int dr = s_row - t_row, dc = s_col - t_col;
                    const unsigned long long k = fs_key((unsigned)acc, (unsigned)(dr * dr + dc * dc), ri, ci);
                    best = k < best ? k : best;
                }
                __syncwarp();
            }
        } else if (nf > 0) {                              // many (flat content flags every candidate): float64 copy of block and window
            for (int i = threadIdx.x; i < bs * bs; i += blockDim.x) {
                const int r = s_row + i / bs, c = s_col + i % bs;
                double d = 0.0;
                if (r < H && c < W) { const uint8_t *q = rgb_s + ((size_t)r * W + c) * 3; d = luma_f64(q[0], q[1], q[2]); }
                dref[i] = d;
            }
            for (int i = threadIdx.x; i < ww * ww; i += blockDim.x) {
                const int r = s_row - R + i / ww, c = s_col - R + i % ww;
                double d = 0.0;
                if (r >= 0 && r < H && c >= 0 && c < W) { const uint8_t *q = rgb_t + ((size_t)r * W + c) * 3; d = luma_f64(q[0], q[1], q[2]); }
                dwin[i] = d;
            }
            __syncthreads();
            for (int c = threadIdx.x; c < ncand; c += blockDim.x) {
                if (!(bits[c >> 5] & (1u << (c & 31)))) continue;
                const int ri = c / nc, ci = c - ri * nc;
                const int t_row = b.r0 + ri, t_col = b.c0 + ci;
                const int woff = (t_row - (s_row - R)) * ww + (t_col - (s_col - R));
                // sequential f64 accumulation in the reference's order (full_search.py:17)
                double acc = 0.0;
                for (int i = 0; i < bs; ++i)
                    for (int j = 0; j < bs; ++j)
                        acc = __dadd_rn(acc, fabs(__dsub_rn(dref[i * bs + j], dwin[woff + i * ww + j])));
                const int dr = s_row - t_row, dc = s_col - t_col;
                const unsigned long long k = fs_key((unsigned)acc, (unsigned)(dr * dr + dc * dc), ri, ci);
                best = k < best ? k : best;
            }
        }
        best = shfl_min_u64(best);
        if (lane == 0) wbest[warp] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < FS_REDO_NT / 32; ++w) best = wbest[w] < best ? wbest[w] : best;
            if (best != KEY_NONE) {
                const int ri = (int)((best >> 8) & 255), ci = (int)(best & 255);
                out_vec[blk] = make_float2((float)(b.r0 + ri - s_row), (float)(b.c0 + ci - s_col));
                out_sad[blk] = (float)(unsigned)(best >> 32);
            }
        }
        __syncthreads();
    }
}

// =========================================================================================
// Tiled kernel.
//
// A tile = nb horizontally adjacent blocks of one block row.  The lead warp stages, with two TMA
// tile loads (cp.async.bulk.tensor.2d, zero fill outside the frame = the reference's np.pad),
//   win[box_h][256]  luma1000 of the target frame, rows s_row-R .. s_row+bs+R-1, cols x0-R ..
//   ref[bs][256]     luma1000 of the source frame, the nb blocks themselves
// into one of up to three shared-memory stages while earlier tiles are being searched.
//
// Work item = (block b, column offset dx, group g of G consecutive row offsets).  A thread
// keeps the 8x8 reference sub-block in 64 registers and G accumulators, walks the G+7 window
// rows the group touches (8 conflict-free LDS.32 per row: neighbouring lanes = neighbouring dx)
// and issues one VABSDIFF (|a-b|+acc) per pixel-candidate.  Blocks larger than 8 are searched
// as (bs/8)^2 sub-blocks accumulating into the same registers.  Items are dense over the
// *valid* candidates (frame-edge clipping and the :37 bug remove 23 % of them at 1080p), and
// the host packs tiles to ~3 items per thread so no lane idles.
//
// Argmin: per item in registers, then one shared-memory atomicMin per item on the 64-bit key.
// Flagged candidates are tracked with SAD-1 as a lower bound; a block whose best lower bound
// beats its best fast key is appended to the redo list and re-resolved by fs_redo_kernel.
// =========================================================================================
struct FsParams {
    int H, W, bs, R, Rp, m, nbr, nbc, n_tiles, box_h, n_stages, ks;   // ks = bits reserved for dist^2 in the 32-bit key   // Rp = R rounded up to 4: TMA x origin stays 16-byte aligned
    int lead_warp;               // the warp that writes the tile's vectors and issues the TMA loads
    unsigned long long *trace;   // debug (MEMCI_FS_TRACE=file): per (tile, warp) of CTA 0: clock at wait / start / end of the column
    unsigned one;                // the constant 1, opaque to ptxas (keeps IMADs from being folded into ALU adds)
    int n_tiles_dir;             // tiles of one direction; tile index >= n_tiles_dir = reverse direction (pair launch)
    const FsTile *tiles;
    float2 *out_vec[2];
    float *out_sad[2];
    int *redo_list;
    int *counters;
    int redo_cap;
    const int *gate;             // non-null: run only if *gate != 0 (the prefilter path's survivor list overflowed)
};

struct FsTileTab {
    int by, bx0, nb, s_row, dir;
    int r0, n_dy, total_dx, uniform_ndx;   // uniform_ndx > 0: every block has that many valid column offsets
    int dx_lo[FS_NBMAX];
    int pre[FS_NBMAX + 1];
};

// geometry of one tile, computed by warp 0 (lane = block within tile)
__device__ inline void fs_fill_tab(FsTileTab *tab, const FsTile t, const FsParams &p, int lane, int dir)
{
    int s_row = t.by * p.bs;
    int n_dx = 0, dx_lo = 0;
    if (lane < t.nb) {
        int s_col = (t.bx0 + lane) * p.bs;
        FsBounds b = fs_bounds(p.H, p.W, p.bs, p.R, s_row, s_col);
        n_dx = b.c1 - b.c0 > 0 ? b.c1 - b.c0 : 0;
        dx_lo = b.c0 - s_col;
    }
    int incl = n_dx;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane < FS_NBMAX) { tab->dx_lo[lane] = dx_lo; tab->pre[lane + 1] = incl; }
    if (lane == 0) {
        FsBounds b = fs_bounds(p.H, p.W, p.bs, p.R, s_row, t.bx0 * p.bs);
        tab->pre[0] = 0;
        tab->by = t.by; tab->bx0 = t.bx0; tab->nb = t.nb; tab->s_row = s_row; tab->dir = dir;
        tab->r0 = b.r0;
        tab->n_dy = b.r1 - b.r0 > 0 ? b.r1 - b.r0 : 0;
    }
    int total = __shfl_sync(0xffffffffu, incl, 31);
    const int n0 = __shfl_sync(0xffffffffu, n_dx, 0);
    const bool uni = __all_sync(0xffffffffu, lane >= t.nb || n_dx == n0);
    if (lane == 0) { tab->total_dx = total; tab->uniform_ndx = (uni && n0 > 0) ? n0 : 0; }
}

// Per-thread argmin over the G row offsets of one group.  dx is fixed and the row offset ascends
// with gi, so the 32-bit key  ((sad << ks) + d2 + 1) << GB | gi   (GB = 4 bits for G <= 16, else 6)  under plain unsigned min
// reproduces the reference's update rule (lower SAD, then smaller distance, then earlier row).
// kl is the same minimum with every flagged candidate (S % 1000 == 0) lowered to sad-1 and one
// notch further, so that "kl < best key" also covers a flagged candidate that would TIE the
// winner on (sad-1, d2) and could beat it on scan order.  S == 0 also has remainder 0; its
// lowered key wraps to a huge value and is ignored by the min.
//
// The kernel is bound by the integer-ALU pipe (VABSDIFF, 2 issue cycles per warp instruction), while the
// FMA pipe idles.  So the whole per-candidate bookkeeping is written as IMAD / IMAD.HI (inline PTX keeps
// ptxas from turning them back into ALU-pipe shifts, adds and selects); only the running minima (one
// VIMNMX3 per two candidates and chain) stay on the ALU pipe.  Per candidate, with Sm1 = S - 1 (the SAD
// accumulators start at 0xffffffff so that S - 1 comes for free):
//   S    = Sm1 * 1 + 1                      hi = mulhi(S, 0x10624dd3)     q = mulhi(hi, 2^26) = S / 1000
//   rem1 = q * -1000 + Sm1 = (S % 1000) - 1  flag = mulhi(rem1, 2) = (S % 1000 == 0)
//   u    = two_dy0 * (GM gi) + base          t = one * (GM gi^2 + gi) + u   (one = 1 from the parameter block:
//                                                                            a literal 1 would be folded into an add)
//   key  = q * mulg + t                      lo = flag * -lower + key
template <int G, bool CHECK>
__device__ __forceinline__ void fs_finalize(const unsigned (&acc)[G], int n_valid, int dy0, int dx, unsigned mulg, unsigned one,
                                            unsigned &kb, unsigned &kl)
{
    constexpr unsigned GM = G <= 16 ? 16u : 64u;
    const unsigned base = ((unsigned)(dy0 * dy0 + dx * dx) + 1u) * GM;
    const unsigned a_dy = (unsigned)(2 * dy0) * GM;
    const unsigned neg_lower = 0u - (mulg + GM);              // one SAD step + one notch, in key units
    kb = 0xffffffffu; kl = 0xffffffffu;
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
        const unsigned Sm1 = acc[gi];
        unsigned S, hi, q, rem1, flag, u, t, key, lo;
        asm("mad.lo.u32 %0, %1, %2, 1;" : "=r"(S) : "r"(Sm1), "r"(one));
        asm("mul.hi.u32 %0, %1, 0x10624dd3;" : "=r"(hi) : "r"(S));
        asm("mul.hi.u32 %0, %1, 0x4000000;" : "=r"(q) : "r"(hi));
        asm("mad.lo.u32 %0, %1, 0xfffffc18, %2;" : "=r"(rem1) : "r"(q), "r"(Sm1));
        asm("mul.hi.u32 %0, %1, 2;" : "=r"(flag) : "r"(rem1));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(u) : "r"(a_dy), "r"((unsigned)gi), "r"(base));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(one), "r"(GM * (unsigned)(gi * gi) + (unsigned)gi), "r"(u));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(key) : "r"(q), "r"(mulg), "r"(t));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(lo) : "r"(flag), "r"(neg_lower), "r"(key));
        if (CHECK) { if (gi >= n_valid) { key = 0xffffffffu; lo = 0xffffffffu; } }   // rows past the window lose
        kb = min(kb, key);
        kl = min(kl, lo);
    }
}

// Work item = (block b, column offset dx): one thread walks all row-offset groups of its
// candidate column, so the per-item bookkeeping (decode, 8x8 reference load when bs == 8,
// warp reduction, shared-memory atomic) is paid once per 8*8*(2R+1) SAD pixel-ops.
//
// Synchronisation is mbarrier-only: full[s] = "tile data landed in stage s" (TMA complete_tx),
// done[j] = "all 16 warps finished searching the tile using result buffer j".  Only the lead warp
// ever waits on done[] (it writes the tile's motion vectors, re-arms the stage and issues the
// next TMA), so a warp that finishes early moves straight on to the next tile.
template <int G, bool SINGLE>
__global__ void __launch_bounds__(FS_NT, 1)
fs_tiled_kernel(const __grid_constant__ CUtensorMap tm_src, const __grid_constant__ CUtensorMap tm_tgt,
                const __grid_constant__ CUtensorMap tm_src1, const __grid_constant__ CUtensorMap tm_tgt1,
                const FsParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[3];
    __shared__ __align__(8) uint64_t done[3];
    __shared__ FsTileTab tabs[3];
    __shared__ unsigned long long best_fast[3][FS_NBMAX];
    __shared__ unsigned best_low[3][FS_NBMAX];

    if (p.gate && *p.gate == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int stage_ints = (p.box_h + p.bs) * FS_BOXW;
    int *const stage0 = reinterpret_cast<int *>(smem_raw);
    const unsigned stage_bytes = (unsigned)stage_ints * 4u;
    const int m = SINGLE ? 1 : p.m;
    constexpr int GB = G <= 16 ? 4 : 6;
    const unsigned mulg = (1u << GB) << p.ks;

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(&full[i], 1); mbar_init(&done[i], FS_NT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 3 * FS_NBMAX) {
        (&best_fast[0][0])[tid] = KEY_NONE;
        (&best_low[0][0])[tid] = 0xffffffffu;
    }
    __syncthreads();

    const int first = blockIdx.x, stride = gridDim.x;
    if (first >= p.n_tiles) return;

    // Ring indices: tile k lives in data stage k % NS and in table / result buffer k % NB.
    //   NS == 3 (the 8x8 +-16 case: 3 x 48 KB): warp 0 does its housekeeping AFTER its own share of tile k --
    //     by then done[k-1] completed long ago, so it never spins -- writes tile k-1's vectors and issues the
    //     TMA for tile k+2 into the stage tile k-1 just vacated.  Two tiles are always in flight or landed.
    //   NS == 2: warp 0 waits for done[k-1] BEFORE tile k (it needs that stage for tile k+1).
    //   NS == 1: load, search, output, strictly in turn.
    const int NS = p.n_stages, NB = NS == 3 ? 3 : 2;
    auto issue = [&](int tile_idx, int s, int j) {        // lead warp only: tile table + two TMA tile loads
        const int dir = tile_idx >= p.n_tiles_dir ? 1 : 0;
        FsTile t = p.tiles[tile_idx - dir * p.n_tiles_dir];
        fs_fill_tab(&tabs[j], t, p, lane, dir);
        __syncwarp();
        if (lane == 0) {
            int *win = stage0 + (size_t)s * stage_ints;
            int *ref = win + p.box_h * FS_BOXW;
            mbar_expect_tx(&full[s], stage_bytes);
            tma_load_2d(win, dir ? &tm_tgt1 : &tm_tgt, t.bx0 * p.bs - p.Rp, t.by * p.bs - p.R, &full[s]);
            tma_load_2d(ref, dir ? &tm_src1 : &tm_src, t.bx0 * p.bs, t.by * p.bs, &full[s]);
        }
    };

    // warp 0, lanes < nb: motion vectors of the tile in buffer j out, result buffers re-armed, redo list fed
    auto output = [&](int j) {
        const FsTileTab *tab = &tabs[j];
        if (lane < tab->nb) {
            const unsigned long long kb = best_fast[j][lane];
            const unsigned kl = best_low[j][lane];
            best_fast[j][lane] = KEY_NONE;
            best_low[j][lane] = 0xffffffffu;
            const int bc = tab->bx0 + lane;
            const int blk = tab->by * p.nbc + bc;
            const int s_row = tab->s_row, s_col = bc * p.bs;
            float dy, dx, sad;
            if (kb == KEY_NONE) {                       // empty window: target_index stays (0,0), sad = inf
                dy = (float)(-s_row); dx = (float)(-s_col); sad = __int_as_float(0x7f800000);
            } else {
                const int ri = (int)((kb >> 8) & 255), ci = (int)(kb & 255);
                dy = (float)(tab->r0 - s_row + ri);
                dx = (float)(tab->dx_lo[lane] + ci);
                sad = (float)(unsigned)(kb >> 32);
            }
            (tab->dir ? p.out_vec[1] : p.out_vec[0])[blk] = make_float2(dy, dx);     // no dynamic index into the param struct
            (tab->dir ? p.out_sad[1] : p.out_sad[0])[blk] = sad;
            // lowered key of some flagged candidate reaches the winner's (sad, d2): exact redo
            if (kb != KEY_NONE && kl < ((unsigned)(kb >> 32) << p.ks) + (unsigned)((kb >> 16) & 0xffffu) + 1u) {
                int pos = atomicAdd(&p.counters[0], 1);
                if (pos < p.redo_cap) p.redo_list[pos] = blk | (tab->dir ? FS_DIR_BIT : 0);
            }
        }
        __syncwarp();
    };

    if (warp == p.lead_warp) {
        issue(first, 0, 0);
        if (NS == 3 && first + stride < p.n_tiles) issue(first + stride, 1, 1);
    }

    int k = 0;
    int s = 0, j = 0;                                  // k % NS, k % NB
    unsigned full_phase[3] = {0u, 0u, 0u}, done_phase[3] = {0u, 0u, 0u};
    for (int tile_idx = first; tile_idx < p.n_tiles; tile_idx += stride, ++k) {
        const int next = tile_idx + stride;
        const int jp = j == 0 ? NB - 1 : j - 1;        // buffer of tile k-1
        const int jn = j + 1 == NB ? 0 : j + 1;        // buffer of tile k+1
        if (warp == p.lead_warp && NS == 2) {
            if (k >= 1) {
                mbar_wait(&done[jp], done_phase[jp]);
                done_phase[jp] ^= 1u;
                output(jp);
            }
            if (next < p.n_tiles) issue(next, s ^ 1, jn);     // streams in while this tile is searched
        }

        unsigned long long t_wait = 0, t_start = 0;
        if (p.trace && blockIdx.x == 0 && lane == 0) t_wait = clock64();
        mbar_wait(&full[s], full_phase[s]);
        full_phase[s] ^= 1u;
        if (p.trace && blockIdx.x == 0 && lane == 0) t_start = clock64();
        const FsTileTab *tab = &tabs[j];
        const int *win = stage0 + (size_t)s * stage_ints;
        const int *ref = win + p.box_h * FS_BOXW;
        const int nb = tab->nb, n_items = tab->total_dx, n_dy = tab->n_dy, ndx0 = tab->uniform_ndx;
        const int n_groups = (n_dy + G - 1) / G;
        const int s_row = tab->s_row;
        const int dy_lo = tab->r0 - s_row;               // first valid row offset
        const int wrow_base = tab->r0 - (s_row - p.R);   // its row inside win[]

        for (int base = warp * 32; base < n_items; base += FS_NT) {
            const int item = base + lane;
            const bool active = item < n_items;
            unsigned long long kbest = KEY_NONE;
            unsigned klow = 0xffffffffu;
            int b = 0;
            if (active) {
                if (ndx0 > 0) b = item / ndx0;
                else while (b + 1 < nb && tab->pre[b + 1] <= item) ++b;
                const int dxi = item - tab->pre[b];
                const int dx = tab->dx_lo[b] + dxi;
                const int wc = b * p.bs + dx + p.Rp;
                const int rc = b * p.bs;
                int r[8][8];
                if (SINGLE) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        int4 a = *reinterpret_cast<const int4 *>(ref + kk * FS_BOXW + rc);
                        int4 c = *reinterpret_cast<const int4 *>(ref + kk * FS_BOXW + rc + 4);
                        r[kk][0] = a.x; r[kk][1] = a.y; r[kk][2] = a.z; r[kk][3] = a.w;
                        r[kk][4] = c.x; r[kk][5] = c.y; r[kk][6] = c.z; r[kk][7] = c.w;
                    }
                }
                unsigned kb_run = 0xffffffffu;
                int g_run = 0;
#pragma unroll 1
                for (int g = 0; g < n_groups; ++g) {
                    const int wr0 = wrow_base + g * G;
                    unsigned acc[G];
#pragma unroll
                    for (int i = 0; i < G; ++i) acc[i] = 0xffffffffu;      // S - 1: see fs_finalize
#pragma unroll 1
                    for (int sub = 0; sub < m * m; ++sub) {
                        const int si = SINGLE ? 0 : sub / m, sj = SINGLE ? 0 : sub - si * m;
                        if (!SINGLE) {
                            const int *rp = ref + (si * 8) * FS_BOXW + rc + sj * 8;
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk) {
                                int4 a = *reinterpret_cast<const int4 *>(rp + kk * FS_BOXW);
                                int4 c = *reinterpret_cast<const int4 *>(rp + kk * FS_BOXW + 4);
                                r[kk][0] = a.x; r[kk][1] = a.y; r[kk][2] = a.z; r[kk][3] = a.w;
                                r[kk][4] = c.x; r[kk][5] = c.y; r[kk][6] = c.z; r[kk][7] = c.w;
                            }
                        }
                        const int *wp = win + (wr0 + si * 8) * FS_BOXW + wc + sj * 8;
#pragma unroll
                        for (int w = 0; w < G + 7; ++w) {
                            int px[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) px[j] = wp[w * FS_BOXW + j];
                            // pixel-major order: consecutive VABSDIFFs feed different accumulators
                            // (independent chains), so one warp alone can keep the pipe busy
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
#pragma unroll
                                for (int gi = 0; gi < G; ++gi) {
                                    const int kk = w - gi;
                                    if (kk >= 0 && kk < 8) acc[gi] = __sad(r[kk][j], px[j], acc[gi]);
                                }
                            }
                        }
                    }
                    const int n_valid = n_dy - g * G;       // > 0
                    const int dy0 = dy_lo + g * G;
                    unsigned kb32, kl32;
                    if (n_valid >= G) fs_finalize<G, false>(acc, n_valid, dy0, dx, mulg, p.one, kb32, kl32);
                    else fs_finalize<G, true>(acc, n_valid, dy0, dx, mulg, p.one, kb32, kl32);
                    // groups ascend in row offset: strict '<' on (sad, d2) keeps the earlier row on ties
                    if ((kb32 >> GB) < (kb_run >> GB)) { kb_run = kb32; g_run = g; }
                    klow = min(klow, kl32);
                }
                if (kb_run != 0xffffffffu) {
                    const unsigned kq = (kb_run >> GB) - 1u;
                    kbest = fs_key(kq >> p.ks, kq & ((1u << p.ks) - 1u), (unsigned)(g_run * G) + (kb_run & ((1u << GB) - 1u)), (unsigned)dxi);
                }
            }
            // Segmented warp reduction: the 32 items of a warp belong to a few adjacent blocks (usually
            // two); one REDUX-based 64-bit min + two shared-memory atomics per (warp, block) instead of
            // per-lane CAS loops on the same address.
            unsigned todo = __ballot_sync(0xffffffffu, active);
            while (todo) {
                const int leader = __ffs(todo) - 1;
                const int bb = __shfl_sync(0xffffffffu, b, leader);
                const bool mine = active && b == bb;
                const unsigned hi = mine ? (unsigned)(kbest >> 32) : 0xffffffffu;
                const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
                const unsigned lo = (mine && hi == mh) ? (unsigned)kbest : 0xffffffffu;
                const unsigned ml = __reduce_min_sync(0xffffffffu, lo);
                const unsigned kl_b = __reduce_min_sync(0xffffffffu, mine ? klow : 0xffffffffu);
                if (lane == leader) {
                    const unsigned long long kw = ((unsigned long long)mh << 32) | ml;
                    if (kw != KEY_NONE) atomicMin(&best_fast[j][bb], kw);
                    if (kl_b != 0xffffffffu) atomicMin(&best_low[j][bb], kl_b >> GB);
                }
                todo &= ~__ballot_sync(0xffffffffu, mine);
            }
        }
        __syncwarp();
        if (p.trace && blockIdx.x == 0 && lane == 0 && k < 64) {
            unsigned long long *t = p.trace + ((size_t)k * (FS_NT / 32) + warp) * 3;
            t[0] = t_wait; t[1] = t_start; t[2] = clock64();
        }
        if (lane == 0) mbar_arrive(&done[j]);

        if (warp == p.lead_warp && NS == 1) {
            mbar_wait(&done[j], done_phase[j]);
            done_phase[j] ^= 1u;
            output(j);
            if (next < p.n_tiles) issue(next, 0, jn);
        }
        if (warp == p.lead_warp && NS == 3) {
            if (k >= 1) {
                mbar_wait(&done[jp], done_phase[jp]);     // completed about one tile ago: no spin
                done_phase[jp] ^= 1u;
                output(jp);
            }
            // tile k+2 goes where tile k-1 was (stage and buffer (k+2) % 3 == (k-1) % 3)
            if (next + stride < p.n_tiles) issue(next + stride, jp, jp);
        }
        s = s + 1 == NS ? 0 : s + 1;
        j = jn;
    }
    if (warp == p.lead_warp && NS >= 2 && k >= 1) {
        const int jl = j == 0 ? NB - 1 : j - 1;          // buffer of the last tile
        mbar_wait(&done[jl], done_phase[jl]);
        output(jl);
    }
}

// =========================================================================================
// Host side: plan (tile table), tensor maps, launch.
// =========================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// Tiled tensor map, no swizzle / interleave, zero fill outside `dims`.  rank 2 or 3; strides[i] = bytes between
// consecutive indices of dimension i + 1 (multiples of 16); box[0] * element size must be a multiple of 16 bytes.
int memci_tmap_encode(memci_ctx *ctx, CUtensorMap *tm, int dtype_u8, int rank, const void *base,
                      const unsigned long long *dims, const unsigned long long *strides, const unsigned *box)
{
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return memci_fail(ctx, MEMCI_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t d[3], st[2];
    cuuint32_t bx[3], estr[3] = {1, 1, 1};
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) st[i] = strides[i];
    CUresult r = enc(tm, dtype_u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_INT32, (cuuint32_t)rank, const_cast<void *>(base), d, st, bx, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return memci_fail(ctx, MEMCI_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d", (int)r);
    return MEMCI_OK;
}

static int make_tmap(memci_ctx *ctx, CUtensorMap *tm, const int32_t *plane, int H, int W, int Wp, int box_w, int box_h)
{
    const unsigned long long dims[2] = {(unsigned long long)W, (unsigned long long)H};        // true extent: x >= W / y >= H zero-fill
    const unsigned long long strides[1] = {(unsigned long long)Wp * 4};
    const unsigned box[2] = {(unsigned)box_w, (unsigned)box_h};
    return memci_tmap_encode(ctx, tm, 0, 2, plane, dims, strides, box);
}

static int pick_G(int R)
{
    if (const char *e = getenv("MEMCI_FS_G")) {          // tuning override
        int g = atoi(e);
        if (g == 5 || g == 8 || g == 11 || g == 13 || g == 17 || g == 33) return g;
    }
    // group size dividing 2R+1 where possible (no half-empty last group)
    const int n = 2 * R + 1;
    const int cands[4] = {11, 13, 8, 5};
    int best = 11, best_waste = 1 << 30;
    for (int i = 0; i < 4; ++i) {
        int g = cands[i];
        int waste = ((n + g - 1) / g) * g - n;
        // relative waste, prefer larger G on ties (fewer window re-reads)
        int score = waste * 1000 / n * 16 + (16 - g);
        if (score < best_waste) { best_waste = score; best = g; }
    }
    return best;
}

static int fs_smem_bytes(int bs, int box_h, int n_stages, int G)
{
    return ((box_h + bs) * n_stages + G + 8) * FS_BOXW * 4;
}

static FsPlan *fs_get_plan(memci_ctx *ctx, int bs, int R)
{
    for (int i = 0; i < 4; ++i)
        if (ctx->fs_plan[i].valid && ctx->fs_plan[i].bs == bs && ctx->fs_plan[i].R == R) return &ctx->fs_plan[i];
    FsPlan *pl = nullptr;
    for (int i = 0; i < 4; ++i)
        if (!ctx->fs_plan[i].valid) { pl = &ctx->fs_plan[i]; break; }
    if (!pl) {                       // evict slot 0
        pl = &ctx->fs_plan[0];
        if (pl->d_tiles) cudaFree(pl->d_tiles);
        free(pl->row_tile_start); free(pl->row_cand);
        memset(pl, 0, sizeof(*pl));
    }
    const int H = ctx->H, W = ctx->W;
    const int nbr = ceil_div_i(H, bs), nbc = ceil_div_i(W, bs);
    pl->bs = bs; pl->R = R; pl->G = pick_G(R);
    pl->box_w = FS_BOXW; pl->box_h = bs + 2 * R;
    // exact candidate count -- same bounds as the kernels
    long long n_cand = 0;
    pl->row_cand = (long long *)calloc(nbr, sizeof(long long));
    pl->row_tile_start = (int *)calloc(nbr + 1, sizeof(int));
    for (int br = 0; br < nbr; ++br) {
        for (int bc = 0; bc < nbc; ++bc) {
            FsBounds b = fs_bounds(H, W, bs, R, br * bs, bc * bs);
            if (b.r1 > b.r0 && b.c1 > b.c0) pl->row_cand[br] += (long long)(b.r1 - b.r0) * (b.c1 - b.c0);
        }
        n_cand += pl->row_cand[br];
    }
    pl->n_cand = n_cand;
    pl->n_tiles = 0; pl->d_tiles = nullptr;
    int key_bits = 0;
    { int ks = 1; while ((1 << ks) <= 2 * R * R + 1) ++ks; int qb = 1; while ((1ll << qb) <= (long long)bs * bs * 255) ++qb; key_bits = ks + qb + (pl->G <= 16 ? 4 : 6); }
    const bool tiled_ok = !getenv("MEMCI_FS_FORCE_GENERIC") && (bs % 8 == 0) && bs <= 64 && R <= 120 && key_bits <= 32 && (bs + 2 * R + 4) <= FS_BOXW &&
                          fs_smem_bytes(bs, pl->box_h, 1, pl->G) <= 200 * 1024;
    if (tiled_ok) {
        const int Rp = (R + 3) & ~3;
        // one item (candidate column of one block) per thread per tile
        const int budget = FS_NT;
        FsTile *h = (FsTile *)malloc(sizeof(FsTile) * (size_t)nbr * nbc);
        int nt = 0;
        for (int br = 0; br < nbr; ++br) {
            pl->row_tile_start[br] = nt;
            int bc = 0;
            while (bc < nbc) {
                int nb = 0, items = 0;
                while (bc + nb < nbc && nb < FS_NBMAX) {
                    FsBounds b = fs_bounds(H, W, bs, R, br * bs, (bc + nb) * bs);
                    int n_dx = b.c1 - b.c0 > 0 ? b.c1 - b.c0 : 0;
                    int add = n_dx;
                    // rightmost window column this block reads: origin is x0 - Rp, last offset is c1 - 1
                    int right = Rp + nb * bs + (n_dx > 0 ? (b.c1 - 1 - (bc + nb) * bs) : 0) + bs;
                    if (right > FS_BOXW || (nb + 1) * bs > FS_BOXW) break;
                    if (nb > 0 && items + add > budget) break;
                    items += add; ++nb;
                }
                h[nt].by = br; h[nt].bx0 = bc; h[nt].nb = nb; h[nt].pad = items;
                ++nt; bc += nb;
            }
        }
        if (cudaMalloc(&pl->d_tiles, sizeof(FsTile) * nt) != cudaSuccess ||
            cudaMemcpyAsync(pl->d_tiles, h, sizeof(FsTile) * nt, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            free(h);
            memci_fail(ctx, MEMCI_ERR_CUDA, "fs plan upload failed");
            return nullptr;
        }
        free(h);
        pl->row_tile_start[nbr] = nt;
        pl->n_tiles = nt;
    }
    pl->valid = 1;
    return pl;
}

template <int G, bool SINGLE>
static int fs_launch_tiled2(memci_ctx *ctx, const CUtensorMap *tm, const FsParams &p, int smem)
{
    CK(ctx, cudaFuncSetAttribute(fs_tiled_kernel<G, SINGLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int grid = ctx->num_sms < p.n_tiles ? ctx->num_sms : p.n_tiles;
    fs_tiled_kernel<G, SINGLE><<<grid, FS_NT, smem, ctx->stream>>>(tm[0], tm[1], tm[2], tm[3], p);
    CKL(ctx);
    return MEMCI_OK;
}

template <int G>
static int fs_launch_tiled(memci_ctx *ctx, const CUtensorMap *tm, const FsParams &p, int smem)
{
    return p.m == 1 ? fs_launch_tiled2<G, true>(ctx, tm, p, smem) : fs_launch_tiled2<G, false>(ctx, tm, p, smem);
}

// out2 != nullptr: pair launch -- the same persistent kernel also searches the reverse direction
// (slot_tgt -> slot_src) into out2 (one launch, one redo pass for both motion fields).
int memci_fs_run(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int R, MVField *out, int br0, int br1, MVField *out2)
{
    const int H = ctx->H, W = ctx->W, Wp = ctx->Wp;
    if (bs <= 0 || bs > 64 || R < 0 || R > 127)
        return memci_fail(ctx, MEMCI_ERR_ARG, "full search: block_size %d / target_region %d out of range (1..64, 0..127)", bs, R);
    int rc;
    if ((rc = memci_ensure_luma(ctx, slot_src))) return rc;
    if ((rc = memci_ensure_luma(ctx, slot_tgt))) return rc;
    const int nbr = ceil_div_i(H, bs), nbc = ceil_div_i(W, bs);
    if ((rc = memci_mv_reserve(ctx, out, bs, nbr, nbc, bs, nbr, nbc))) return rc;
    if (out2 && (rc = memci_mv_reserve(ctx, out2, bs, nbr, nbc, bs, nbr, nbc))) return rc;
    const int ndir = out2 ? 2 : 1;
    FsPlan *pl = fs_get_plan(ctx, bs, R);
    if (!pl) return MEMCI_ERR_CUDA;
    if (br1 < 0) { br0 = 0; br1 = nbr; }
    if (br0 < 0 || br1 > nbr || br0 >= br1) return memci_fail(ctx, MEMCI_ERR_ARG, "block row range [%d, %d) outside [0, %d)", br0, br1, nbr);
    long long cand = 0;
    for (int r = br0; r < br1; ++r) cand += pl->row_cand[r];
    ctx->fs_last_cand = cand * ndir;
    ctx->fs_last_ops = cand * ndir * bs * bs;
    FrameSlot &fs = ctx->frames[slot_src], &ft = ctx->frames[slot_tgt];
    if ((size_t)ctx->redo_cap < 2 * (size_t)nbr * nbc) {
        if (ctx->d_redo_list) cudaFree(ctx->d_redo_list);
        ctx->d_redo_list = nullptr;
        CK(ctx, cudaMalloc(&ctx->d_redo_list, sizeof(int) * 2 * (size_t)nbr * nbc));
        ctx->redo_cap = 2 * nbr * nbc;
    }
    ctx->fs_timed = 0;
    // Two-phase exact search (fs8.cu): 8-bit prefilter + exact refinement of the survivors; the exact tiled kernel below
    // then runs only if the survivor list overflowed (device-side gate, no host round trip).
    const char *pf_env = getenv("MEMCI_FS_PREFILTER");
    bool use8 = pl->n_tiles > 0 && !(pf_env && atoi(pf_env) == 0) && memci_fs8_eligible(ctx, bs, R);
    // Content feedback: the prefilter only pays where it prunes.  The counters of a recent two-phase search come back
    // through an asynchronous copy; if they show an overflow, or survivors and redo blocks that cost more than the exact
    // kernel would have, the next 32 searches go straight to the exact kernel (then the prefilter is probed again).  MEMCI_FS_PREFILTER=1 pins it on.
    if (use8 && !(pf_env && atoi(pf_env) == 1)) {
        if (ctx->fs8_feedback_pending && cudaEventQuery(ctx->ev_fs8_feedback) == cudaSuccess) {
            const int *h = ctx->h_fs8_feedback;
            ctx->fs8_feedback_pending = 0;
            // Estimated GPU time of what that search did against the exact kernel's, from measured rates (1080p 8 x 8 +-16 and
            // 4K 16 x 16 +-32, profiles/): exact 12.9 T pixel-ops/s, prefilter 36 T, refine 0.15 + 0.001 bs^2 ns per
            // survivor, redo pass 4 T pixel-ops/s.  An overflow (h[0]) means the exact kernel ran on top of everything.
            const double ops = (double)ctx->fs8_feedback_cand * bs * bs;
            const double per_block = ops / ((double)nbr * nbc * ndir);
            const double est_us = ops / 36e6 + h[1] * (0.15 + 0.001 * bs * bs) * 1e-3 + h[3] * per_block / 4e6;
            if (h[0] != 0 || est_us > ops / 12.9e6) ctx->fs8_backoff = 32;
        }
        if (ctx->fs8_backoff > 0) { --ctx->fs8_backoff; use8 = false; }
    }
    int prof = -1;
    int *redo_count = use8 ? ctx->d_counters + 11 : ctx->d_counters;     // the prefilter path zeroes [8..11] in one memset
    ctx->fs_used_prefilter = use8 ? 1 : 0;
    if (!use8) CK(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(int), ctx->stream));
    if (use8) {
        if (ctx->timing) CK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
        if ((rc = memci_fs8_run(ctx, slot_src, slot_tgt, bs, R, out, br0, br1, out2))) return rc;   // per-launch profile events: around the prefilter kernel
    }
    if (pl->n_tiles > 0) {
        CUtensorMap tm[4];                     // src/tgt of direction 0, src/tgt of direction 1
        if ((rc = make_tmap(ctx, &tm[0], fs.luma, H, W, Wp, FS_BOXW, bs))) return rc;
        if ((rc = make_tmap(ctx, &tm[1], ft.luma, H, W, Wp, FS_BOXW, pl->box_h))) return rc;
        if (out2) {
            if ((rc = make_tmap(ctx, &tm[2], ft.luma, H, W, Wp, FS_BOXW, bs))) return rc;
            if ((rc = make_tmap(ctx, &tm[3], fs.luma, H, W, Wp, FS_BOXW, pl->box_h))) return rc;
        } else { tm[2] = tm[0]; tm[3] = tm[1]; }
        FsParams p;
        p.H = H; p.W = W; p.bs = bs; p.R = R; p.Rp = (R + 3) & ~3; p.m = bs / 8; p.nbr = nbr; p.nbc = nbc;
        p.one = 1u;
        // warp schedulers favour the highest warp id, so the last warp runs ahead of the others: as the lead it writes results and
        // issues the next loads early instead of late (1.8 % on the 1080p kernel; start-up staggering of the warps and per-item
        // result slots reduced by the lead warp were measured too: no gain)
        p.lead_warp = FS_NT / 32 - 1;
        p.trace = nullptr;
        const char *trace_path = getenv("MEMCI_FS_TRACE");
        const size_t trace_n = (size_t)64 * (FS_NT / 32) * 3;
        if (trace_path) { CK(ctx, cudaMalloc(&p.trace, trace_n * 8)); CK(ctx, cudaMemsetAsync(p.trace, 0, trace_n * 8, ctx->stream)); }
        p.n_tiles_dir = pl->row_tile_start[br1] - pl->row_tile_start[br0]; p.n_tiles = ndir * p.n_tiles_dir; p.box_h = pl->box_h;
        { int ks = 1; while ((1 << ks) <= 2 * R * R + 1) ++ks; p.ks = ks; }
        p.n_stages = fs_smem_bytes(bs, pl->box_h, 3, pl->G) <= 200 * 1024 ? 3 : fs_smem_bytes(bs, pl->box_h, 2, pl->G) <= 220 * 1024 ? 2 : 1;
        if (const char *e = getenv("MEMCI_FS_STAGES")) { int v = atoi(e); if (v >= 1 && v < p.n_stages) p.n_stages = v; }   // tuning override
        p.tiles = pl->d_tiles + pl->row_tile_start[br0];
        p.out_vec[0] = out->vec; p.out_sad[0] = out->sad;
        p.out_vec[1] = out2 ? out2->vec : out->vec; p.out_sad[1] = out2 ? out2->sad : out->sad;
        p.redo_list = ctx->d_redo_list; p.counters = redo_count; p.redo_cap = ctx->redo_cap;
        p.gate = use8 ? ctx->d_counters + 8 : nullptr;
        const int smem = fs_smem_bytes(bs, pl->box_h, p.n_stages, pl->G);
        if (!use8) {
            if (ctx->timing) CK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
            prof = memci_prof_begin(ctx, MEMCI_PROF_FS, (double)ctx->fs_last_ops);
        } else {
            prof = memci_prof_begin(ctx, MEMCI_PROF_REFINE, 0.0);      // the gated exact kernel (normally a no-op) and the redo pass
        }
        switch (pl->G) {
        case 5: rc = fs_launch_tiled<5>(ctx, tm, p, smem); break;
        case 8: rc = fs_launch_tiled<8>(ctx, tm, p, smem); break;
        case 13: rc = fs_launch_tiled<13>(ctx, tm, p, smem); break;
        case 17: rc = fs_launch_tiled<17>(ctx, tm, p, smem); break;
        case 33: rc = fs_launch_tiled<33>(ctx, tm, p, smem); break;
        default: rc = fs_launch_tiled<11>(ctx, tm, p, smem); break;
        }
        if (!use8) { memci_prof_end(ctx, prof); prof = memci_prof_begin(ctx, MEMCI_PROF_REFINE, 0.0); }   // the redo pass counts as refinement either way
        if (p.trace) {                                   // debug dump: one line per (tile, warp)
            unsigned long long *h = (unsigned long long *)malloc(trace_n * 8);
            cudaStreamSynchronize(ctx->stream);
            cudaMemcpy(h, p.trace, trace_n * 8, cudaMemcpyDeviceToHost);
            if (FILE *f = fopen(trace_path, "w")) {
                for (size_t i = 0; i < trace_n / 3; ++i) fprintf(f, "%zu %zu %llu %llu %llu\n", i / (FS_NT / 32), i % (FS_NT / 32), h[3 * i], h[3 * i + 1], h[3 * i + 2]);
                fclose(f);
            }
            free(h); cudaFree(p.trace);
        }
        if (rc) return rc;
        if (ctx->timing) { CK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream)); ctx->fs_timed = 1; }
        // exact pass over the (normally few) blocks on the redo list
        const int redo_smem = fs_redo_smem_bytes(bs, R);
        if (redo_smem <= 160 * 1024) {
            CK(ctx, cudaFuncSetAttribute(fs_redo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, redo_smem));
            fs_redo_kernel<<<ctx->num_sms * 4, FS_REDO_NT, redo_smem, ctx->stream>>>(fs.luma, ft.luma, fs.rgb, ft.rgb, H, W, Wp, bs, R, nbc,
                                                                              ctx->d_redo_list, redo_count, out->vec, out->sad,
                                                                              p.out_vec[1], p.out_sad[1]);
        } else {
            fs_generic_kernel<<<ctx->num_sms, 256, 0, ctx->stream>>>(fs.luma, ft.luma, fs.rgb, ft.rgb, H, W, Wp, bs, R, 0, 0, nbc,
                                                                     ctx->d_redo_list, redo_count, out->vec, out->sad,
                                                                     p.out_vec[1], p.out_sad[1]);
        }
        if (prof >= 0) memci_prof_end(ctx, prof);
        CKL(ctx);
    } else {
        if (ctx->timing) CK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
        fs_generic_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(fs.luma, ft.luma, fs.rgb, ft.rgb, H, W, Wp, bs, R, br0 * nbc, br1 * nbc, nbc,
                                                                     nullptr, nullptr, out->vec, out->sad, nullptr, nullptr);
        CKL(ctx);
        if (out2) {
            fs_generic_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ft.luma, fs.luma, ft.rgb, fs.rgb, H, W, Wp, bs, R, br0 * nbc, br1 * nbc, nbc,
                                                                         nullptr, nullptr, out2->vec, out2->sad, nullptr, nullptr);
            CKL(ctx);
        }
        if (ctx->timing) { CK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream)); ctx->fs_timed = 1; }
    }
    if (use8 && !ctx->fs8_feedback_pending) {
        if (!ctx->h_fs8_feedback) {
            CK(ctx, cudaHostAlloc((void **)&ctx->h_fs8_feedback, 4 * sizeof(int), cudaHostAllocDefault));
            CK(ctx, cudaEventCreateWithFlags(&ctx->ev_fs8_feedback, cudaEventDisableTiming));
        }
        CK(ctx, cudaMemcpyAsync(ctx->h_fs8_feedback, ctx->d_counters + 8, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaEventRecord(ctx->ev_fs8_feedback, ctx->stream));
        ctx->fs8_feedback_pending = 1;
        ctx->fs8_feedback_cand = ctx->fs_last_cand;
    }
    out->valid = 1;
    if (out2) out2->valid = 1;
    return MEMCI_OK;
}
