// Full search, two-phase exact form: an 8-bit prefilter that prunes, then exact evaluation of what is left.
//
// The exact cost (fs.cu) needs 18-bit luma*1000 values: one VABSDIFF per pixel pair, 18.4 T pixel-ops/s on a B200.
// With luma rounded to 8 bits, L8 = round(L1000 / 1000), VABSDIFF4.U8.ACC handles four pixel pairs per instruction
// (66 T pixel-ops/s measured, profiles/microbench.py mode 15).  The approximate cost A(c) = sum |L8_a - L8_b| brackets
// the exact one: each rounded luma is within 0.5 of the real one, so every |a - b| term moves by at most 1 and
//     | S1000(c) / 1000 - A(c) | <= n,   n = bs * bs,
// hence the reference's SAD (floor(S1000/1000), or one less for a flagged candidate, fs.cu) lies in
// [A(c) - n - 1, A(c) + n].  Let A_min = min_c A(c), attained at c*.  SAD(c*) <= A_min + n, so a candidate can only win
// (or tie and then win on distance / scan order) if A(c) - n - 1 <= A_min + n:
//     survivors = { c : A(c) <= A_min + 2n + 1 }.
// The reference's winner is always among them; every survivor is then evaluated exactly (same arithmetic, same
// total order as fs.cu) and the minimum key taken.  The result is bit-identical; only the amount of exact work
// depends on the content (textured frames: ~0.5 % of the candidates survive; flat frames: all of them -- those blocks go
// to the block-level redo pass, and past a limit the exact kernel searches the pair, gated on a device-side flag).
//
//   fs8_prefilter_kernel  persistent CTAs over tiles of adjacent blocks of one block row; thread = (block, dx) candidate
//                         column; survivors appended to per-CTA segments of the list
//   fs8_refine_kernel     8 lanes per survivor: exact luma-SAD (+ float64 replay of flagged sums), 64-bit atomicMin
//   fs8_output_kernel     thread per block: key -> (dy, dx, sad); decides the pair-level fallback; re-arms the buffers
// Blocks that cannot be pruned (flat content, segment overflow, flagged sums of 16 x 16 blocks) go to fs_redo_kernel (fs.cu).
#include <stdlib.h>
#include <type_traits>

#include "memci_internal.cuh"

#define F8_G 11
#define F8_NBMAX 32
#define F8_MAXSEG 1024          // CTAs of the prefilter grid = segments of the survivor list

struct F8Params {
    int H, W, bs, m, R, Rp, nbr, nbc;
    int nb_tile, tiles_per_row, n_tiles_dir, br0, br1, ndir;
    int rpt;                     // block rows per tile (2 when the taller window fits: staging and barriers per block halve)
    int pitch8;                  // bytes per row of the padded L8 planes
    int win_rows;                // staged window rows incl. slack for the last group
    int rs;                      // words per window row (per pre-shifted plane)
    int plane_words;             // words per plane, = 8 (mod 32): the four planes of a row land in different banks
    int ref_rs;                  // words per row of the staged source blocks
    int a_stride;                // uint16 per candidate column in the A array (multiple of 4)
    int T;                       // survivor threshold 2 n + 1
    const uint8_t *l8_src[2], *l8_tgt[2];   // per direction: pixel (0, 0) inside the zero-padded plane
    unsigned *list; int seg_cap;   // survivor list: one segment of seg_cap entries per CTA
    unsigned *seg_counts;         // [grid] entries written per segment
    unsigned *redo_mark;          // [2][n_blk] 0 = block already on the redo list
    int *redo_list; int redo_cap; // blocks for fs_redo_kernel: flat / overflowing blocks here, flagged sums in the refine pass
    int redo_limit;               // more redo blocks than this: the pair falls back to the exact kernel (fs8_output_kernel decides)
    int *counters;               // [8] survivor list overflowed, [9] survivors
};

struct FsBounds8 { int r0, r1, c0, c1; };
// candidate window of one block: full_search.py:33-43 incl. the H-for-W comparison at :37 (same as fs.cu fs_bounds)
__device__ __forceinline__ FsBounds8 f8_bounds(int H, int W, int bs, int R, int s_row, int s_col)
{
    FsBounds8 b;
    const int tmr = (s_row + bs >= H) ? s_row + 1 : H - bs;
    const int tmc = (s_col + bs >= H) ? s_col + 1 : W - bs;
    b.r0 = max(s_row - R, 0); b.r1 = min(tmr, s_row + R + 1);
    b.c0 = max(s_col - R, 0); b.c1 = min(tmc, s_col + R + 1);
    return b;
}

// acc + sum of the four byte-wise |a - b| (VABSDIFF4.U8.ACC with the accumulator as third operand; written as PTX
// because nvcc otherwise emits the instruction with a zero accumulator plus a separate IADD3)
__device__ __forceinline__ unsigned f8_sad4(unsigned a, unsigned b, unsigned acc)
{
    unsigned r;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
    return r;
}

__device__ __forceinline__ void f8_cp_async4(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void f8_cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
}

// One group of F8_G consecutive row offsets of one candidate column: 8 x 8 source sub-block in rl / rh, window rows
// streamed from the lane's pre-shifted plane.  MASK: the group sticks out of the valid row range (last group only).
template <bool MASK>
__device__ __forceinline__ void f8_finish_group(unsigned (&acc)[F8_G], int n_valid, unsigned short *cg, unsigned &cmin)
{
    if (MASK) {
#pragma unroll
        for (int gi = 0; gi < F8_G; ++gi) acc[gi] = gi < n_valid ? acc[gi] : 0xffffu;
    }
#pragma unroll
    for (int gi = 0; gi < F8_G; ++gi) cg[gi] = (unsigned short)acc[gi];
#pragma unroll
    for (int gi = 0; gi + 1 < F8_G; gi += 2) cmin = min(cmin, min(acc[gi], acc[gi + 1]));     // VIMNMX3
    if (F8_G & 1) cmin = min(cmin, acc[F8_G - 1]);
}

__device__ __forceinline__ void f8_accumulate(unsigned (&acc)[F8_G], const unsigned (&rl)[8], const unsigned (&rh)[8],
                                              const unsigned *wp, int rs)
{
#pragma unroll
    for (int w = 0; w < F8_G + 7; ++w) {
        const unsigned lo = wp[w * rs], hi = wp[w * rs + 1];
#pragma unroll
        for (int gi = 0; gi < F8_G; ++gi) {
            const int kk = w - gi;
            if (kk >= 0 && kk < 8) acc[gi] = f8_sad4(rh[kk], hi, f8_sad4(rl[kk], lo, acc[gi]));
        }
    }
}

// Persistent: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  While a tile is searched, the raw window
// and source rows of the CTA's next tile stream into shared memory with cp.async; after the search they are expanded
// into the four byte-shifted planes (shared -> shared), so no global-memory latency sits between two tiles.
template <bool M1, int F8_NT>
__global__ void __launch_bounds__(F8_NT, 1024 / F8_NT)
fs8_prefilter_kernel(const F8Params p)
{
    extern __shared__ __align__(16) unsigned f8_smem[];
    __shared__ unsigned bmin_s[2][2][F8_NBMAX];  // [tile parity][block row of the tile]; per tile parity: the next tile's geometry is written while slow warps still emit
    __shared__ int pre_s[2][F8_NBMAX + 1], dx_lo_ss[2][F8_NBMAX];
    __shared__ unsigned seg_count;
    __shared__ int uniform_s[2];                                   // > 0: every block of the tile has that many column offsets
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bs = p.bs, R = p.R, raw_rs = p.rs + 1;
    unsigned *planes = f8_smem;                                    // [4][plane_words]
    unsigned *raw = planes + 4 * p.plane_words;                    // [win_rows][rs + 1]   next tile's window, unshifted
    const int ref_rows = p.rpt * bs;
    unsigned *ref0 = raw + p.win_rows * raw_rs;                    // [2][rpt * bs][ref_rs]  source blocks, current / next tile
    unsigned short *A = reinterpret_cast<unsigned short *>(ref0 + 2 * ref_rows * p.ref_rs);   // [threads][a_stride]
    const int n_tiles = p.n_tiles_dir * p.ndir;

    auto tile_of = [&](int t, int &dir, int &by, int &bx0, int &nb) {
        dir = t >= p.n_tiles_dir ? 1 : 0;
        const int tt = t - dir * p.n_tiles_dir, trow = tt / p.tiles_per_row;
        by = p.br0 + trow * p.rpt; bx0 = (tt - trow * p.tiles_per_row) * p.nb_tile;      // first block row of the tile
        nb = min(p.nb_tile, p.nbc - bx0);
    };
    auto issue_loads = [&](int dir, int by, int bx0, int nb, int buf) {
        const int s_row = by * bs;
        const uint8_t *g = (dir ? p.l8_tgt[1] : p.l8_tgt[0]) + (ptrdiff_t)(s_row - R) * p.pitch8 + (bx0 * bs - p.Rp);
        const int n_w = (nb * bs + 2 * p.Rp) / 4 + 2;              // words per row incl. the one the last funnel shift reads
        for (int r = warp; r < p.win_rows; r += F8_NT / 32) {
            const unsigned *row = reinterpret_cast<const unsigned *>(g + (ptrdiff_t)r * p.pitch8);
            for (int j = lane; j < n_w; j += 32) f8_cp_async4(raw + r * raw_rs + j, row + j);
        }
        const uint8_t *gs = (dir ? p.l8_src[1] : p.l8_src[0]) + (ptrdiff_t)s_row * p.pitch8 + bx0 * bs;
        unsigned *ref = ref0 + buf * ref_rows * p.ref_rs;
        const int n_r = nb * bs / 4;
        for (int r = warp; r < ref_rows; r += F8_NT / 32)                  // rows past the frame are zero in the padded plane
            for (int j = lane; j < n_r; j += 32)
                f8_cp_async4(ref + r * p.ref_rs + j, reinterpret_cast<const unsigned *>(gs + (ptrdiff_t)r * p.pitch8) + j);
    };
    // raw -> four byte-shifted planes; tile geometry (warp 0)
    auto build = [&](int by, int bx0, int nb, int par) {
        int *pre = pre_s[par], *dx_lo_s = dx_lo_ss[par];
        const int n_w = (nb * bs + 2 * p.Rp) / 4 + 1;
        for (int r = warp; r < p.win_rows; r += F8_NT / 32) {
            const unsigned *src = raw + r * raw_rs;
            unsigned *dst = planes + r * p.rs;
            for (int j = lane; j < n_w; j += 32) {
                const unsigned w0 = src[j], w1 = src[j + 1];
                dst[j] = w0;
                dst[p.plane_words + j] = __funnelshift_r(w0, w1, 8);
                dst[2 * p.plane_words + j] = __funnelshift_r(w0, w1, 16);
                dst[3 * p.plane_words + j] = __funnelshift_r(w0, w1, 24);
            }
        }
        if (warp == 0) {
            int n_dx = 0, dxl = 0;
            if (lane < nb) {
                const FsBounds8 b = f8_bounds(p.H, p.W, bs, R, by * bs, (bx0 + lane) * bs);
                n_dx = max(b.c1 - b.c0, 0);
                dxl = b.c0 - (bx0 + lane) * bs;
            }
            int incl = n_dx;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            pre[lane + 1] = incl; dx_lo_s[lane] = dxl; bmin_s[par][0][lane] = 0xffffffffu; bmin_s[par][1][lane] = 0xffffffffu;
            const int n0 = __shfl_sync(0xffffffffu, n_dx, 0);
            const bool uni = __all_sync(0xffffffffu, lane >= nb || n_dx == n0);
            if (lane == 0) { pre[0] = 0; uniform_s[par] = (uni && n0 > 0) ? n0 : 0; }
        }
    };

    unsigned *seg = p.list + (size_t)blockIdx.x * p.seg_cap;
    if (tid == 0) seg_count = 0u;
    int t = blockIdx.x;
    if (t >= n_tiles) return;
    int cur = 0;
    int dir, by, bx0, nb;                                          // geometry of the tile being searched
    tile_of(t, dir, by, bx0, nb);
    issue_loads(dir, by, bx0, nb, 0);
    f8_cp_async_wait_all();
    __syncwarp();
    build(by, bx0, nb, 0);                                        // every warp expands the rows it copied itself
    __syncthreads();

    for (;;) {
        const int tn = t + gridDim.x;
        int dir_n = 0, by_n = 0, bx0_n = 0, nb_n = 0;              // ... and of the CTA's next tile
        if (tn < n_tiles) {
            tile_of(tn, dir_n, by_n, bx0_n, nb_n);
            issue_loads(dir_n, by_n, bx0_n, nb_n, cur ^ 1);        // streams in while this tile is searched
        }

        const int *pre = pre_s[cur], *dx_lo_s = dx_lo_ss[cur];
        const int n_items = pre[nb];
        int b = 0, dxi = 0, dx = 0;                                // the thread's candidate column: the same for every block row of the tile
        if (tid < n_items) {
            if (uniform_s[cur] > 0) b = tid / uniform_s[cur];
            else while (b + 1 < nb && pre[b + 1] <= tid) ++b;
            dxi = tid - pre[b];
            dx = dx_lo_s[b] + dxi;
        }
        unsigned short *col = A + (size_t)tid * p.a_stride;
        const int nv = min(p.rpt, p.br1 - by);
#pragma unroll 1
        for (int v = 0; v < nv; ++v) {
            unsigned *bmin = bmin_s[cur][v];
            const int s_row = (by + v) * bs;
            const unsigned *ref = ref0 + (cur * ref_rows + v * bs) * p.ref_rs;
            const FsBounds8 b0 = f8_bounds(p.H, p.W, bs, R, s_row, bx0 * bs);
            const int n_dy = max(b0.r1 - b0.r0, 0);
            const int wrow_base = b0.r0 - (by * bs - R);               // window rows count from the tile's first block row
            const bool active = tid < n_items && n_dy > 0;
            unsigned cmin = 0xffffffffu;
            if (active) {
                unsigned rl[8], rh[8];
                if (M1) {
                    const unsigned *rp = ref + ((b * 8) >> 2);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint2 rv = *reinterpret_cast<const uint2 *>(rp + kk * p.ref_rs);
                        rl[kk] = rv.x; rh[kk] = rv.y;
                    }
                }
                const int wc0 = b * bs + dx + p.Rp;                    // byte offset inside a window row
                // full groups first, then (clipped windows only) the one group that sticks out of the valid row range: two
                // separate copies of the loop body, so that the common one carries no masking selects
                auto group = [&](int g, auto masked) {
                    unsigned acc[F8_G];
#pragma unroll
                    for (int i = 0; i < F8_G; ++i) acc[i] = 0u;
                    if (M1) {
                        f8_accumulate(acc, rl, rh, planes + (wc0 & 3) * p.plane_words + (wrow_base + g * F8_G) * p.rs + (wc0 >> 2), p.rs);
                    } else {
                        const int m = p.m;
#pragma unroll 1
                        for (int sub = 0; sub < m * m; ++sub) {
                            const int si = sub / m, sj = sub - si * m;
                            const unsigned *rp = ref + (si * 8) * p.ref_rs + ((b * bs + sj * 8) >> 2);
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk) {
                                const uint2 rv = *reinterpret_cast<const uint2 *>(rp + kk * p.ref_rs);
                                rl[kk] = rv.x; rh[kk] = rv.y;
                            }
                            const int wc = wc0 + sj * 8;
                            f8_accumulate(acc, rl, rh, planes + (wc & 3) * p.plane_words + (wrow_base + g * F8_G + si * 8) * p.rs + (wc >> 2), p.rs);
                        }
                    }
                    f8_finish_group<decltype(masked)::value>(acc, n_dy - g * F8_G, col + g * F8_G, cmin);
                };
                const int n_full = n_dy / F8_G;
#pragma unroll 1
                for (int g = 0; g < n_full; ++g) group(g, std::false_type{});
                if (n_full * F8_G < n_dy) group(n_full, std::true_type{});
            }
            // ---- block minima: segmented warp reduction, one shared-memory atomic per (warp, block) ----
            {
                unsigned todo = __ballot_sync(0xffffffffu, active);
                while (todo) {
                    const int leader = __ffs(todo) - 1;
                    const int bb = __shfl_sync(0xffffffffu, b, leader);
                    const bool mine = active && b == bb;
                    const unsigned mn = __reduce_min_sync(0xffffffffu, mine ? cmin : 0xffffffffu);
                    if (lane == leader) atomicMin(&bmin[bb], mn);
                    todo &= ~__ballot_sync(0xffffffffu, mine);
                }
            }
            __syncthreads();
            // ---- survivors: the warp visits its flagged columns one at a time, lanes = row offsets.  Each CTA appends to
            //      its own segment of the list (shared-memory cursor, no global atomics); fs8_refine_kernel gets the
            //      segment lengths at the end ----
            {
                const unsigned thr = active ? bmin[b] + (unsigned)p.T : 0u;
                const unsigned flagged0 = __ballot_sync(0xffffffffu, active && cmin <= thr);
                const unsigned head_own = (dir ? 0x80000000u : 0u) | ((unsigned)((by + v) * p.nbc + bx0 + b) << 14) | (unsigned)dxi;
                const unsigned short *cw = A + (size_t)(tid & ~31) * p.a_stride;
                for (unsigned f = flagged0; f; f &= f - 1) {
                    const int src = __ffs(f) - 1;
                    const unsigned thr_c = __shfl_sync(0xffffffffu, thr, src);
                    const unsigned head = __shfl_sync(0xffffffffu, head_own, src);
                    // lane l looks at row offsets 2l and 2l + 1 (one 32-bit load): 64 offsets per pass
                    const unsigned *cs = reinterpret_cast<const unsigned *>(cw + src * p.a_stride);
                    int hits = 0;
                    for (int i0 = 0; i0 < n_dy; i0 += 64) {
                        const int i = i0 + 2 * lane;
                        const unsigned pv = i < n_dy ? cs[i >> 1] : 0xffffffffu;
                        hits += __popc(__ballot_sync(0xffffffffu, i < n_dy && (pv & 0xffffu) <= thr_c)) +
                                __popc(__ballot_sync(0xffffffffu, i + 1 < n_dy && (pv >> 16) <= thr_c));
                    }
                    int pos = 0;
                    if (lane == 0) {
                        // A column that survives whole is flat content: nothing to prune in this block.  It, and any block
                        // whose survivors no longer fit the segment, goes to the exact block-level redo pass instead (once).
                        pos = hits == n_dy ? -1 : (int)atomicAdd(&seg_count, (unsigned)hits);
                        if (pos >= 0 && pos + hits > p.seg_cap) { atomicSub(&seg_count, (unsigned)hits); pos = -1; }
                        if (pos < 0) {
                            const unsigned slot = (head >> 31) * (unsigned)(p.nbr * p.nbc) + ((head >> 14) & 0x1ffffu);
                            if (atomicExch(&p.redo_mark[slot], 0u) != 0u) {
                                const int rp = atomicAdd(&p.counters[11], 1);
                                if (rp < p.redo_cap) p.redo_list[rp] = (int)((head >> 14) & 0x1ffffu) | ((head >> 31) ? 0x40000000 : 0);   // FS_DIR_BIT
                            }
                        }
                    }
                    pos = __shfl_sync(0xffffffffu, pos, 0);
                    if (pos < 0) continue;
                    for (int i0 = 0; i0 < n_dy; i0 += 64) {
                        const int i = i0 + 2 * lane;
                        const unsigned pv = i < n_dy ? cs[i >> 1] : 0xffffffffu;
                        const bool h0 = i < n_dy && (pv & 0xffffu) <= thr_c, h1 = i + 1 < n_dy && (pv >> 16) <= thr_c;
                        const unsigned b0m = __ballot_sync(0xffffffffu, h0), b1m = __ballot_sync(0xffffffffu, h1);
                        const unsigned below = (1u << lane) - 1u;
                        const int off = __popc(b0m & below) + __popc(b1m & below);      // entries of lower lanes come first (order is irrelevant to the result)
                        if (h0) seg[pos + off] = head | ((unsigned)i << 7);
                        if (h1) seg[pos + off + (h0 ? 1 : 0)] = head | ((unsigned)(i + 1) << 7);
                        pos += __popc(b0m) + __popc(b1m);
                    }
                }
            }
        }                                                          // block rows of the tile
        if (tn >= n_tiles) break;
        f8_cp_async_wait_all();
        __syncwarp();                                              // the warp's own rows of the next tile landed; planes are free since the barrier above
        build(by_n, bx0_n, nb_n, cur ^ 1);
        __syncthreads();
        t = tn; cur ^= 1;
        dir = dir_n; by = by_n; bx0 = bx0_n; nb = nb_n;
    }
    __syncthreads();
    if (tid == 0) { p.seg_counts[blockIdx.x] = min(seg_count, (unsigned)p.seg_cap); atomicAdd(&p.counters[9], (int)seg_count); }
}

__device__ __forceinline__ unsigned long long f8_key(unsigned sad, unsigned d2, unsigned ri, unsigned ci)
{
    return ((unsigned long long)sad << 32) | ((unsigned long long)d2 << 16) | (ri << 8) | ci;    // = fs.cu fs_key
}

// Eight lanes per survivor: lane l covers row l of each 8 x 8 sub-block -- the source row as two aligned 16-byte
// vectors, the target row as one 32-byte run -- so a survivor costs a few dozen sectors instead of 128 scattered words.
// Blocks already on the redo list (dense blocks, earlier flagged sums) are skipped.
__global__ void __launch_bounds__(256)
fs8_refine_kernel(const int32_t *__restrict__ L0, const int32_t *__restrict__ L1,
                  const uint8_t *__restrict__ rgb0, const uint8_t *__restrict__ rgb1,
                  int H, int W, int Wp, int bs, int R, int nbc, int n_blk,
                  const unsigned *__restrict__ list, int seg_cap, const unsigned *__restrict__ seg_counts, int n_seg,
                  int *__restrict__ counters,
                  unsigned long long *__restrict__ best, unsigned *__restrict__ redo_mark,
                  int *__restrict__ redo_list, int redo_cap)
{
    if (counters[8]) return;                                       // overflow: the exact kernel takes the pair
    // exclusive prefix of the segment lengths (n_seg <= F8_MAXSEG, every CTA computes it for itself)
    __shared__ int seg_pre[F8_MAXSEG + 1];
    {
        __shared__ int wsum[8];
        const int per = (n_seg + 255) / 256;
        int loc = 0;
        for (int k = 0; k < per; ++k) { const int si = threadIdx.x * per + k; if (si < n_seg) loc += (int)seg_counts[si]; }
        int incl = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += v; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += wsum[w];
        int run = wbase + incl - loc;
        for (int k = 0; k < per; ++k) { const int si = threadIdx.x * per + k; if (si < n_seg) { seg_pre[si] = run; run += (int)seg_counts[si]; } }
        if (threadIdx.x == 255) seg_pre[n_seg] = run;
        __syncthreads();
    }
    const int n = seg_pre[n_seg];
    const int lane = threadIdx.x & 31, pr = lane & 7;
    const int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, n_grp = (gridDim.x * blockDim.x) >> 3;
    const int m = bs >> 3;
    const float inv_nbc = 1.0f / (float)nbc;
    // neighbouring 8-lane groups take neighbouring survivors (same block, same column: shared source rows and target
    // rows in L1 at the same time)
    for (int i0 = 0; i0 < n; i0 += n_grp) {                        // warp-uniform trip count (shuffles inside)
        const int i = i0 + grp;
        const bool on = i < n;
        unsigned e = 0u;
        if (on) {                                                  // global index -> (segment, offset)
            int lo = 0, hi = n_seg;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seg_pre[mid] <= i) lo = mid; else hi = mid; }
            e = list[(size_t)lo * seg_cap + (i - seg_pre[lo])];
        }
        const int dir = e >> 31, blk = (int)((e >> 14) & 0x1ffffu), dyi = (int)((e >> 7) & 127u), dxi = (int)(e & 127u);
        const int32_t *Ls = dir ? L1 : L0, *Lt = dir ? L0 : L1;
        int br = (int)(((float)blk + 0.5f) * inv_nbc);             // blk / nbc (blk < 2^17: the float quotient is off by at most one)
        br -= (br * nbc > blk); br += ((br + 1) * nbc <= blk);
        const int bc = blk - br * nbc;
        const int s_row = br * bs, s_col = bc * bs;
        const int t_row = max(s_row - R, 0) + dyi, t_col = max(s_col - R, 0) + dxi;             // window origin (r0, c0) of f8_bounds
        unsigned S = 0;
        if (on) {
            if (s_row + bs <= H && s_col + bs <= W && t_row + bs <= H && t_col + bs <= W) {      // no padding involved
                for (int sub = 0; sub < m * m; ++sub) {
                    const int si = sub / m, sj = sub - si * m;
                    const int r = si * 8 + pr, c = sj * 8;
                    const int32_t *ps = Ls + (size_t)(s_row + r) * Wp + s_col + c;
                    const int4 a0 = *reinterpret_cast<const int4 *>(ps), a1 = *reinterpret_cast<const int4 *>(ps + 4);
                    const int32_t *q = Lt + (size_t)(t_row + r) * Wp + t_col + c;
                    S = __sad(a0.x, q[0], S); S = __sad(a0.y, q[1], S); S = __sad(a0.z, q[2], S); S = __sad(a0.w, q[3], S);
                    S = __sad(a1.x, q[4], S); S = __sad(a1.y, q[5], S); S = __sad(a1.z, q[6], S); S = __sad(a1.w, q[7], S);
                }
            } else {                                                // zero padding (np.pad) on either side
                for (int sub = 0; sub < m * m; ++sub) {
                    const int si = sub / m, sj = sub - si * m;
                    const int r1 = s_row + si * 8 + pr, r2 = t_row + si * 8 + pr;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int c1 = s_col + sj * 8 + k, c2 = t_col + sj * 8 + k;
                        const int a = (r1 < H && c1 < W) ? Ls[(size_t)r1 * Wp + c1] : 0;
                        const int v = (r2 < H && c2 < W) ? Lt[(size_t)r2 * Wp + c2] : 0;
                        S = __sad(a, v, S);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
        // Flagged sum (S a positive multiple of 1000): the reference's float64 accumulation may land just below the
        // integer and truncate to S / 1000 - 1 (fs.cu).  8 x 8 blocks are replayed here in the reference's operation
        // order -- each lane forms the |dL| terms of its row, then the 64 terms are added row by row, column by column --
        // larger blocks go to the block-level redo pass.
        const bool flagged = on && S != 0u && (S / 1000u) * 1000u == S;
        unsigned sad = S / 1000u;
        if (bs == 8 && __any_sync(0xffffffffu, flagged)) {
            double tr[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) tr[c] = 0.0;
            if (flagged) {
                const uint8_t *rgb_s = dir ? rgb1 : rgb0, *rgb_t = dir ? rgb0 : rgb1;
                const int r1 = s_row + pr, r2 = t_row + pr;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int c1 = s_col + c, c2 = t_col + c;
                    double l1 = 0.0, l2 = 0.0;
                    if (r1 < H && c1 < W) { const uint8_t *q = rgb_s + ((size_t)r1 * W + c1) * 3; l1 = luma_f64(q[0], q[1], q[2]); }
                    if (r2 < H && c2 < W) { const uint8_t *q = rgb_t + ((size_t)r2 * W + c2) * 3; l2 = luma_f64(q[0], q[1], q[2]); }
                    tr[c] = fabs(__dsub_rn(l1, l2));
                }
            }
            double acc = 0.0;
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int c = 0; c < 8; ++c) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, tr[c], i, 8));
            }
            if (flagged) { sad = (unsigned)acc; if (pr == 0) atomicAdd(&counters[10], 1); }
        }
        if (on && pr == 0) {
            if (flagged && bs != 8) {
                if (atomicExch(&redo_mark[(size_t)dir * n_blk + blk], 0u) != 0u) {
                    const int pos = atomicAdd(&counters[11], 1);
                    if (pos < redo_cap) redo_list[pos] = blk | (dir ? 0x40000000 : 0);           // FS_DIR_BIT
                }
            }
            const int dr = s_row - t_row, dc = s_col - t_col;
            atomicMin(&best[(size_t)dir * n_blk + blk], f8_key(sad, (unsigned)(dr * dr + dc * dc), (unsigned)dyi, (unsigned)dxi));
        }
    }
}

// Also re-arms the key / redo-mark buffers for the next search (no memset per call).
__global__ void fs8_output_kernel(unsigned long long *__restrict__ best, unsigned *__restrict__ redo_mark, int *__restrict__ counters, int redo_limit,
                                  int H, int W, int bs, int R, int nbc, int n_blk, int blk_begin, int blk_end, int ndir,
                                  float2 *__restrict__ out_vec0, float *__restrict__ out_sad0,
                                  float2 *__restrict__ out_vec1, float *__restrict__ out_sad1)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = blk_end - blk_begin;
    if (i >= per * ndir) return;
    // too many blocks for the block-level redo pass (mostly flat content): the exact kernel takes the whole pair
    const bool fallback = counters[8] != 0 || counters[11] > redo_limit;
    if (i == 0 && fallback) { counters[8] = 1; counters[11] = 0; }
    const int dir = i / per, blk = blk_begin + (i - dir * per);
    const size_t slot = (size_t)dir * n_blk + blk;
    const unsigned long long k = best[slot];
    best[slot] = 0xffffffffffffffffull;
    redo_mark[slot] = 0xffffffffu;
    if (fallback) return;                                          // the exact kernel writes the vectors
    const int br = blk / nbc, bc = blk - br * nbc, s_row = br * bs, s_col = bc * bs;
    float dy, dx, sad;
    if (k == 0xffffffffffffffffull) {                              // empty window: target_index stays (0,0), sad = inf
        dy = (float)(-s_row); dx = (float)(-s_col); sad = __int_as_float(0x7f800000);
    } else {
        const FsBounds8 b = f8_bounds(H, W, bs, R, s_row, s_col);
        dy = (float)(b.r0 + (int)((k >> 8) & 255) - s_row); dx = (float)(b.c0 + (int)(k & 255) - s_col);
        sad = (float)(unsigned)(k >> 32);
    }
    (dir ? out_vec1 : out_vec0)[blk] = make_float2(dy, dx);
    (dir ? out_sad1 : out_sad0)[blk] = sad;
}

// ---- rounded 8-bit luma plane with a zero border (np.pad of the reference and every out-of-frame read) ----
__global__ void luma8_kernel(const int32_t *__restrict__ L, int H, int W, int Wp, uint8_t *__restrict__ l8, int pitch8)
{
    const int quads = Wp >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * quads) return;
    const int y = idx / quads, x = (idx - y * quads) << 2;
    const int4 v = *reinterpret_cast<const int4 *>(L + (size_t)y * Wp + x);     // pad columns of L are 0
    const unsigned w = (unsigned)((v.x + 500) / 1000) | ((unsigned)((v.y + 500) / 1000) << 8) |
                       ((unsigned)((v.z + 500) / 1000) << 16) | ((unsigned)((v.w + 500) / 1000) << 24);
    *reinterpret_cast<unsigned *>(l8 + (size_t)y * pitch8 + x) = w;
}

int memci_ensure_luma8(memci_ctx *ctx, int slot)
{
    FrameSlot &f = ctx->frames[slot];
    int rc = memci_ensure_luma(ctx, slot);
    if (rc) return rc;
    if (f.luma8_ok) return MEMCI_OK;
    const int pitch8 = ctx->Wp + 2 * MEMCI_L8_PAD;
    if (!f.luma8) {
        const size_t bytes = (size_t)(ctx->H + 2 * MEMCI_L8_PAD) * pitch8 + 16;
        CK(ctx, cudaMalloc(&f.luma8, bytes));
        CK(ctx, cudaMemsetAsync(f.luma8, 0, bytes, ctx->stream));
    }
    uint8_t *org = f.luma8 + (size_t)MEMCI_L8_PAD * pitch8 + MEMCI_L8_PAD;
    const int n = ctx->H * (ctx->Wp >> 2);
    luma8_kernel<<<ceil_div_i(n, 256), 256, 0, ctx->stream>>>(f.luma, ctx->H, ctx->W, ctx->Wp, org, pitch8);
    CKL(ctx);
    f.luma8_ok = 1;
    return MEMCI_OK;
}

// Is the prefilter path applicable?  Block sizes 8k, list entry fields (17-bit block, 7-bit offsets), border of the
// padded plane, one candidate column per thread, shared memory.
static int f8_smem_bytes(int nt, int rpt, int bs, int R, int nb_tile, int *rs_out, int *plane_words_out, int *ref_rs_out, int *a_stride_out, int *win_rows_out)
{
    const int Rp = (R + 3) & ~3;
    const int n_dy = 2 * R + 1;
    const int n_groups = (n_dy + F8_G - 1) / F8_G;
    const int win_rows = rpt * bs + 2 * R + (n_groups * F8_G - n_dy) + 1;
    const int rs = (nb_tile * bs + 2 * Rp) / 4 + 1;
    int plane_words = win_rows * rs;
    plane_words += (8 - (plane_words & 31) + 32) & 31;
    const int ref_rs = nb_tile * bs / 4;
    const int a_stride = (n_groups * F8_G + 3) & ~3;
    *rs_out = rs; *plane_words_out = plane_words; *ref_rs_out = ref_rs; *a_stride_out = a_stride; *win_rows_out = win_rows;
    return (4 * plane_words + win_rows * (rs + 1) + 2 * rpt * bs * ref_rs) * 4 + nt * a_stride * 2;
}

// CTA shape: 512 threads (two CTAs per SM, wider tiles: 15 blocks of 33 column offsets fill 97 % of the lanes at +-16)
// when the shared memory of two such CTAs fits, else 256 threads; two block rows per tile when that still fits.
// MEMCI_F8_NT / MEMCI_F8_RPT override (tuning).
static void f8_shape(int bs, int R, int *nt_out, int *rpt_out)
{
    int a, b, c, d, e2;
    int nt = 256, rpt = 1;
    {
        int nb_tile = 512 / (2 * R + 1);
        if (nb_tile > F8_NBMAX) nb_tile = F8_NBMAX;
        if (nb_tile >= 1 && f8_smem_bytes(512, 1, bs, R, nb_tile, &a, &b, &c, &d, &e2) <= 110 * 1024) nt = 512;
    }
    if (const char *e = getenv("MEMCI_F8_NT")) { const int v = atoi(e); if (v == 256 || v == 512) nt = v; }
    {
        int nb_tile = nt / (2 * R + 1);
        if (nb_tile > F8_NBMAX) nb_tile = F8_NBMAX;
        if (nb_tile >= 1 && R + 2 * bs + 16 <= MEMCI_L8_PAD && f8_smem_bytes(nt, 2, bs, R, nb_tile, &a, &b, &c, &d, &e2) <= 110 * 1024) rpt = 2;
    }
    if (const char *e = getenv("MEMCI_F8_RPT")) { const int v = atoi(e); if (v == 1) rpt = 1; }
    *nt_out = nt; *rpt_out = rpt;
}

int memci_fs8_eligible(memci_ctx *ctx, int bs, int R)
{
    if ((bs != 8 && bs != 16) || R < 1 || R > 63 || R + bs + 16 > MEMCI_L8_PAD) return 0;      // A fits uint16 up to 16 x 16 blocks
    const long long n_blk = (long long)ceil_div_i(ctx->H, bs) * ceil_div_i(ctx->W, bs);
    if (n_blk >= (1 << 17)) return 0;
    int nt, rpt;
    f8_shape(bs, R, &nt, &rpt);
    int nb_tile = nt / (2 * R + 1);
    if (nb_tile < 1) return 0;
    int a, b, c, d, e;
    if (nb_tile > F8_NBMAX) nb_tile = F8_NBMAX;
    return f8_smem_bytes(nt, rpt, bs, R, nb_tile, &a, &b, &c, &d, &e) <= 110 * 1024;
}

// Enqueue prefilter + refine + output for block rows [br0, br1).  counters[8] is left non-zero when the survivor
// list overflowed; the caller then runs the exact kernel gated on that flag.
int memci_fs8_run(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int R, MVField *out, int br0, int br1, MVField *out2)
{
    const int H = ctx->H, W = ctx->W;
    int rc;
    if ((rc = memci_ensure_luma8(ctx, slot_src))) return rc;
    if ((rc = memci_ensure_luma8(ctx, slot_tgt))) return rc;
    const int nbr = ceil_div_i(H, bs), nbc = ceil_div_i(W, bs), n_blk = nbr * nbc, ndir = out2 ? 2 : 1;
    F8Params p;
    memset(&p, 0, sizeof(p));
    p.H = H; p.W = W; p.bs = bs; p.m = bs / 8; p.R = R; p.Rp = (R + 3) & ~3; p.nbr = nbr; p.nbc = nbc;
    int nt, rpt;
    f8_shape(bs, R, &nt, &rpt);
    p.rpt = rpt;
    p.nb_tile = nt / (2 * R + 1);
    if (p.nb_tile > F8_NBMAX) p.nb_tile = F8_NBMAX;
    if (p.nb_tile > nbc) p.nb_tile = nbc;
    p.tiles_per_row = ceil_div_i(nbc, p.nb_tile);
    p.n_tiles_dir = p.tiles_per_row * ceil_div_i(br1 - br0, rpt); p.br0 = br0; p.br1 = br1; p.ndir = ndir;
    p.pitch8 = ctx->Wp + 2 * MEMCI_L8_PAD;
    const int smem = f8_smem_bytes(nt, rpt, bs, R, p.nb_tile, &p.rs, &p.plane_words, &p.ref_rs, &p.a_stride, &p.win_rows);
    p.T = 2 * bs * bs + 1;
    const FrameSlot &fs = ctx->frames[slot_src], &ft = ctx->frames[slot_tgt];
    const size_t org = (size_t)MEMCI_L8_PAD * p.pitch8 + MEMCI_L8_PAD;
    p.l8_src[0] = fs.luma8 + org; p.l8_tgt[0] = ft.luma8 + org;
    p.l8_src[1] = ft.luma8 + org; p.l8_tgt[1] = fs.luma8 + org;
    // survivor list: an eighth of the candidates, at most 16 M entries (64 MB).  Textured content needs about 0.5 %, but the
    // list is cut into one segment per prefilter CTA and flat regions load a few segments heavily; a segment that fills
    // up sends its blocks to the redo pass, so the size is a performance knob, not a correctness one.
    const size_t cand_max = (size_t)n_blk * (2 * R + 1) * (2 * R + 1) * 2;
    size_t cap = cand_max / 8 < (1u << 20) ? (1u << 20) : cand_max / 8;
    if (cap > (16u << 20)) cap = 16u << 20;
    size_t have = ctx->fs8_list_cap;
    if ((rc = memci_reserve(ctx, (void **)&ctx->d_fs8_list, &have, (cap + F8_MAXSEG) * 4))) return rc;
    ctx->fs8_list_cap = have;
    size_t have_b = ctx->fs8_best_cap;
    if ((rc = memci_reserve(ctx, (void **)&ctx->d_fs8_best, &have_b, (size_t)n_blk * 2 * 12))) return rc;       // keys + redo marks
    if (have_b != ctx->fs8_best_cap) CK(ctx, cudaMemsetAsync(ctx->d_fs8_best, 0xff, have_b, ctx->stream));    // new buffer: armed once, then by fs8_output_kernel
    ctx->fs8_best_cap = have_b;
    p.counters = ctx->d_counters;
    p.redo_mark = reinterpret_cast<unsigned *>(ctx->d_fs8_best + (size_t)n_blk * 2);
    p.redo_list = ctx->d_redo_list; p.redo_cap = ctx->redo_cap; p.redo_limit = n_blk * ndir / 16 + 64;
    CK(ctx, cudaMemsetAsync(ctx->d_counters + 8, 0, 4 * sizeof(int), ctx->stream));   // [8] overflow [9] survivors [10] float64 replays [11] redo blocks
    if (!ctx->fs8_attr_set) {                                      // per context: the attribute belongs to the device
        CK(ctx, cudaFuncSetAttribute(fs8_prefilter_kernel<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        CK(ctx, cudaFuncSetAttribute(fs8_prefilter_kernel<false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        CK(ctx, cudaFuncSetAttribute(fs8_prefilter_kernel<true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        CK(ctx, cudaFuncSetAttribute(fs8_prefilter_kernel<false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        ctx->fs8_attr_set = 1;
    }
    const int n_tiles = p.n_tiles_dir * ndir;
    int per_sm = 1024 / nt;
    if (const char *e = getenv("MEMCI_F8_CTAS")) per_sm = atoi(e);          // tuning: persistent CTAs per SM
    int grid = n_tiles < per_sm * ctx->num_sms ? n_tiles : per_sm * ctx->num_sms;
    if (grid > F8_MAXSEG) grid = F8_MAXSEG;
    p.seg_counts = ctx->d_fs8_list;                                // [F8_MAXSEG] segment lengths, then the segments
    p.list = ctx->d_fs8_list + F8_MAXSEG;
    p.seg_cap = (int)((have / 4 - F8_MAXSEG) / grid);
    if (const char *e = getenv("MEMCI_F8_SEG_CAP")) { const int v = atoi(e); if (v >= 0 && v < p.seg_cap) p.seg_cap = v; }   // tests: force the overflow fallback
    const int prof = memci_prof_begin(ctx, 0);
    if (nt == 512) {
        if (bs == 8) fs8_prefilter_kernel<true, 512><<<grid, 512, smem, ctx->stream>>>(p);
        else fs8_prefilter_kernel<false, 512><<<grid, 512, smem, ctx->stream>>>(p);
    } else {
        if (bs == 8) fs8_prefilter_kernel<true, 256><<<grid, 256, smem, ctx->stream>>>(p);
        else fs8_prefilter_kernel<false, 256><<<grid, 256, smem, ctx->stream>>>(p);
    }
    memci_prof_end(ctx, prof);
    CKL(ctx);
    fs8_refine_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(fs.luma, ft.luma, fs.rgb, ft.rgb, H, W, ctx->Wp, bs, R, nbc, n_blk,
                                                                 p.list, p.seg_cap, p.seg_counts, grid, ctx->d_counters, ctx->d_fs8_best,
                                                                 reinterpret_cast<unsigned *>(ctx->d_fs8_best + (size_t)n_blk * 2), ctx->d_redo_list, ctx->redo_cap);
    CKL(ctx);
    const int n_out = (br1 - br0) * nbc * ndir;
    fs8_output_kernel<<<ceil_div_i(n_out, 256), 256, 0, ctx->stream>>>(ctx->d_fs8_best, reinterpret_cast<unsigned *>(ctx->d_fs8_best + (size_t)n_blk * 2), ctx->d_counters, p.redo_limit,
                                                                      H, W, bs, R, nbc, n_blk, br0 * nbc, br1 * nbc, ndir,
                                                                      out->vec, out->sad, out2 ? out2->vec : out->vec, out2 ? out2->sad : out->sad);
    CKL(ctx);
    return MEMCI_OK;
}
