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
//   luma8_kernel          L8 as FOUR planes, plane q shifted left by q bytes (zero border): a thread then reads the 8 window
//                         bytes of any column offset as two aligned words of plane (offset & 3), and TMA -- whose inner
//                         coordinate must be 16-byte aligned (probed: any other x raises an illegal instruction) -- can
//                         stage all four shifts with ONE 3-D tile load
//   fs8_prefilter_kernel  persistent CTAs, dynamic job queue; job = adjacent blocks of one block row (of two, side by side,
//                         where the column range is clipped), staged by TMA (cp.async.bulk.tensor.3d + mbarrier, two
//                         stages); thread = (block, dx) candidate column; survivors appended to per-CTA segments of the list
//   fs8_refine_kernel     8 lanes per survivor: exact luma-SAD, 64-bit atomicMin; flagged sums go on a short list
//   fs8_replay_kernel     float64 replay of the flagged sums in the reference's operation order
//   fs8_output_kernel     thread per block: key -> (dy, dx, sad); decides the pair-level fallback; re-arms the buffers
// Blocks that cannot be pruned (flat content, segment overflow, flagged sums of 16 x 16 blocks) go to fs_redo_kernel (fs.cu).
#include <stdlib.h>
#include <type_traits>

#include "memci_internal.cuh"

#define F8_G 11
#define F8_NBMAX 32
#define F8_MAXSEG 1024          // CTAs of the prefilter grid = segments of the survivor list
#define F8_FLAG_CAP 16384        // flagged survivors (sum a multiple of 1000) replayed in float64 per search
#define F8_EMIT_WORDS 9          // votes (eight row offsets each) a warp keeps in registers while it emits survivors

struct F8Params {
    int H, W, bs, m, R, Rp, nbr, nbc;
    int n_tiles;
    const struct F8Tile *tiles;  // the jobs, in queue order (host table, cached per search geometry)
    int n_stage;                 // shared-memory stages of (window planes + source blocks): 2 = the next tile lands while this one is searched
    unsigned c_two, c_8k, c_magic; float c_fmagic;   // the constants of F8Conv, opaque to the compiler
    int ablate;                  // timing experiments only (MEMCI_F8_ABLATE=1): no survivor emission (nothing survives: wrong results)
    int flush_at;                // the hand-over of the previous job happens after this many + 1 row-offset groups of the current one
    int box_x0, box_y0;          // padded-plane coordinates of frame pixel (0, 0)
    int rs;                      // words per staged window row (= TMA box width / 4)
    int plane_words;             // words per staged plane, = 8 or 24 (mod 32): the four planes of a row land in different banks
    int ref_rs, ref_rows;        // words per row / rows of the staged source blocks
    int a_rows;                  // row offsets per A buffer (groups * F8_G)
    unsigned tx_bytes;           // bytes one tile's TMA loads deliver
    int T;                       // survivor threshold 2 n + 1
    unsigned *list; int seg_cap;   // survivor list: one segment of seg_cap entries per CTA
    unsigned *seg_counts;         // [grid] entries written per segment
    unsigned *redo_mark;          // [2][n_blk] 0 = block already on the redo list
    int *redo_list; int redo_cap; // blocks for fs_redo_kernel: flat / overflowing blocks here, flagged sums in the refine pass
    int redo_limit;               // more redo blocks than this: the pair falls back to the exact kernel (fs8_output_kernel decides)
    int *counters;               // [8] survivor list overflowed, [9] survivors, [11] redo blocks, [12] tile queue cursor
};

struct F8Tile { short dir, nv; short by, bx0; short nb, pad; };   // direction, block rows (1 or 2), first block row / column, blocks

struct F8Geo {                   // geometry of one job; written by warp 0 before it arms the job's loads
    int t;                       // index in the table, < 0: the queue is empty
    int dir, by, bx0, nb, nv;    // direction, first block row / column, blocks, block rows
    int n_items, n_pad;          // candidate columns of one block row; two-row jobs: row 1's columns sit at thread n_pad + i
    int uniform;                 // > 0: every block of the tile has that many column offsets
    int dwin, dref;              // bytes between the 16-byte aligned TMA origin and the wanted one (window / source blocks)
    int pre[F8_NBMAX + 1], dx_lo[F8_NBMAX];
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

// One group of F8_G consecutive row offsets of one candidate column is finished.  The accumulators hold A + 1.  The A
// values go out as an 8-bit MONOTONE code -- code(v) = bits 19..26 of float(2 v): 4 exponent bits (2 <= 2 v < 2^18) + 4
// mantissa bits, i.e. 16 steps per octave -- so the survivor test later compares code(A + 1) with code(A_min + 1 + T): a
// superset of the exact test A <= A_min + T (monotone => nothing is lost; a few extra survivors within one step, 1/16
// of the value, above the threshold).  Half the shared memory of uint16 -- which is what pays for the second stage of
// the window planes -- with the full 16-bit range, and nothing on the integer-ALU pipe the SADs are bound by: IMAD
// (2 v + 2^23-as-float bits), FADD (- 2^23: the classic int-to-float without I2F), IMAD.HI (>> 19), all three on the FMA
// pipe, with constants that come from the parameter block (literals would be turned back into ALU shifts / logic ops).
// Layout A[row offset][thread], A_PITCH = threads + 4 bytes per row: the 32 stores of a warp fall into 8 consecutive
// words, and eight consecutive row offsets of one column fall into eight different banks.  The column minimum stays
// exact.  MASK: the group sticks out of the valid row range (last group only).
struct F8Conv { unsigned two, c8k, magic; float fmagic; };       // 2, 8192, 0x4B000000 and 8388608.0f, opaque to the compiler
__device__ __forceinline__ unsigned f8_code(unsigned v, const F8Conv cv)   // v in [1, 65535]
{
    unsigned t, x;
    float f;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(v), "r"(cv.two), "r"(cv.magic));
    asm("sub.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(t)), "f"(cv.fmagic));
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(x) : "r"(__float_as_uint(f)), "r"(cv.c8k));
    return x;                                                      // the code is the low byte
}
template <bool MASK, int A_PITCH>
__device__ __forceinline__ void f8_finish_group(unsigned (&acc)[F8_G], int n_valid, unsigned char *cg, unsigned &cmin, const F8Conv cv)
{
    if (MASK) {
#pragma unroll
        for (int gi = 0; gi < F8_G; ++gi) acc[gi] = gi < n_valid ? acc[gi] : 0xffffu;
    }
#pragma unroll
    for (int gi = 0; gi < F8_G; ++gi) cg[gi * A_PITCH] = (unsigned char)f8_code(acc[gi], cv);
#pragma unroll
    for (int gi = 0; gi + 1 < F8_G; gi += 2) cmin = min(cmin, min(acc[gi], acc[gi + 1]));     // VIMNMX3
    if (F8_G & 1) cmin = min(cmin, acc[F8_G - 1]);
}

// 8 x 8 source sub-block in rl / rh, F8_G + 7 window rows streamed from the lane's pre-shifted plane: 176 VABSDIFF4.
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

// Persistent.  A JOB is one tile of the host's table (F8Tile): nb adjacent blocks of one block row -- or, where the
// reference's H-for-W column bound (full_search.py:37) leaves a block only 17 of its 33 column offsets at 1080p, of TWO
// block rows searched side by side by the two halves of the CTA -- so every job keeps ~500 of the 512 threads busy and
// all jobs cost the same.  Jobs come off a queue (one global atomic per job, taken by warp 0); a job's target window --
// all four byte shifts, one 3-D TMA load -- and source blocks land in one of `n_stage` shared-memory stages and signal
// bar_full.  Per job j:
//   search the first row-offset group of the thread's candidate column                          -> A[j & 1]
//   hand-over of job j - 1: its block minima were completed a third of a search ago, so the wait on bar_min[(j - 1) & 1]
//     falls through; warp 0 refills the stage job j - 1 vacated (job j + 1 lands while job j is searched); the warp
//     emits job j - 1's survivors
//   search the other groups; column minima -> block minima (shared atomicMin, generation-tagged: never reset);
//   arrive bar_min[j & 1]
// No warp idles for the slowest one, no __syncthreads() after the prologue.  With one stage (wide searches whose window
// does not fit twice) the hand-over of job j happens at its own end.
template <bool M1, int F8_NT>
__global__ void __launch_bounds__(F8_NT, 1024 / F8_NT)
fs8_prefilter_kernel(const __grid_constant__ CUtensorMap tm_win_a, const __grid_constant__ CUtensorMap tm_win_b,
                     const __grid_constant__ CUtensorMap tm_ref_a, const __grid_constant__ CUtensorMap tm_ref_b,
                     const F8Params p)
{
    constexpr int NW = F8_NT / 32;
    constexpr int A_PITCH = F8_NT + 4;
    extern __shared__ __align__(128) unsigned f8_smem[];
    __shared__ __align__(8) uint64_t bar_full[2], bar_min[2];
    __shared__ F8Geo geo_s[2];
    __shared__ unsigned bmin_s[4][2][F8_NBMAX];                     // [job & 3][block row of the tile][block]: (generation << 16) | (A_min + 1)
    __shared__ unsigned seg_count, seg_valid;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bs = p.bs, R = p.R;
    const F8Conv cv = {p.c_two, p.c_8k, p.c_magic, p.c_fmagic};
    const int stage_words = 4 * p.plane_words + ((p.ref_rows * p.ref_rs + 31) & ~31);   // planes [4][plane_words], then source rows [ref_rows][ref_rs]
    unsigned char *A_base = reinterpret_cast<unsigned char *>(f8_smem + p.n_stage * stage_words);   // [2][a_rows][A_PITCH]
    const int a_bytes = p.a_rows * A_PITCH;
    unsigned *seg = p.list + (size_t)blockIdx.x * p.seg_cap;

    if (tid == 0) {
        mbar_init(&bar_full[0], 1); mbar_init(&bar_full[1], 1); mbar_init(&bar_min[0], NW); mbar_init(&bar_min[1], NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        seg_count = 0u; seg_valid = 0xffffffffu;
    }
    if (tid < 4 * 2 * F8_NBMAX) (&bmin_s[0][0][0])[tid] = 0xffffffffu;
    __syncthreads();

    // warp 0: take the next job off the queue, publish its geometry (local job number kk), arm the stage's barrier and
    // issue the two TMA loads
    auto issue_next = [&](int kk) {
        const int sb = p.n_stage == 2 ? (kk & 1) : 0;
        int t = 0;
        if (lane == 0) t = atomicAdd(&p.counters[12], 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        F8Geo *g = &geo_s[kk & 1];
        if (t >= p.n_tiles) {
            if (lane == 0) { g->t = -1; mbar_arrive(&bar_full[sb]); }
            return;
        }
        const F8Tile tile = p.tiles[t];
        const int dir = tile.dir, by = tile.by, bx0 = tile.bx0, nb = tile.nb, nv = tile.nv;
        int n_dx = 0, dxl = 0;
        if (lane < nb) {                                           // column bounds do not depend on the block row
            const FsBounds8 b = f8_bounds(p.H, p.W, bs, R, by * bs, (bx0 + lane) * bs);
            n_dx = max(b.c1 - b.c0, 0);
            dxl = b.c0 - (bx0 + lane) * bs;
        }
        int incl = n_dx;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        g->pre[lane + 1] = incl; g->dx_lo[lane] = dxl;
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int n0 = __shfl_sync(0xffffffffu, n_dx, 0);
        const bool uni = __all_sync(0xffffffffu, lane >= nb || n_dx == n0);
        const int xw = p.box_x0 + bx0 * bs - p.Rp, xr = p.box_x0 + bx0 * bs;          // wanted x origins (multiples of 4)
        if (lane == 0) {
            g->pre[0] = 0; g->t = t; g->dir = dir; g->by = by; g->bx0 = bx0; g->nb = nb; g->nv = nv;
            g->n_items = total; g->n_pad = (total + 31) & ~31;
            g->uniform = (uni && n0 > 0) ? n0 : 0;
            g->dwin = xw & 15; g->dref = xr & 15;
        }
        __syncwarp();
        if (lane == 0) {
            unsigned *st = f8_smem + sb * stage_words;
            mbar_expect_tx(&bar_full[sb], p.tx_bytes);
            tma_load_3d(st, dir ? &tm_win_a : &tm_win_b, xw & ~15, p.box_y0 + by * bs - R, 0, &bar_full[sb]);
            tma_load_3d(st + 4 * p.plane_words, dir ? &tm_ref_b : &tm_ref_a, xr & ~15, p.box_y0 + by * bs, 0, &bar_full[sb]);
        }
    };

    // ---- survivors of one finished job.  A warp takes its flagged columns (column minimum within T of the block minimum)
    //      four at a time: lane = (column slot, row offset mod 8), so one vote covers eight row offsets of all four
    //      columns; the votes stay in registers, the first lane of each slot counts its column's hits and reserves room,
    //      then every lane writes the entry of each of its own hits.  Each CTA appends to its own segment of the list
    //      through a shared-memory cursor that only ever moves forward: a reservation that does not fit marks where the
    //      valid prefix ends (later ones cannot fit either) and its block goes to the redo pass.  fs8_refine_kernel gets
    //      the length of the valid prefix ----
    auto emit = [&](const unsigned char *Abuf, unsigned bmin_word, unsigned head_own, unsigned cmin, bool active, int n_dy) {
        const unsigned thr = active ? (bmin_word & 0xffffu) + (unsigned)p.T : 1u;       // A_min + 1 + T: the accumulators, and so cmin, carry the same + 1
        unsigned flagged = __ballot_sync(0xffffffffu, active && cmin <= thr);
        if (!flagged) return;
        const unsigned thr_code = f8_code(min(thr, 65535u), cv) & 0xffu;
        const int slot = lane >> 3, sub = lane & 7;
        const unsigned my_bits = 0xffu << (slot * 8), below = my_bits & ((1u << lane) - 1u);
        const unsigned char *cw = Abuf + (tid & ~31);
        const int n_it = (n_dy + 7) >> 3;                          // n_dy is the same for every lane of a warp
        while (flagged) {
            int src = -1;
#pragma unroll
            for (int k = 0; k < 4; ++k) {                          // slot k takes the k-th flagged column
                const int sbit = __ffs(flagged) - 1;               // -1 when none is left
                if (slot == k) src = sbit;
                flagged &= flagged - 1;
            }
            const bool on = src >= 0;
            const unsigned thr_c = __shfl_sync(0xffffffffu, thr_code, on ? src : 0);
            const unsigned head = __shfl_sync(0xffffffffu, head_own, on ? src : 0);
            const unsigned char *cs = cw + (on ? src : 0) + sub * A_PITCH;
            // one vote per eight row offsets; up to F8_EMIT_WORDS votes stay in registers (n_dy <= 72: every range up to +-35),
            // longer columns are voted on twice
            unsigned bal[F8_EMIT_WORDS];
            int hits = 0;
#pragma unroll
            for (int m = 0; m < F8_EMIT_WORDS; ++m) {
                bal[m] = 0u;
                if (m < n_it) {
                    const bool h = on && m * 8 + sub < n_dy && cs[m * 8 * A_PITCH] <= thr_c;
                    bal[m] = __ballot_sync(0xffffffffu, h);
                    hits += __popc(bal[m] & my_bits);
                }
            }
            for (int m = F8_EMIT_WORDS; m < n_it; ++m) {
                const bool h = on && m * 8 + sub < n_dy && cs[m * 8 * A_PITCH] <= thr_c;
                hits += __popc(__ballot_sync(0xffffffffu, h) & my_bits);
            }
            int pos = -1;
            if (on && sub == 0) {
                // A column that survives whole is flat content: nothing to prune in this block.  It, and any block whose
                // survivors no longer fit the segment, goes to the exact block-level redo pass instead (once).
                if (hits != n_dy) {
                    pos = (int)atomicAdd(&seg_count, (unsigned)hits);
                    if (pos + hits > p.seg_cap) { atomicMin(&seg_valid, (unsigned)pos); pos = -1; }
                }
                if (pos < 0) {
                    const unsigned bslot = (head >> 31) * (unsigned)(p.nbr * p.nbc) + ((head >> 14) & 0x1ffffu);
                    if (atomicExch(&p.redo_mark[bslot], 0u) != 0u) {
                        const int rp = atomicAdd(&p.counters[11], 1);
                        if (rp < p.redo_cap) p.redo_list[rp] = (int)((head >> 14) & 0x1ffffu) | ((head >> 31) ? 0x40000000 : 0);   // FS_DIR_BIT
                    }
                }
            }
            pos = __shfl_sync(0xffffffffu, pos, slot * 8);
            if (pos >= 0) {                                        // order inside a column is irrelevant to the result
#pragma unroll
                for (int m = 0; m < F8_EMIT_WORDS; ++m) {
                    if (m < n_it) {
                        if ((bal[m] >> lane) & 1u) seg[pos + __popc(bal[m] & below)] = head | ((unsigned)(m * 8 + sub) << 7);
                        pos += __popc(bal[m] & my_bits);
                    }
                }
            }
            if (n_it > F8_EMIT_WORDS && __any_sync(0xffffffffu, pos >= 0)) {
                for (int m = F8_EMIT_WORDS; m < n_it; ++m) {
                    const int i = m * 8 + sub;
                    const bool h = pos >= 0 && i < n_dy && cs[m * 8 * A_PITCH] <= thr_c;
                    const unsigned b2 = __ballot_sync(0xffffffffu, h);
                    if (h) seg[pos + __popc(b2 & below)] = head | ((unsigned)i << 7);
                    pos += __popc(b2 & my_bits);
                }
            }
        }
    };

    if (warp == 0) { issue_next(0); if (p.n_stage == 2) issue_next(1); }
    int n_issued = p.n_stage;

    // the job whose survivors are still to be emitted
    bool pv_valid = false, pv_active = false;
    unsigned pv_head = 0u, pv_cmin = 0u;
    int pv_ndy = 0, pv_bslot = 0;
    int j = 0;                                                     // jobs searched so far

    auto flush_prev = [&](bool refill, bool wait) {                 // hand-over of job j - 1 (see above)
        const int q = j - 1;
        if (wait) mbar_wait_sleep(&bar_min[q & 1], (unsigned)((q >> 1) & 1));
        if (refill) {                                              // every warp is done with that job's stage
            if (warp == 0) issue_next(n_issued);
            ++n_issued;
        }
        if (!(p.ablate & 1)) emit(A_base + (q & 1) * a_bytes, (&bmin_s[0][0][0])[pv_bslot], pv_head, pv_cmin, pv_active, pv_ndy);
        pv_valid = false;
    };

#pragma unroll 1
    for (;; ++j) {
        const int sb = p.n_stage == 2 ? (j & 1) : 0;
        mbar_wait_sleep(&bar_full[sb], (unsigned)((p.n_stage == 2 ? (j >> 1) : j) & 1));
        const F8Geo *g = &geo_s[j & 1];
        if (g->t < 0) break;
        const int dir = g->dir, by = g->by, bx0 = g->bx0, nb = g->nb, n_items = g->n_items;
        const unsigned *planes = f8_smem + sb * stage_words;
        const unsigned *ref0 = planes + 4 * p.plane_words;
        // the thread's candidate column (block, dx); in a two-row job the upper half of the CTA takes block row 1
        int it = tid, v = 0;
        if (g->nv == 2 && tid >= g->n_pad) { it = tid - g->n_pad; v = 1; }
        const bool has_item = it < n_items;
        int b = 0, dxi = 0, dx = 0;
        if (has_item) {
            if (g->uniform > 0) b = it / g->uniform;
            else while (b + 1 < nb && g->pre[b + 1] <= it) ++b;
            dxi = it - g->pre[b];
            dx = g->dx_lo[b] + dxi;
        }
        const int wq = b * bs + dx + p.Rp + g->dwin;                 // byte offset of the column inside a staged window row
        const unsigned *pl = planes + (wq & 3) * p.plane_words + (wq >> 2);
        const int s_row = (by + v) * bs;
        const unsigned *ref = ref0 + (v * bs) * p.ref_rs + ((b * bs + g->dref) >> 2);
        const FsBounds8 b0 = f8_bounds(p.H, p.W, bs, R, s_row, bx0 * bs);
        const int n_dy = max(b0.r1 - b0.r0, 0);
        const int wrow_base = b0.r0 - (by * bs - R);                   // window rows count from the job's first block row
        const bool active = has_item && n_dy > 0;
        unsigned char *col = A_base + (j & 1) * a_bytes + tid;
        unsigned cmin = 0xffffffffu;
        // Lanes without a column of their own (the tail of the last warp) search block 0 / dx 0 along with their warp: the
        // hand-over below contains warp-wide votes and shuffles, so control flow stays warp-uniform; their results are ignored.
        if (__any_sync(0xffffffffu, active)) {
            unsigned rl[8], rh[8];
            if (M1) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint2 rv = *reinterpret_cast<const uint2 *>(ref + kk * p.ref_rs);
                    rl[kk] = rv.x; rh[kk] = rv.y;
                }
            }
            // full groups first, then (clipped windows only) the one group that sticks out of the valid row range: two
            // separate copies of the loop body, so that the common one carries no masking selects
            auto group = [&](int gq, auto masked) {
                unsigned acc[F8_G];
#pragma unroll
                for (int i = 0; i < F8_G; ++i) acc[i] = 1u;                  // A + 1: see f8_finish_group
                if (M1) {
                    f8_accumulate(acc, rl, rh, pl + (wrow_base + gq * F8_G) * p.rs, p.rs);
                } else {
                    const int m = p.m;
#pragma unroll 1
                    for (int sub = 0; sub < m * m; ++sub) {
                        const int si = sub / m, sj = sub - si * m;
                        const unsigned *rp = ref + (si * 8) * p.ref_rs + sj * 2;
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            const uint2 rv = *reinterpret_cast<const uint2 *>(rp + kk * p.ref_rs);
                            rl[kk] = rv.x; rh[kk] = rv.y;
                        }
                        f8_accumulate(acc, rl, rh, pl + (wrow_base + gq * F8_G + si * 8) * p.rs + sj * 2, p.rs);
                    }
                }
                f8_finish_group<decltype(masked)::value, A_PITCH>(acc, n_dy - gq * F8_G, col + gq * F8_G * A_PITCH, cmin, cv);
            };
            const int n_full = n_dy / F8_G;
#pragma unroll 1
            for (int gq = 0; gq < n_full; ++gq) {
                group(gq, std::false_type{});
                // hand-over of the previous job as soon as its block minima are complete -- probed, not waited for: a warp that
                // runs ahead of its CTA keeps searching, so the warps of a CTA do not all emit at the same moment
                if (pv_valid && gq >= p.flush_at) {
                    const int q = j - 1;
                    if (__shfl_sync(0xffffffffu, (int)mbar_test(&bar_min[q & 1], (unsigned)((q >> 1) & 1)), 0)) flush_prev(true, false);
                }
            }
            if (n_full * F8_G < n_dy) group(n_full, std::true_type{});
        }
        if (pv_valid) flush_prev(true, true);                       // not complete at any probe (or no full group in this job): wait
        // block minima: segmented warp reduction, one shared-memory atomic per (warp, block).  The generation tag in the
        // upper half makes every later job's entries smaller than any earlier job's, so the slots are never reset (a slot
        // is re-used four jobs later; its last reader finished before the hand-over two jobs earlier).
        const int bslot = ((j & 3) * 2 + v) * F8_NBMAX;
        {
            const unsigned tag = (0xffffu - (unsigned)(j >> 2)) << 16;
            unsigned todo = __ballot_sync(0xffffffffu, active);
            while (todo) {
                const int leader = __ffs(todo) - 1;
                const int bb = __shfl_sync(0xffffffffu, b, leader);
                const bool mine = active && b == bb;
                const unsigned mn = __reduce_min_sync(0xffffffffu, mine ? cmin : 0xffffffffu);
                if (lane == leader) atomicMin(&(&bmin_s[0][0][0])[bslot + bb], tag | mn);
                todo &= ~__ballot_sync(0xffffffffu, mine);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_min[j & 1]);
        // this job becomes the pending one
        pv_valid = true; pv_active = active; pv_cmin = cmin; pv_ndy = n_dy; pv_bslot = bslot + b;
        pv_head = (dir ? 0x80000000u : 0u) | ((unsigned)((by + v) * p.nbc + bx0 + b) << 14) | (unsigned)dxi;
        if (p.n_stage == 1) { ++j; flush_prev(true, true); --j; }        // one stage: the next job's loads cannot start before this hand-over
    }
    if (pv_valid) flush_prev(false, true);
    __syncthreads();
    if (tid == 0) {
        const unsigned n = seg_valid != 0xffffffffu ? seg_valid : seg_count;     // the prefix every entry of which was written
        p.seg_counts[blockIdx.x] = n;
        atomicAdd(&p.counters[9], (int)n);
    }
}

__device__ __forceinline__ unsigned long long f8_key(unsigned sad, unsigned d2, unsigned ri, unsigned ci)
{
    return ((unsigned long long)sad << 32) | ((unsigned long long)d2 << 16) | (ri << 8) | ci;    // = fs.cu fs_key
}

// Exact SAD of one survivor by eight lanes: lane pr covers row pr of each 8 x 8 sub-block -- the source row as two
// aligned 16-byte vectors, the target row as one 32-byte run; returns the lane's partial sum of |dL1000|.
// (Measured alternatives, both slower at 8 x 8: a whole warp per survivor with lane = (row, pixel pair), 102 vs 76 us at
// 1080p -- the per-survivor decode and hand-over then run 32 instead of 8 lanes wide -- and the target row as the three
// aligned 16-byte vectors that cover it, 83 us: half as many load instructions, 1.5 x the bytes.)
__device__ __forceinline__ unsigned f8_survivor_sad(const int32_t *__restrict__ Ls, const int32_t *__restrict__ Lt, int H, int W, int Wp,
                                                    int bs, int m, int s_row, int s_col, int t_row, int t_col, int pr)
{
    unsigned S = 0;
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
    return S;
}

struct F8Entry { int dir, blk, dyi, dxi, s_row, s_col, t_row, t_col; };
__device__ __forceinline__ F8Entry f8_decode(unsigned e, int nbc, float inv_nbc, int bs, int R)
{
    F8Entry x;
    x.dir = e >> 31; x.blk = (int)((e >> 14) & 0x1ffffu); x.dyi = (int)((e >> 7) & 127u); x.dxi = (int)(e & 127u);
    int br = (int)(((float)x.blk + 0.5f) * inv_nbc);          // blk / nbc (blk < 2^17: the float quotient is off by at most one)
    br -= (br * nbc > x.blk); br += ((br + 1) * nbc <= x.blk);
    const int bc = x.blk - br * nbc;
    x.s_row = br * bs; x.s_col = bc * bs;
    x.t_row = max(x.s_row - R, 0) + x.dyi; x.t_col = max(x.s_col - R, 0) + x.dxi;      // window origin (r0, c0) of f8_bounds
    return x;
}

// Every CTA takes one contiguous range of the survivor list (the concatenation of the per-CTA segments of the
// prefilter); its 32 eight-lane groups walk that range side by side, so neighbouring groups take neighbouring survivors
// (same block, same column offset: the same source rows and overlapping target rows -- loads of the same line coalesce)
// and a group finds its segment by stepping forward, not by searching.  A flagged sum (S a positive multiple of 1000: the
// reference's float64 accumulation may land just below the integer and truncate to S / 1000 - 1, fs.cu) is NOT decided
// here: the survivor goes on a short list for fs8_replay_kernel, which keeps this kernel at 32 registers and the SM full.
// Blocks already on the redo list (dense blocks) still get their keys; the redo pass overwrites them.
__global__ void __launch_bounds__(256, 8)
fs8_refine_kernel(const int32_t *__restrict__ L0, const int32_t *__restrict__ L1,
                  int H, int W, int Wp, int bs, int R, int nbc, int n_blk,
                  const unsigned *__restrict__ list, int seg_cap, const unsigned *__restrict__ seg_counts, int n_seg,
                  int *__restrict__ counters, unsigned long long *__restrict__ best,
                  unsigned *__restrict__ flag_list, int flag_cap,
                  unsigned *__restrict__ redo_mark, int *__restrict__ redo_list, int redo_cap)
{
    if (counters[8]) return;                                       // overflow: the exact kernel takes the pair
    // exclusive prefix of the segment lengths (n_seg <= F8_MAXSEG, every CTA computes it for itself)
    __shared__ int seg_pre[F8_MAXSEG + 1];
    __shared__ int lo_s;
    // Minimum keys are combined per CTA before they touch global memory: a CTA's range is contiguous, i.e. a handful of
    // blocks, and a block whose content is nearly flat can own a thousand survivors -- a thousand atomics on ONE address
    // would serialise in L2 and set the duration of the whole kernel.  Direct-mapped, 128 (direction, block) slots; a
    // survivor whose slot belongs to another block goes to global memory directly.
    __shared__ unsigned long long ckey[128];
    __shared__ int cid[128];
    if (threadIdx.x < 128) { ckey[threadIdx.x] = 0xffffffffffffffffull; cid[threadIdx.x] = -1; }
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
    const int per_cta = (((n + gridDim.x - 1) / gridDim.x) + 31) & ~31;
    const int cs = blockIdx.x * per_cta, ce = min(n, cs + per_cta);
    if (cs >= ce) return;
    if (threadIdx.x == 0) {                                        // segment of the range's first survivor
        int lo = 0, hi = n_seg;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seg_pre[mid] <= cs) lo = mid; else hi = mid; }
        lo_s = lo;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, pr = lane & 7, grp = threadIdx.x >> 3;
    const int m = bs >> 3;
    const float inv_nbc = 1.0f / (float)nbc;
    int lo = lo_s;
    for (int i0 = cs; i0 < ce; i0 += 32) {                         // warp-uniform trip count (shuffles inside)
        const int i = i0 + grp;
        const bool on = i < ce;
        unsigned e = 0u;
        if (on) {
            while (seg_pre[lo + 1] <= i) ++lo;                     // empty segments are stepped over
            e = list[(size_t)lo * seg_cap + (i - seg_pre[lo])];
        }
        const F8Entry x = f8_decode(e, nbc, inv_nbc, bs, R);
        unsigned S = on ? f8_survivor_sad(x.dir ? L1 : L0, x.dir ? L0 : L1, H, W, Wp, bs, m, x.s_row, x.s_col, x.t_row, x.t_col, pr) : 0u;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
        if (on && pr == 0) {
            const unsigned sad = S / 1000u;
            bool decided = true;
            if (S != 0u && sad * 1000u == S) {                     // flagged: decided by the replay pass
                const int fp = atomicAdd(&counters[13], 1);
                if (fp < flag_cap) { flag_list[fp] = e; decided = false; }
                else if (atomicExch(&redo_mark[(size_t)x.dir * n_blk + x.blk], 0u) != 0u) {      // list full: the block is searched exactly
                    const int pos = atomicAdd(&counters[11], 1);
                    if (pos < redo_cap) redo_list[pos] = x.blk | (x.dir ? 0x40000000 : 0);       // FS_DIR_BIT
                }
            }
            if (decided) {
                const int dr = x.s_row - x.t_row, dc = x.s_col - x.t_col;
                const unsigned long long key = f8_key(sad, (unsigned)(dr * dr + dc * dc), (unsigned)x.dyi, (unsigned)x.dxi);
                const int id = x.dir * n_blk + x.blk, h = id & 127;
                const int owner = atomicCAS(&cid[h], -1, id);
                if (owner == -1 || owner == id) atomicMin(&ckey[h], key);
                else atomicMin(&best[id], key);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 128 && cid[threadIdx.x] >= 0) atomicMin(&best[cid[threadIdx.x]], ckey[threadIdx.x]);
}

// Flagged survivors (a few hundred per 1080p pair): 8 x 8 blocks are replayed in float64 in the reference's operation
// order -- each lane forms the |dL| terms of its row, then the 64 terms are added row by row, column by column -- larger
// blocks go to the block-level redo pass (their integer SAD is entered meanwhile; the redo pass overwrites the block).
__global__ void __launch_bounds__(256)
fs8_replay_kernel(const int32_t *__restrict__ L0, const int32_t *__restrict__ L1,
                  const uint8_t *__restrict__ rgb0, const uint8_t *__restrict__ rgb1,
                  int H, int W, int Wp, int bs, int R, int nbc, int n_blk,
                  const unsigned *__restrict__ flag_list, int flag_cap, int *__restrict__ counters,
                  unsigned long long *__restrict__ best, unsigned *__restrict__ redo_mark,
                  int *__restrict__ redo_list, int redo_cap)
{
    if (counters[8]) return;
    const int n = min(counters[13], flag_cap);
    const int lane = threadIdx.x & 31, pr = lane & 7;
    const int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, n_grp = (gridDim.x * blockDim.x) >> 3;
    const int m = bs >> 3;
    const float inv_nbc = 1.0f / (float)nbc;
    for (int i0 = 0; i0 < n; i0 += n_grp) {                        // warp-uniform trip count (shuffles inside)
        const int i = i0 + grp;
        const bool on = i < n;
        const unsigned e = on ? flag_list[i] : 0u;
        const F8Entry x = f8_decode(e, nbc, inv_nbc, bs, R);
        unsigned sad = 0u;
        if (bs == 8) {
            double tr[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) tr[c] = 0.0;
            if (on) {
                const uint8_t *rgb_s = x.dir ? rgb1 : rgb0, *rgb_t = x.dir ? rgb0 : rgb1;
                const int r1 = x.s_row + pr, r2 = x.t_row + pr;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int c1 = x.s_col + c, c2 = x.t_col + c;
                    double l1 = 0.0, l2 = 0.0;
                    if (r1 < H && c1 < W) { const uint8_t *q = rgb_s + ((size_t)r1 * W + c1) * 3; l1 = luma_f64(q[0], q[1], q[2]); }
                    if (r2 < H && c2 < W) { const uint8_t *q = rgb_t + ((size_t)r2 * W + c2) * 3; l2 = luma_f64(q[0], q[1], q[2]); }
                    tr[c] = fabs(__dsub_rn(l1, l2));
                }
            }
            double acc = 0.0;
#pragma unroll 1
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int c = 0; c < 8; ++c) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, tr[c], r, 8));
            }
            sad = (unsigned)acc;
            if (on && pr == 0) atomicAdd(&counters[10], 1);
        } else {
            unsigned S = on ? f8_survivor_sad(x.dir ? L1 : L0, x.dir ? L0 : L1, H, W, Wp, bs, m, x.s_row, x.s_col, x.t_row, x.t_col, pr) : 0u;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
            sad = S / 1000u;
            if (on && pr == 0 && atomicExch(&redo_mark[(size_t)x.dir * n_blk + x.blk], 0u) != 0u) {
                const int pos = atomicAdd(&counters[11], 1);
                if (pos < redo_cap) redo_list[pos] = x.blk | (x.dir ? 0x40000000 : 0);           // FS_DIR_BIT
            }
        }
        if (on && pr == 0) {
            const int dr = x.s_row - x.t_row, dc = x.s_col - x.t_col;
            atomicMin(&best[(size_t)x.dir * n_blk + x.blk], f8_key(sad, (unsigned)(dr * dr + dc * dc), (unsigned)x.dyi, (unsigned)x.dxi));
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

// ---- rounded 8-bit luma, four byte-shifted planes with a zero border (np.pad of the reference and every out-of-frame
//      read): plane q, padded row y, padded column x holds L8[y][x + q].  One thread per word of a row, plus one word to
//      the left of the frame (planes 1..3 carry the first q pixels of a row there). ----
__global__ void luma8_kernel(const int32_t *__restrict__ L, int H, int W, int Wp, uint8_t *__restrict__ l8, int pitch8, size_t plane_bytes)
{
    const int quads = (Wp >> 2) + 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * quads) return;
    const int y = idx / quads, x = ((idx - y * quads) << 2) - 4;                    // x = -4, 0, 4, ...
    auto pack = [&](int xx) -> unsigned {
        if (xx < 0 || xx >= Wp) return 0u;
        const int4 v = *reinterpret_cast<const int4 *>(L + (size_t)y * Wp + xx);     // pad columns of L are 0
        return (unsigned)((v.x + 500) / 1000) | ((unsigned)((v.y + 500) / 1000) << 8) |
               ((unsigned)((v.z + 500) / 1000) << 16) | ((unsigned)((v.w + 500) / 1000) << 24);
    };
    const unsigned w0 = pack(x), w1 = pack(x + 4);
    uint8_t *dst = l8 + (size_t)(MEMCI_L8_PAD + y) * pitch8 + MEMCI_L8_PAD + x;
    *reinterpret_cast<unsigned *>(dst) = w0;
    *reinterpret_cast<unsigned *>(dst + plane_bytes) = __funnelshift_r(w0, w1, 8);
    *reinterpret_cast<unsigned *>(dst + 2 * plane_bytes) = __funnelshift_r(w0, w1, 16);
    *reinterpret_cast<unsigned *>(dst + 3 * plane_bytes) = __funnelshift_r(w0, w1, 24);
}

static int f8_pitch8(const memci_ctx *ctx) { return (ctx->Wp + 2 * MEMCI_L8_PAD + 63) & ~63; }
static size_t f8_plane_bytes(const memci_ctx *ctx) { return ((size_t)(ctx->H + 2 * MEMCI_L8_PAD) * f8_pitch8(ctx) + 255) & ~(size_t)255; }

int memci_ensure_luma8(memci_ctx *ctx, int slot)
{
    FrameSlot &f = ctx->frames[slot];
    int rc = memci_ensure_luma(ctx, slot);
    if (rc) return rc;
    if (f.luma8_ok) return MEMCI_OK;
    const int pitch8 = f8_pitch8(ctx);
    const size_t plane_bytes = f8_plane_bytes(ctx);
    if (!f.luma8) {
        CK(ctx, cudaMalloc(&f.luma8, 4 * plane_bytes));
        CK(ctx, cudaMemsetAsync(f.luma8, 0, 4 * plane_bytes, ctx->stream));
    }
    const int n = ctx->H * ((ctx->Wp >> 2) + 1);
    luma8_kernel<<<ceil_div_i(n, 256), 256, 0, ctx->stream>>>(f.luma, ctx->H, ctx->W, ctx->Wp, f.luma8, pitch8, plane_bytes);
    CKL(ctx);
    f.luma8_ok = 1;
    return MEMCI_OK;
}

// ---- launch shape ----------------------------------------------------------------------------------------------
struct F8Shape {
    int nt, ctas, rpt, n_stage, nb_tile;
    int win_rows, box_rows, box_w, rs, plane_words, ref_w, ref_rs, a_rows, smem;
};

// Shared memory of one CTA and the TMA boxes for (threads, block rows per tile, stages).  Returns false when the
// shape is impossible (box wider than 256 bytes / taller than 256 rows, no column per thread).
static bool f8_layout(int nt, int rpt, int n_stage, int bs, int R, int nbc, F8Shape *s)
{
    const int Rp = (R + 3) & ~3;
    const int n_dy = 2 * R + 1;
    const int n_groups = (n_dy + F8_G - 1) / F8_G;
    int nb_tile = nt / (2 * R + 1);
    if (nb_tile > F8_NBMAX) nb_tile = F8_NBMAX;
    if (nb_tile > nbc) nb_tile = nbc;
    const int win_rows = rpt * bs + 2 * R + (n_groups * F8_G - n_dy) + 1;
    int best_bytes = 0;
    for (; nb_tile >= 1 && !best_bytes; --nb_tile) {
        // bytes of a staged row a thread may touch: column offset + 12 bytes of TMA alignment slack + second word of the pair
        const int need_w = 4 * (((nb_tile * bs + R + Rp + 4) >> 2) + 2);
        for (int bw = (need_w + 15) & ~15; bw <= 256; bw += 16) {
            for (int e = 0; e < 8 && win_rows + e <= 256; ++e) {
                const int pw = (bw / 4) * (win_rows + e);          // words per plane: the four planes must start 8 banks apart
                if ((pw & 31) != 8 && (pw & 31) != 24) continue;
                if (!best_bytes || pw * 4 < best_bytes) {
                    best_bytes = pw * 4;
                    s->box_w = bw; s->box_rows = win_rows + e; s->rs = bw / 4; s->plane_words = pw; s->nb_tile = nb_tile;
                }
                break;
            }
        }
    }
    if (!best_bytes) return false;
    s->nt = nt; s->rpt = rpt; s->n_stage = n_stage; s->win_rows = win_rows;
    s->ref_w = (s->nb_tile * bs + 8 + 15) & ~15;
    s->ref_rs = s->ref_w / 4;
    s->a_rows = n_groups * F8_G;
    const int ref_words = (rpt * bs * s->ref_rs + 31) & ~31;
    s->smem = n_stage * (4 * s->plane_words + ref_words) * 4 + 2 * s->a_rows * (nt + 4);      // two A buffers (uint8 codes): jobs j and j - 1
    return true;
}

// Preference: 512 threads x 2 CTAs per SM, two block rows per tile, two stages; fewer stages / one CTA per SM / one block
// row / 256 threads as shared memory dictates (wide searches: 16 x 16 +-32 runs one 512-thread CTA per SM).
// MEMCI_F8_NT / MEMCI_F8_RPT / MEMCI_F8_STAGES / MEMCI_F8_CTAS override (tuning).
static bool f8_shape(int bs, int R, int nbc, F8Shape *out)
{
    static const int cand[][4] = {      // threads, CTAs per SM, block rows per tile, stages
        {512, 2, 2, 2}, {512, 2, 2, 1}, {512, 1, 2, 2}, {512, 1, 2, 1}, {512, 2, 1, 2}, {512, 2, 1, 1}, {256, 2, 2, 2}, {256, 2, 2, 1},
        {256, 2, 1, 2}, {256, 2, 1, 1}, {512, 1, 1, 2}, {512, 1, 1, 1}, {256, 1, 1, 1}};
    int f_nt = 0, f_rpt = 0, f_st = 0, f_ctas = 0;
    if (const char *e = getenv("MEMCI_F8_NT")) f_nt = atoi(e);
    if (const char *e = getenv("MEMCI_F8_RPT")) f_rpt = atoi(e);
    if (const char *e = getenv("MEMCI_F8_STAGES")) f_st = atoi(e);
    if (const char *e = getenv("MEMCI_F8_CTAS")) f_ctas = atoi(e);
    const bool rpt2_ok = R + 2 * bs + 16 <= MEMCI_L8_PAD;
    for (const auto &c : cand) {
        if ((f_nt && c[0] != f_nt) || (f_ctas && c[1] != f_ctas) || (f_rpt && c[2] != f_rpt) || (f_st && c[3] != f_st)) continue;
        if (c[2] == 2 && !rpt2_ok) continue;
        F8Shape s;
        if (!f8_layout(c[0], c[2], c[3], bs, R, nbc, &s)) continue;
        const int budget = (c[1] == 2 ? 113 : 226) * 1024 - 2560;   // 228 KB per SM, 1 KB reserved per CTA, static shared memory
        if (s.smem > budget) continue;
        s.ctas = c[1];
        *out = s;
        return true;
    }
    return false;
}

// Is the prefilter path applicable?  Block sizes 8 / 16 (A fits uint16), list entry fields (17-bit block, 7-bit offsets),
// border of the padded planes, a launch shape that fits shared memory.
int memci_fs8_eligible(memci_ctx *ctx, int bs, int R)
{
    if ((bs != 8 && bs != 16) || R < 1 || R > 63 || R + bs + 16 > MEMCI_L8_PAD) return 0;
    const long long n_blk = (long long)ceil_div_i(ctx->H, bs) * ceil_div_i(ctx->W, bs);
    if (n_blk >= (1 << 17)) return 0;
    F8Shape s;
    return f8_shape(bs, R, ceil_div_i(ctx->W, bs), &s) ? 1 : 0;
}

template <bool M1, int NT>
static int f8_launch(memci_ctx *ctx, const CUtensorMap *tm, const F8Params &p, int grid, int smem)
{
    CK(ctx, cudaFuncSetAttribute(fs8_prefilter_kernel<M1, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    fs8_prefilter_kernel<M1, NT><<<grid, NT, smem, ctx->stream>>>(tm[0], tm[1], tm[2], tm[3], p);
    return MEMCI_OK;
}

// The jobs of one search geometry, in queue order (cached per context; rebuilt when the geometry changes): row-major over
// (direction, block row, tile column).  A tile column whose blocks keep so few column offsets that two block rows fit the
// CTA side by side becomes ONE two-row job per pair of block rows; every other tile is one job per block row.
static int f8_job_table(memci_ctx *ctx, const F8Shape &sh, int bs, int R, int br0, int br1, int ndir, const F8Tile **tiles, int *n_tiles)
{
    const int H = ctx->H, W = ctx->W, nbc = ceil_div_i(W, bs);
    const int no_split = getenv("MEMCI_F8_NOSPLIT") ? 1 : 0;
    const int key[8] = {bs, R, br0, br1, ndir, sh.nb_tile, sh.nt * 4 + sh.rpt * 2 + no_split, 1};
    if (ctx->d_f8_tiles && memcmp(key, ctx->f8_tiles_key, sizeof(key)) == 0) {
        *tiles = reinterpret_cast<const F8Tile *>(ctx->d_f8_tiles); *n_tiles = ctx->f8_tiles_n;
        return MEMCI_OK;
    }
    const int n_cols = ceil_div_i(nbc, sh.nb_tile);
    F8Tile *h = (F8Tile *)malloc(sizeof(F8Tile) * (size_t)ndir * (br1 - br0) * n_cols);
    if (!h) return memci_fail(ctx, MEMCI_ERR_ARG, "out of host memory");
    unsigned char *two = (unsigned char *)calloc(n_cols, 1);
    for (int tc = 0; tc < n_cols; ++tc) {                          // column bounds do not depend on the block row
        const int bx0 = tc * sh.nb_tile, nb = nbc - bx0 < sh.nb_tile ? nbc - bx0 : sh.nb_tile;
        int items = 0;
        for (int b = 0; b < nb; ++b) {
            const int s_col = (bx0 + b) * bs;
            const int tmc = (s_col + bs >= H) ? s_col + 1 : W - bs;        // full_search.py:37 (sic: compares against H)
            const int c0 = s_col - R > 0 ? s_col - R : 0, c1 = tmc < s_col + R + 1 ? tmc : s_col + R + 1;
            items += c1 > c0 ? c1 - c0 : 0;
        }
        two[tc] = sh.rpt == 2 && !no_split && items > 0 && ((items + 31) & ~31) + items <= sh.nt;
    }
    int n = 0;
    for (int dir = 0; dir < ndir; ++dir)
        for (int by = br0; by < br1; ++by)
            for (int tc = 0; tc < n_cols; ++tc) {
                const int bx0 = tc * sh.nb_tile, nb = nbc - bx0 < sh.nb_tile ? nbc - bx0 : sh.nb_tile;
                if (two[tc] && ((by - br0) & 1)) continue;          // second row of a two-row job
                F8Tile t;
                t.dir = (short)dir; t.nv = (short)((two[tc] && by + 1 < br1) ? 2 : 1); t.by = (short)by; t.bx0 = (short)bx0; t.nb = (short)nb; t.pad = 0;
                h[n++] = t;
            }
    free(two);
    if (ctx->d_f8_tiles) { cudaFree(ctx->d_f8_tiles); ctx->d_f8_tiles = nullptr; }
    cudaError_t e = cudaMalloc(&ctx->d_f8_tiles, sizeof(F8Tile) * (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_f8_tiles, h, sizeof(F8Tile) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    free(h);
    if (e != cudaSuccess) { ctx->sticky = 1; return memci_fail(ctx, MEMCI_ERR_CUDA, "prefilter job table: %s", cudaGetErrorString(e)); }
    memcpy(ctx->f8_tiles_key, key, sizeof(key));
    ctx->f8_tiles_n = n;
    *tiles = reinterpret_cast<const F8Tile *>(ctx->d_f8_tiles); *n_tiles = n;
    return MEMCI_OK;
}

// Enqueue prefilter + refine + output for block rows [br0, br1).  counters[8] is left non-zero when the pair has to be
// searched by the exact kernel; the caller runs that kernel gated on the flag.
int memci_fs8_run(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int R, MVField *out, int br0, int br1, MVField *out2)
{
    const int H = ctx->H, W = ctx->W;
    int rc;
    if ((rc = memci_ensure_luma8(ctx, slot_src))) return rc;
    if ((rc = memci_ensure_luma8(ctx, slot_tgt))) return rc;
    const int nbr = ceil_div_i(H, bs), nbc = ceil_div_i(W, bs), n_blk = nbr * nbc, ndir = out2 ? 2 : 1;
    F8Shape sh;
    if (!f8_shape(bs, R, nbc, &sh)) return memci_fail(ctx, MEMCI_ERR_ARG, "two-phase full search: no launch shape for block %d, range %d", bs, R);
    F8Params p;
    memset(&p, 0, sizeof(p));
    p.H = H; p.W = W; p.bs = bs; p.m = bs / 8; p.R = R; p.Rp = (R + 3) & ~3; p.nbr = nbr; p.nbc = nbc;
    p.n_stage = sh.n_stage;
    p.c_two = 2u; p.c_8k = 8192u; p.c_magic = 0x4B000000u; p.c_fmagic = 8388608.0f;
    p.flush_at = 0;
    if (const char *e = getenv("MEMCI_F8_ABLATE")) p.ablate = atoi(e);
    if (const char *e = getenv("MEMCI_F8_FLUSH_AT")) p.flush_at = atoi(e);
    if ((rc = f8_job_table(ctx, sh, bs, R, br0, br1, ndir, &p.tiles, &p.n_tiles))) return rc;
    p.box_x0 = MEMCI_L8_PAD; p.box_y0 = MEMCI_L8_PAD;
    p.rs = sh.rs; p.plane_words = sh.plane_words; p.ref_rs = sh.ref_rs; p.ref_rows = sh.rpt * bs; p.a_rows = sh.a_rows;
    p.tx_bytes = (unsigned)(4 * sh.plane_words * 4 + sh.ref_w * sh.rpt * bs);
    p.T = 2 * bs * bs + 1;
    const FrameSlot &fs = ctx->frames[slot_src], &ft = ctx->frames[slot_tgt];
    // tensor maps: the four planes of a frame as one 3-D tensor {pitch8, H + 2 PAD, 4}; window box = all four planes,
    // source-block box = plane 0 only
    CUtensorMap tm[4];
    {
        const unsigned long long dims[3] = {(unsigned long long)f8_pitch8(ctx), (unsigned long long)(H + 2 * MEMCI_L8_PAD), 4ull};
        const unsigned long long strides[2] = {(unsigned long long)f8_pitch8(ctx), (unsigned long long)f8_plane_bytes(ctx)};
        const unsigned box_win[3] = {(unsigned)sh.box_w, (unsigned)sh.box_rows, 4u};
        const unsigned box_ref[3] = {(unsigned)sh.ref_w, (unsigned)(sh.rpt * bs), 1u};
        if ((rc = memci_tmap_encode(ctx, &tm[0], 1, 3, fs.luma8, dims, strides, box_win))) return rc;
        if ((rc = memci_tmap_encode(ctx, &tm[1], 1, 3, ft.luma8, dims, strides, box_win))) return rc;
        if ((rc = memci_tmap_encode(ctx, &tm[2], 1, 3, fs.luma8, dims, strides, box_ref))) return rc;
        if ((rc = memci_tmap_encode(ctx, &tm[3], 1, 3, ft.luma8, dims, strides, box_ref))) return rc;
    }
    // survivor list: an eighth of the candidates, at most 16 M entries (64 MB).  Textured content needs about 0.5 %, but the
    // list is cut into one segment per prefilter CTA and flat regions load a few segments heavily; a segment that fills
    // up sends its blocks to the redo pass, so the size is a performance knob, not a correctness one.
    const size_t cand_max = (size_t)n_blk * (2 * R + 1) * (2 * R + 1) * 2;
    size_t cap = cand_max / 8 < (1u << 20) ? (1u << 20) : cand_max / 8;
    if (cap > (16u << 20)) cap = 16u << 20;
    size_t have = ctx->fs8_list_cap;
    if ((rc = memci_reserve(ctx, (void **)&ctx->d_fs8_list, &have, (cap + F8_MAXSEG) * 4))) return rc;
    if (have != ctx->fs8_list_cap) CK(ctx, cudaMemsetAsync(ctx->d_fs8_list, 0, have, ctx->stream));     // new buffer: no stale entries, ever
    ctx->fs8_list_cap = have;
    size_t have_b = ctx->fs8_best_cap;
    if ((rc = memci_reserve(ctx, (void **)&ctx->d_fs8_best, &have_b, (size_t)n_blk * 2 * 12 + F8_FLAG_CAP * 4))) return rc;       // keys + redo marks + flagged survivors
    if (have_b != ctx->fs8_best_cap) CK(ctx, cudaMemsetAsync(ctx->d_fs8_best, 0xff, have_b, ctx->stream));    // new buffer: armed once, then by fs8_output_kernel
    ctx->fs8_best_cap = have_b;
    p.counters = ctx->d_counters;
    p.redo_mark = reinterpret_cast<unsigned *>(ctx->d_fs8_best + (size_t)n_blk * 2);
    p.redo_list = ctx->d_redo_list; p.redo_cap = ctx->redo_cap; p.redo_limit = n_blk * ndir / 16 + 64;
    CK(ctx, cudaMemsetAsync(ctx->d_counters + 8, 0, 6 * sizeof(int), ctx->stream));   // [8] overflow [9] survivors [10] float64 replays [11] redo blocks [12] tile queue [13] flagged survivors
    const int n_tiles = p.n_tiles;
    int grid = n_tiles < sh.ctas * ctx->num_sms ? n_tiles : sh.ctas * ctx->num_sms;
    if (grid > F8_MAXSEG) grid = F8_MAXSEG;
    p.seg_counts = ctx->d_fs8_list;                                // [F8_MAXSEG] segment lengths, then the segments
    p.list = ctx->d_fs8_list + F8_MAXSEG;
    p.seg_cap = (int)((have / 4 - F8_MAXSEG) / grid);
    if (const char *e = getenv("MEMCI_F8_SEG_CAP")) { const int v = atoi(e); if (v >= 0 && v < p.seg_cap) p.seg_cap = v; }   // tests: force the overflow fallback
    const int prof = memci_prof_begin(ctx, MEMCI_PROF_FS, (double)ctx->fs_last_ops);
    if (sh.nt == 512) rc = bs == 8 ? f8_launch<true, 512>(ctx, tm, p, grid, sh.smem) : f8_launch<false, 512>(ctx, tm, p, grid, sh.smem);
    else rc = bs == 8 ? f8_launch<true, 256>(ctx, tm, p, grid, sh.smem) : f8_launch<false, 256>(ctx, tm, p, grid, sh.smem);
    if (rc) return rc;
    memci_prof_end(ctx, prof);
    CKL(ctx);
    if (getenv("MEMCI_F8_DEBUG")) {
        int hc[5]; unsigned hs[8];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(hc, ctx->d_counters + 8, sizeof(hc), cudaMemcpyDeviceToHost);
        cudaMemcpy(hs, p.seg_counts, sizeof(hs), cudaMemcpyDeviceToHost);
        fprintf(stderr, "f8 debug: err=%s counters[8..12]=%d %d %d %d %d  n_tiles=%d grid=%d seg_cap=%d segs=%u %u %u %u  shape nt=%d ctas=%d rpt=%d stages=%d nb_tile=%d box=%dx%d smem=%d\n",
                cudaGetErrorString(cudaGetLastError()), hc[0], hc[1], hc[2], hc[3], hc[4], n_tiles, grid, p.seg_cap, hs[0], hs[1], hs[2], hs[3],
                sh.nt, sh.ctas, sh.rpt, sh.n_stage, sh.nb_tile, sh.box_w, sh.box_rows, sh.smem);
    }
    const int prof_rf = memci_prof_begin(ctx, MEMCI_PROF_REFINE, 0.0);
    // the flagged-survivor list lives behind the redo marks (F8_FLAG_CAP entries)
    unsigned *d_marks = reinterpret_cast<unsigned *>(ctx->d_fs8_best + (size_t)n_blk * 2);
    unsigned *d_flags = d_marks + (size_t)n_blk * 2;
    fs8_refine_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(fs.luma, ft.luma, H, W, ctx->Wp, bs, R, nbc, n_blk,
                                                                 p.list, p.seg_cap, p.seg_counts, grid, ctx->d_counters, ctx->d_fs8_best,
                                                                 d_flags, F8_FLAG_CAP, d_marks, ctx->d_redo_list, ctx->redo_cap);
    CKL(ctx);
    fs8_replay_kernel<<<32, 256, 0, ctx->stream>>>(fs.luma, ft.luma, fs.rgb, ft.rgb, H, W, ctx->Wp, bs, R, nbc, n_blk, d_flags, F8_FLAG_CAP,
                                                   ctx->d_counters, ctx->d_fs8_best, d_marks, ctx->d_redo_list, ctx->redo_cap);
    CKL(ctx);
    const int n_out = (br1 - br0) * nbc * ndir;
    fs8_output_kernel<<<ceil_div_i(n_out, 256), 256, 0, ctx->stream>>>(ctx->d_fs8_best, reinterpret_cast<unsigned *>(ctx->d_fs8_best + (size_t)n_blk * 2), ctx->d_counters, p.redo_limit,
                                                                      H, W, bs, R, nbc, n_blk, br0 * nbc, br1 * nbc, ndir,
                                                                      out->vec, out->sad, out2 ? out2->vec : out->vec, out2 ? out2->sad : out->sad);
    memci_prof_end(ctx, prof_rf);
    CKL(ctx);
    return MEMCI_OK;
}
