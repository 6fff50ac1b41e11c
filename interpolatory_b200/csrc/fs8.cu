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
// depends on the content (textured frames: ~0.3 % of the candidates survive; flat frames: all of them -- those blocks go
// to the block-level redo pass, and past a limit the exact kernel searches the pair, gated on a device-side flag).
//
//   luma8_kernel          L8 as FOUR planes, plane q shifted left by q bytes (zero border): a thread then reads the 8 window
//                         bytes of any column offset as two aligned words of plane (offset & 3), and TMA -- whose inner
//                         coordinate must be 16-byte aligned (probed: any other x raises an illegal instruction) -- can
//                         stage all four shifts with ONE 3-D tile load
//   fs8_prefilter_kernel  pure search: persistent CTAs, dynamic job queue; job = adjacent blocks of one block row (of two,
//                         side by side, where the column range is clipped), staged by TMA (cp.async.bulk.tensor.3d +
//                         mbarrier, two or three stages); thread = (block, dx) candidate column.  Out: an 8-bit monotone code
//                         of every A, eleven to a 16-byte store, column-major (L2-resident until the scan pass reads it),
//                         the exact minimum of every column, and the block minima (one global atomicMin per warp and
//                         block).  No survivor logic, no CTA-wide hand-over: 0.55 of the VABSDIFF4 issue rate at 1080p
//   fs8_scan_kernel       CTA per job: column minima against the block's threshold, the columns that pass (15 - 30 %) are
//                         compacted; thread per listed column compares its codes four at a time and appends its
//                         survivors to one of 64 sub-lists (one global atomic per warp); flat columns -> redo list
//   fs8_refine_kernel     8 lanes per survivor (lane = column of the block): exact luma-SAD, keys combined per CTA, 64-bit
//                         atomicMin; flagged sums go on a short list
//   fs8_replay_kernel     float64 replay of the flagged sums in the reference's operation order
//   fs8_output_kernel     thread per block: key -> (dy, dx, sad); decides the pair-level fallback; re-arms the buffers
// Blocks that cannot be pruned (flat content, flagged sums of 16 x 16 blocks) go to fs_redo_kernel (fs.cu).
#include <stdlib.h>
#include <type_traits>

#include "memci_internal.cuh"

#define F8_G 11
#define F8_NBMAX 32
#define F8_COL_HITS 12           // survivors a candidate column may hold; more: the block has nothing to prune, exact block search (redo pass)
#define F8_NLIST 64             // sub-lists of the survivor list (spreads the scan pass's atomics over as many addresses)
#define F8_CURSOR_STRIDE 32     // words between the cursors of two sub-lists: one 128-byte line each
#define F8_FLAG_CAP 16384        // flagged survivors (sum a multiple of 1000) replayed in float64 per search
#define F8_CODE_CAP ((size_t)512 << 20)   // bytes of candidate codes kept at a time; longer searches run in chunks of jobs

struct F8Tile { short dir, nv; short by, bx0; short nb, pad; };   // direction, block rows (1 or 2), first block row / column, blocks

struct F8Conv { unsigned one, two, c8k, magic, c256, c64k, c16m, unbias; float fmagic; };   // 1, 2, 8192, 0x4B000000, 2^8, 2^16, 2^24, -0x08080800, 8388608.0f: opaque to the compiler

struct F8Params {
    int H, W, bs, m, R, Rp, nbr, nbc, n_blk;
    int t0, n_tiles;             // jobs [t0, t0 + n_tiles) of the table
    const F8Tile *tiles;         // the jobs, in queue order (host table, cached per search geometry)
    int n_stage;                 // shared-memory stages of (window planes + source blocks): the next jobs land while this one is searched
    F8Conv cv;                   // conversion / packing constants, opaque to the compiler
    int box_x0, box_y0;          // padded-plane coordinates of frame pixel (0, 0)
    int rs;                      // words per staged window row (= TMA box width / 4)
    int plane_words;             // words per staged plane, = 8 or 24 (mod 32): the four planes of a row land in different banks
    int ref_rs, ref_rows;        // words per row / rows of the staged source blocks
    int n_groups;                // groups of F8_G row offsets per candidate column = 16-byte words per column in the code array
    unsigned tx_bytes;           // bytes one job's TMA loads deliver
    uint4 *codes;                // [n_tiles][threads][n_groups]: codes of A + 1 of (job, candidate column), eleven row offsets per word
    unsigned *colinfo;           // [n_tiles][threads]: min(A) + 1 of the candidate column (exact) | slot << 16 | column offset << 22; 0xffffffff: no column
    unsigned *bmin;              // [n_tiles][64] A_min + 1 per (job, slot = block row of the job * 32 + block); 0xffffffff between searches
    int *counters;               // [8] pair falls back to the exact kernel, [9] survivors, [10] float64 replays, [11] redo blocks, [12] job queue cursor, [13] flagged survivors
};

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
// of the value, above the threshold).  Nothing of the conversion touches the integer-ALU pipe the SADs are bound by:
// IMAD (2 v + 2^23-as-float bits), FADD (- 2^23: the classic int-to-float without I2F), IMAD.HI (>> 19; leaves code + 2048),
// then the eleven codes of the group are packed into three words by IMADs (x * 256^k + w; the 2048s come off with one
// more IMAD per word) -- all on the FMA pipe, with constants that come from the parameter block (literals would be
// turned back into ALU shifts / logic ops) -- and leave as ONE 16-byte store.  Layout [job][thread][group]: the codes of
// a candidate column are contiguous (48 bytes at 1080p), which is what the scan pass wants -- it reads only the ~15 % of the
// columns whose minimum is near their block's, and a column stored down a [row][thread] array would cost it 33 sectors
// instead of 2.  The column minimum stays exact.  MASK: the group sticks out of the valid row range (last group only).
__device__ __forceinline__ unsigned f8_code(unsigned v, const F8Conv cv)   // v in [1, 65535]; returns code + 2048
{
    unsigned t, x;
    float f;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(v), "r"(cv.two), "r"(cv.magic));
    asm("sub.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(t)), "f"(cv.fmagic));
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(x) : "r"(__float_as_uint(f)), "r"(cv.c8k));
    return x;                                                      // the code is the low byte
}
__device__ __forceinline__ unsigned f8_mad(unsigned a, unsigned b, unsigned c)
{
    unsigned r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
template <bool MASK>
__device__ __forceinline__ void f8_finish_group(unsigned (&acc)[F8_G], int n_valid, uint4 *cg, unsigned &cmin, const F8Conv cv)
{
    static_assert(F8_G == 11, "three words of packed codes per group");
    if (MASK) {
#pragma unroll
        for (int gi = 0; gi < F8_G; ++gi) acc[gi] = gi < n_valid ? acc[gi] : 0xffffu;
    }
    unsigned x[F8_G];
#pragma unroll
    for (int gi = 0; gi < F8_G; ++gi) x[gi] = f8_code(acc[gi], cv);
    uint4 w;
    w.x = f8_mad(f8_mad(x[3], cv.c16m, f8_mad(x[2], cv.c64k, f8_mad(x[1], cv.c256, x[0]))), cv.one, cv.unbias);
    w.y = f8_mad(f8_mad(x[7], cv.c16m, f8_mad(x[6], cv.c64k, f8_mad(x[5], cv.c256, x[4]))), cv.one, cv.unbias);
    w.z = f8_mad(f8_mad(x[10], cv.c64k, f8_mad(x[9], cv.c256, x[8])), cv.one, cv.unbias);
    w.w = 0u;
    *cg = w;
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
// bar_full[stage].  A warp searches its 32 candidate columns, stores the codes, says "done with the stage"
// (bar_empty[stage]) and folds its column minima into the block minima in global memory (segmented warp reduction, one
// atomicMin per warp and block; a block belongs to one job, so the minima are kept per job and slot); warp 0 refills a stage as soon as all sixteen warps are done with it, probing at its
// row-group boundaries.  Nothing here waits for another warp: no __syncthreads() after the prologue, no survivor logic --
// that is fs8_scan_kernel's, which runs when all block minima are final.
template <bool M1, int F8_NT>
__global__ void __launch_bounds__(F8_NT, 1024 / F8_NT)
fs8_prefilter_kernel(const __grid_constant__ CUtensorMap tm_win_a, const __grid_constant__ CUtensorMap tm_win_b,
                     const __grid_constant__ CUtensorMap tm_ref_a, const __grid_constant__ CUtensorMap tm_ref_b,
                     const F8Params p)
{
    constexpr int NW = F8_NT / 32;
    extern __shared__ __align__(128) unsigned f8_smem[];
    __shared__ __align__(8) uint64_t bar_full[3], bar_empty[3];
    __shared__ F8Geo geo_s[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bs = p.bs, R = p.R, n_stage = p.n_stage;
    const F8Conv cv = p.cv;
    const int stage_words = 4 * p.plane_words + ((p.ref_rows * p.ref_rs + 31) & ~31);   // planes [4][plane_words], then source rows [ref_rows][ref_rs]

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // warp 0: take the next job off the queue, publish its geometry (local job number kk, stage sb), arm the stage's
    // barrier and issue the two TMA loads.  Returns false when the queue is empty (an end marker is published instead).
    auto issue_next = [&](int kk, int sb) -> bool {
        int t = 0;
        if (lane == 0) t = atomicAdd(&p.counters[12], 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        F8Geo *g = &geo_s[kk & 3];
        if (t >= p.n_tiles) {
            if (lane == 0) { g->t = -1; mbar_arrive(&bar_full[sb]); }
            return false;
        }
        t += p.t0;
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
        return true;
    };

    // warp 0's bookkeeping: jobs issued so far, and the oldest job whose stage has not been refilled yet
    int n_issued = 0, rf_stage = 0, rf_round = 0;
    bool queue_open = true;
    if (warp == 0) {
        for (; n_issued < n_stage && queue_open; ++n_issued) queue_open = issue_next(n_issued, n_issued);
    }
    // refill every stage whose job all warps are done with; block (sleeping) only while local job `need` is not issued yet
    auto refill = [&](int need) {
        while (queue_open) {
            bool free_ = mbar_test(&bar_empty[rf_stage], (unsigned)(rf_round & 1));
            free_ = __shfl_sync(0xffffffffu, (int)free_, 0) != 0;
            if (!free_) {
                if (n_issued > need) break;
                mbar_wait_sleep(&bar_empty[rf_stage], (unsigned)(rf_round & 1));
            }
            queue_open = issue_next(n_issued, rf_stage);
            ++n_issued;
            if (++rf_stage == n_stage) { rf_stage = 0; ++rf_round; }
        }
    };

    int sb = 0, round = 0;
#pragma unroll 1
    for (int j = 0;; ++j) {
        mbar_wait_sleep(&bar_full[sb], (unsigned)(round & 1));
        const F8Geo *g = &geo_s[j & 3];
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
        uint4 *col = p.codes + ((size_t)(g->t - p.t0) * F8_NT + tid) * p.n_groups;
        unsigned cmin = 0xffffffffu;
        // Lanes without a column of their own (the tail of the last warp) search block 0 / dx 0 along with their warp, so
        // that control flow stays warp-uniform; their results are ignored.
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
                f8_finish_group<decltype(masked)::value>(acc, n_dy - gq * F8_G, col + gq, cmin, cv);
            };
            const int n_full = n_dy / F8_G;
#pragma unroll 1
            for (int gq = 0; gq < n_full; ++gq) {
                group(gq, std::false_type{});
                if (warp == 0) refill(-1);                          // a probe, never a wait
            }
            if (n_full * F8_G < n_dy) group(n_full, std::true_type{});
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[sb]);                 // this warp is done with the stage
        p.colinfo[(size_t)(g->t - p.t0) * F8_NT + tid] = active ? (min(cmin, 65535u) | ((unsigned)(v * 32 + b) << 16) | ((unsigned)dxi << 22)) : 0xffffffffu;
        // block minima: segmented warp reduction, one global atomicMin per (warp, block)
        {
            unsigned *bmin = p.bmin + (size_t)(g->t - p.t0) * 64 + v * 32;
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
        if (warp == 0) refill(j + 1);                               // the next job must be on its way before we wait for it
        if (++sb == n_stage) { sb = 0; ++round; }
    }
}


// CTA per job, a quarter of the prefilter's threads.  Phase 1, four candidate columns per thread (one 16-byte load): a
// column can only hold survivors if its minimum is within T of its block's minimum; the ones that pass (15 - 30 % on
// texture) are compacted, in order, into a shared-memory list.  Phase 2, thread per listed column: its codes --
// contiguous, two or three 16-byte loads -- against the code of the threshold, four at a time; the hits of a warp are
// added up by a shuffle scan, reserved with one global atomic and appended to the survivor list.  Everything phase 1
// reads is indexed by (job, thread) -- the prefilter leaves slot and column offset next to the column minimum, and the
// block minima per job -- so no load waits for another one: the CTA's life is four L2 round trips (column info, codes,
// atomic, list), and small CTAs keep enough of them resident to go through all jobs in under two waves.
// The list is F8_NLIST sub-lists (CTA and warp pick one): same-address atomics retire one every ~4 ns -- 3382 of them,
// one per job, were 14 of the first version's 39 us.
// A column that survives whole is flat content -- nothing to prune in this block: the block goes to the exact
// block-level redo pass instead (once).  A sub-list that would overflow hands the whole pair to the exact kernel.
template <int NT>
__global__ void __launch_bounds__(NT / 4)
fs8_scan_kernel(int H, int W, int bs, int R, int nbc, int n_blk, int n_groups, int T, int t0, const F8Tile *__restrict__ tiles,
                const uint4 *__restrict__ codes, const uint4 *__restrict__ colinfo, unsigned *__restrict__ bmin,
                const F8Conv cv, unsigned *__restrict__ list, unsigned seg_cap, unsigned *__restrict__ cursors,
                unsigned *__restrict__ redo_mark, int *__restrict__ redo_list, int redo_cap, int *__restrict__ counters, int col_hits)
{
    constexpr int NS = NT / 4, NW = NS / 32;
    __shared__ unsigned thr_s[64];                      // min(A_min + 1 + T, 65535) of the slot; 0: no candidates
    __shared__ unsigned flist[NT];                      // listed columns: thread of the prefilter | slot << 9 | column offset << 15
    __shared__ int wcount[NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint4 info4 = colinfo[(size_t)blockIdx.x * NS + tid];
    const F8Tile tile = tiles[t0 + blockIdx.x];
    if (tid < 64) {
        unsigned *bp = bmin + (size_t)blockIdx.x * 64 + tid;
        const unsigned bm = *bp;
        thr_s[tid] = bm == 0xffffffffu ? 0u : min(bm + (unsigned)T, 65535u);
        if (bm != 0xffffffffu) *bp = 0xffffffffu;                  // re-armed for the next search
    }
    __syncthreads();
    const unsigned info[4] = {info4.x, info4.y, info4.z, info4.w};
    unsigned fl = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) fl |= (info[q] != 0xffffffffu && (info[q] & 0xffffu) <= thr_s[(info[q] >> 16) & 63u]) ? 1u << q : 0u;
    const int mine = __popc(fl);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wcount[warp] = incl;
    __syncthreads();
    int base = 0, n_flag = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) { const int c = wcount[w]; base += w < warp ? c : 0; n_flag += c; }
    {
        int pos = base + incl - mine;
#pragma unroll
        for (int q = 0; q < 4; ++q) if (fl & (1u << q)) flist[pos++] = (unsigned)(tid * 4 + q) | ((info[q] >> 16) << 9);
    }
    __syncthreads();
    // ---- phase 2: thread per listed column ----
    const int dir = tile.dir;
    for (int k0 = warp * 32; k0 < n_flag; k0 += NS) {              // warp-uniform trip count (shuffles inside)
        const bool on = k0 + lane < n_flag;
        unsigned long long hm_lo = 0ull, hm_hi = 0ull;
        int hits = 0, blk = 0, dxi = 0;
        bool redo = false;
        if (on) {
            const unsigned e = flist[k0 + lane];
            const int col_t = (int)(e & 511u), slot = (int)((e >> 9) & 63u);
            dxi = (int)(e >> 15);
            const int row = tile.by + (slot >> 5);
            blk = row * nbc + tile.bx0 + (slot & 31);
            const FsBounds8 b0 = f8_bounds(H, W, bs, R, row * bs, tile.bx0 * bs);
            const int n_dy = max(b0.r1 - b0.r0, 0);
            // four codes per step: byte-wise code <= threshold, SWAR.  With xl = x & 0x7f.., T1 = (t & 0x7f..) | 0x80..: bit 7
            // of each byte of T1 - xl says low7(t) >= low7(x) (no borrow crosses a byte), and
            //     x <= t  <=>  (x7 < t7) or (x7 == t7 and low7(x) <= low7(t)).
            // Hits are rare (one or two per column), so only words with a hit are taken apart -- and a column with more than
            // F8_COL_HITS of them belongs to a block with nothing to prune (flat content; a whole job of such columns kept
            // one CTA busy for 45 k cycles, three times the rest of the kernel): the block goes to the redo pass instead.
            const unsigned t4 = (f8_code(thr_s[slot], cv) & 0xffu) * 0x01010101u;
            const unsigned T1 = (t4 & 0x7f7f7f7fu) | 0x80808080u;
            const uint4 *col = codes + ((size_t)blockIdx.x * NT + col_t) * n_groups;
            for (int g0 = 0; g0 * F8_G < n_dy && !redo; g0 += 3) {   // three groups (33 row offsets) at a time: the loads go out together
                uint4 w[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) w[q] = (g0 + q) * F8_G < n_dy ? col[g0 + q] : make_uint4(~0u, ~0u, ~0u, ~0u);
                unsigned le[9];
                int cnt = 0;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const unsigned ww[3] = {w[q].x, w[q].y, w[q].z};
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const unsigned x = ww[k];
                        const unsigned d = T1 - (x & 0x7f7f7f7fu);
                        const int rem = n_dy - ((g0 + q) * F8_G + k * 4);                 // row offsets left in the window from this word on
                        unsigned valid = k == 2 ? 0x00808080u : 0x80808080u;              // a group is eleven codes: the twelfth byte is padding
                        if (rem < 4) valid = rem <= 0 ? 0u : valid & (0x00808080u >> (8 * (3 - rem)));
                        le[q * 3 + k] = ((~x & t4) | (~(x ^ t4) & d)) & valid;
                        cnt += __popc(le[q * 3 + k]);
                    }
                }
                hits += cnt;
                if (hits > col_hits) { redo = true; break; }
#pragma unroll
                for (int q = 0; q < 3; ++q) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        unsigned m = le[q * 3 + k];
                        while (m) {
                            const int bit = __ffs((int)m) - 1;
                            m &= m - 1u;
                            const int idx = (g0 + q) * F8_G + k * 4 + (bit >> 3);
                            if (idx < 64) hm_lo |= 1ull << idx; else hm_hi |= 1ull << (idx - 64);
                        }
                    }
                }
            }
            if (redo) hits = 0;
        }
        int inc2 = hits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc2, o); if (lane >= o) inc2 += t; }
        const int total = __shfl_sync(0xffffffffu, inc2, 31);
        const unsigned seg = (blockIdx.x + (unsigned)(k0 >> 5)) & (F8_NLIST - 1);
        unsigned pos0 = 0u;
        if (total > 0) {                                           // room for the warp's survivors: one global atomic
            if (lane == 0) {
                pos0 = atomicAdd(&cursors[seg * F8_CURSOR_STRIDE], (unsigned)total);
                if (pos0 + (unsigned)total > seg_cap) { counters[8] = 1; pos0 = 0xffffffffu; }     // overflow: the exact kernel takes the pair
            }
            pos0 = __shfl_sync(0xffffffffu, pos0, 0);
        }
        if (hits > 0 && pos0 != 0xffffffffu) {
            unsigned *dst = list + (size_t)seg * seg_cap + pos0 + (unsigned)(inc2 - hits);
            const unsigned head = (dir ? 0x80000000u : 0u) | ((unsigned)blk << 14) | (unsigned)dxi;
            for (unsigned long long mk = hm_lo; mk; mk &= mk - 1) *dst++ = head | ((unsigned)(__ffsll((long long)mk) - 1) << 7);
            for (unsigned long long mk = hm_hi; mk; mk &= mk - 1) *dst++ = head | ((unsigned)(63 + __ffsll((long long)mk)) << 7);
        }
        if (redo) {
            if (atomicExch(&redo_mark[(size_t)dir * n_blk + blk], 0u) != 0u) {
                const int rp = atomicAdd(&counters[11], 1);
                if (rp < redo_cap) redo_list[rp] = blk | (dir ? 0x40000000 : 0);                   // FS_DIR_BIT
            }
        }
    }
}

__device__ __forceinline__ unsigned long long f8_key(unsigned sad, unsigned d2, unsigned ri, unsigned ci)
{
    return ((unsigned long long)sad << 32) | ((unsigned long long)d2 << 16) | (ri << 8) | ci;    // = fs.cu fs_key
}

// Exact SAD of one survivor by eight lanes: lane pr covers COLUMN pr of each 8 x 8 sub-block, row by row -- the eight
// lanes of a group read one 32-byte run per row (one or two sectors of one line), so a warp-wide load touches the lines of
// four survivors' rows, not 32 different ones (lane = row, the first version, kept the L1 tag stage busy five times as
// long: 76 us at 1080p).  Returns the lane's partial sum of |dL1000|.
// (Also measured, both slower than lane = row: a whole warp per survivor with lane = (row, pixel pair), 102 us, and the
// target row as the three aligned 16-byte vectors that cover it, 83 us.)
__device__ __forceinline__ unsigned f8_survivor_sad(const int32_t *__restrict__ Ls, const int32_t *__restrict__ Lt, int H, int W, int Wp,
                                                    int bs, int m, int s_row, int s_col, int t_row, int t_col, int pr)
{
    unsigned S = 0;
    if (s_row + bs <= H && s_col + bs <= W && t_row + bs <= H && t_col + bs <= W) {      // no padding involved
        for (int sub = 0; sub < m * m; ++sub) {
            const int si = sub / m, sj = sub - si * m;
            const int32_t *ps = Ls + (size_t)(s_row + si * 8) * Wp + s_col + sj * 8 + pr;
            const int32_t *pt = Lt + (size_t)(t_row + si * 8) * Wp + t_col + sj * 8 + pr;
            int a[8], b[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) { a[r] = ps[(size_t)r * Wp]; b[r] = pt[(size_t)r * Wp]; }
#pragma unroll
            for (int r = 0; r < 8; ++r) S = __sad(a[r], b[r], S);
        }
    } else {                                                // zero padding (np.pad) on either side
        for (int sub = 0; sub < m * m; ++sub) {
            const int si = sub / m, sj = sub - si * m;
            const int c1 = s_col + sj * 8 + pr, c2 = t_col + sj * 8 + pr;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int r1 = s_row + si * 8 + r, r2 = t_row + si * 8 + r;
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

// Every CTA takes one contiguous range of the survivor list (the concatenation of the scan pass's sub-lists); its 32 eight-lane groups walk that range side by side, so neighbouring groups take neighbouring survivors
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
    // exclusive prefix of the sub-list lengths (n_seg <= F8_NLIST, every CTA computes it for itself)
    __shared__ int seg_pre[F8_NLIST + 1];
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
        for (int k = 0; k < per; ++k) { const int si = threadIdx.x * per + k; if (si < n_seg) loc += (int)seg_counts[si * F8_CURSOR_STRIDE]; }
        int incl = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += v; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += wsum[w];
        int run = wbase + incl - loc;
        for (int k = 0; k < per; ++k) { const int si = threadIdx.x * per + k; if (si < n_seg) { seg_pre[si] = run; run += (int)seg_counts[si * F8_CURSOR_STRIDE]; } }
        if (threadIdx.x == 255) seg_pre[n_seg] = run;
        __syncthreads();
    }
    const int n = seg_pre[n_seg];
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[9] = n;      // statistics
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
    // the list entry of the NEXT turn is loaded while this turn's rows are in flight (one dependent L2 round trip less per turn;
    // no measurable difference at 1080p, where the turn is bound by the row loads themselves)
    auto fetch = [&](int i) -> unsigned {
        if (i >= ce) return 0u;
        while (seg_pre[lo + 1] <= i) ++lo;                         // empty segments are stepped over
        return list[(size_t)lo * seg_cap + (i - seg_pre[lo])];
    };
    unsigned e_next = fetch(cs + grp);
    for (int i0 = cs; i0 < ce; i0 += 32) {                         // warp-uniform trip count (shuffles inside)
        const int i = i0 + grp;
        const bool on = i < ce;
        const unsigned e = e_next;
        e_next = fetch(i + 32);
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

// Also re-arms the key / redo-mark buffers and the sub-list cursors for the next search (no memset per call).
__global__ void fs8_output_kernel(unsigned long long *__restrict__ best, unsigned *__restrict__ redo_mark, unsigned *__restrict__ cursors, int *__restrict__ counters, int redo_limit,
                                  int H, int W, int bs, int R, int nbc, int n_blk, int blk_begin, int blk_end, int ndir,
                                  float2 *__restrict__ out_vec0, float *__restrict__ out_sad0,
                                  float2 *__restrict__ out_vec1, float *__restrict__ out_sad1)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = blk_end - blk_begin;
    if (i < F8_NLIST) cursors[i * F8_CURSOR_STRIDE] = 0u;
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

// Shared memory of one CTA and the TMA boxes for (threads, block rows per job, stages).  Returns false when the
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
    s->smem = n_stage * (4 * s->plane_words + ref_words) * 4;
    return true;
}

// Preference: 512 threads x 2 CTAs per SM, two-row jobs where columns are clipped, as many stages as fit (3, 2, 1); one
// CTA per SM / no two-row jobs / 256 threads as shared memory dictates (wide searches).
// MEMCI_F8_NT / MEMCI_F8_RPT / MEMCI_F8_STAGES / MEMCI_F8_CTAS override (tuning).
static bool f8_shape(int bs, int R, int nbc, F8Shape *out)
{
    static const int cand[][4] = {      // threads, CTAs per SM, block rows per job (max), stages
        {512, 2, 2, 3}, {512, 2, 2, 2}, {512, 1, 2, 3}, {512, 1, 2, 2}, {512, 2, 2, 1}, {512, 2, 1, 2}, {256, 2, 2, 2}, {512, 1, 2, 1},
        {512, 2, 1, 1}, {256, 2, 2, 1}, {256, 2, 1, 2}, {256, 2, 1, 1}, {512, 1, 1, 2}, {512, 1, 1, 1}, {256, 1, 1, 1}};
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

// Is the prefilter path applicable?  Block sizes 8 / 16 (A fits 16 bits), 7-bit window offsets, border of the padded
// planes, a launch shape that fits shared memory.
int memci_fs8_eligible(memci_ctx *ctx, int bs, int R)
{
    if ((bs != 8 && bs != 16) || R < 1 || R > 63 || R + bs + 16 > MEMCI_L8_PAD) return 0;
    if (ceil_div_i(ctx->H, bs) > 32767 || ceil_div_i(ctx->W, bs) > 32767) return 0;                  // 16-bit block coordinates in the job table
    if ((long long)ceil_div_i(ctx->H, bs) * ceil_div_i(ctx->W, bs) >= (1 << 17)) return 0;        // 17-bit block index in a survivor entry
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
static int f8_job_table(memci_ctx *ctx, const F8Shape &sh, int bs, int R, int br0, int br1, int ndir,
                        const F8Tile **tiles, int *n_tiles)
{
    const int H = ctx->H, W = ctx->W, nbc = ceil_div_i(W, bs);
    const int no_split = getenv("MEMCI_F8_NOSPLIT") ? 1 : 0;
    const int key[8] = {bs, R, br0, br1, ndir, sh.nb_tile, sh.nt * 4 + sh.rpt * 2 + no_split, 2};
    if (!(ctx->d_f8_tiles && memcmp(key, ctx->f8_tiles_key, sizeof(key)) == 0)) {
        const int n_cols = ceil_div_i(nbc, sh.nb_tile);
        F8Tile *h = (F8Tile *)malloc(sizeof(F8Tile) * (size_t)ndir * (br1 - br0) * n_cols);
        unsigned char *two = (unsigned char *)calloc(n_cols, 1);
        int *items = (int *)calloc(n_cols, sizeof(int));
        int *n_dx = (int *)calloc(nbc, sizeof(int));
        if (!h || !two || !items || !n_dx) { free(h); free(two); free(items); free(n_dx); return memci_fail(ctx, MEMCI_ERR_NOMEM, "out of host memory"); }
        for (int bc = 0; bc < nbc; ++bc) {                         // column bounds do not depend on the block row
            const int s_col = bc * bs;
            const int tmc = (s_col + bs >= H) ? s_col + 1 : W - bs;        // full_search.py:37 (sic: compares against H)
            const int c0 = s_col - R > 0 ? s_col - R : 0, c1 = tmc < s_col + R + 1 ? tmc : s_col + R + 1;
            n_dx[bc] = c1 > c0 ? c1 - c0 : 0;
            items[bc / sh.nb_tile] += n_dx[bc];
        }
        for (int tc = 0; tc < n_cols; ++tc)
            two[tc] = sh.rpt == 2 && !no_split && items[tc] > 0 && ((items[tc] + 31) & ~31) + items[tc] <= sh.nt;
        int n = 0;
        for (int dir = 0; dir < ndir; ++dir)
            for (int by = br0; by < br1; ++by)
                for (int tc = 0; tc < n_cols; ++tc) {
                    const int bx0 = tc * sh.nb_tile, nb = nbc - bx0 < sh.nb_tile ? nbc - bx0 : sh.nb_tile;
                    if (two[tc] && ((by - br0) & 1)) continue;      // second row of a two-row job
                    F8Tile t;
                    t.dir = (short)dir; t.nv = (short)((two[tc] && by + 1 < br1) ? 2 : 1); t.by = (short)by; t.bx0 = (short)bx0; t.nb = (short)nb; t.pad = 0;
                    h[n++] = t;
                }
        free(two); free(items); free(n_dx);
        if (ctx->d_f8_tiles) { cudaFree(ctx->d_f8_tiles); ctx->d_f8_tiles = nullptr; }
        cudaError_t e = cudaMalloc(&ctx->d_f8_tiles, sizeof(F8Tile) * (size_t)(n > 0 ? n : 1));
        if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_f8_tiles, h, sizeof(F8Tile) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        free(h);
        if (e != cudaSuccess) { ctx->sticky = 1; return memci_fail(ctx, MEMCI_ERR_CUDA, "prefilter job table: %s", cudaGetErrorString(e)); }
        memcpy(ctx->f8_tiles_key, key, sizeof(key));
        ctx->f8_tiles_n = n;
    }
    *tiles = reinterpret_cast<const F8Tile *>(ctx->d_f8_tiles); *n_tiles = ctx->f8_tiles_n;
    return MEMCI_OK;
}


// Enqueue prefilter + scan (per chunk of jobs), refine, replay and output for block rows [br0, br1).  counters[8] is left
// non-zero when the pair has to be searched by the exact kernel; the caller runs that kernel gated on the flag, and
// fs_redo_kernel on the blocks listed here.
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
    p.H = H; p.W = W; p.bs = bs; p.m = bs / 8; p.R = R; p.Rp = (R + 3) & ~3; p.nbr = nbr; p.nbc = nbc; p.n_blk = n_blk;
    p.n_stage = sh.n_stage;
    p.cv.one = 1u; p.cv.two = 2u; p.cv.c8k = 8192u; p.cv.magic = 0x4B000000u; p.cv.c256 = 256u; p.cv.c64k = 65536u; p.cv.c16m = 1u << 24;
    p.cv.unbias = 0u - 0x08080800u; p.cv.fmagic = 8388608.0f;
    int n_jobs = 0;
    if ((rc = f8_job_table(ctx, sh, bs, R, br0, br1, ndir, &p.tiles, &n_jobs))) return rc;
    p.box_x0 = MEMCI_L8_PAD; p.box_y0 = MEMCI_L8_PAD;
    p.rs = sh.rs; p.plane_words = sh.plane_words; p.ref_rs = sh.ref_rs; p.ref_rows = sh.rpt * bs; p.n_groups = sh.a_rows / F8_G;
    p.tx_bytes = (unsigned)(4 * sh.plane_words * 4 + sh.ref_w * sh.rpt * bs);
    const int T = 2 * bs * bs + 1;
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
    // code array: 16 bytes per (job, thread, group of eleven row offsets), at most F8_CODE_CAP bytes at a time -- longer job lists (8K
    // frames) are searched in chunks of whole jobs, scan pass after each; behind it the column minima of the chunk and
    // the block minima of the search
    const size_t job_bytes = (size_t)(sh.a_rows / F8_G) * 16 * sh.nt;
    size_t chunk_jobs = F8_CODE_CAP / job_bytes;
    if (chunk_jobs < 1) chunk_jobs = 1;
    if (chunk_jobs > (size_t)n_jobs) chunk_jobs = (size_t)n_jobs;
    if (const char *e = getenv("MEMCI_F8_CHUNK_JOBS")) { const int v = atoi(e); if (v >= 1 && (size_t)v < chunk_jobs) chunk_jobs = (size_t)v; }   // tests: force chunking
    const size_t off_colmin = (chunk_jobs * job_bytes + 255) & ~(size_t)255;
    const size_t off_bmin = (off_colmin + chunk_jobs * sh.nt * 4 + 255) & ~(size_t)255;
    size_t have_c = ctx->fs8_codes_cap;
    if ((rc = memci_reserve(ctx, (void **)&ctx->d_fs8_codes, &have_c, off_bmin + chunk_jobs * 64 * 4))) return rc;
    const bool new_codes = have_c != ctx->fs8_codes_cap || off_bmin != ctx->fs8_bmin_off;
    ctx->fs8_codes_cap = have_c; ctx->fs8_bmin_off = off_bmin;
    p.codes = reinterpret_cast<uint4 *>(ctx->d_fs8_codes);
    p.colinfo = reinterpret_cast<unsigned *>(ctx->d_fs8_codes + off_colmin);
    p.bmin = reinterpret_cast<unsigned *>(ctx->d_fs8_codes + off_bmin);
    if (new_codes) CK(ctx, cudaMemsetAsync(p.bmin, 0xff, chunk_jobs * 64 * 4, ctx->stream));          // armed once, then by fs8_scan_kernel
    // survivor list: an eighth of the candidates, at most 16 M entries (64 MB), as F8_NLIST sub-lists behind their cursors.
    // Textured content needs about 0.5 %; a sub-list that fills up hands the pair to the exact kernel, so the size is a
    // performance knob, not a correctness one.
    const size_t cand_max = (size_t)n_blk * (2 * R + 1) * (2 * R + 1) * 2;
    size_t cap = cand_max / 8 < (1u << 20) ? (1u << 20) : cand_max / 8;
    if (cap > (16u << 20)) cap = 16u << 20;
    const size_t cursor_words = (size_t)F8_NLIST * F8_CURSOR_STRIDE;
    size_t have = ctx->fs8_list_cap;
    if ((rc = memci_reserve(ctx, (void **)&ctx->d_fs8_list, &have, (cap + cursor_words) * 4))) return rc;
    if (have != ctx->fs8_list_cap) CK(ctx, cudaMemsetAsync(ctx->d_fs8_list, 0, cursor_words * 4, ctx->stream));   // new buffer: cursors zeroed once, then by fs8_output_kernel
    ctx->fs8_list_cap = have;
    unsigned *d_cursors = ctx->d_fs8_list, *d_list = ctx->d_fs8_list + cursor_words;
    unsigned seg_cap = (unsigned)((have / 4 - cursor_words) / F8_NLIST);
    if (const char *e = getenv("MEMCI_F8_SEG_CAP")) { const int v = atoi(e); if (v >= 0 && (unsigned)v < seg_cap) seg_cap = (unsigned)v; }   // tests: force the overflow fallback
    size_t have_b = ctx->fs8_best_cap;
    if ((rc = memci_reserve(ctx, (void **)&ctx->d_fs8_best, &have_b, (size_t)n_blk * 2 * 12 + F8_FLAG_CAP * 4))) return rc;       // keys + redo marks + flagged survivors
    if (have_b != ctx->fs8_best_cap) CK(ctx, cudaMemsetAsync(ctx->d_fs8_best, 0xff, have_b, ctx->stream));    // new buffer: armed once, then by fs8_output_kernel
    ctx->fs8_best_cap = have_b;
    unsigned *d_marks = reinterpret_cast<unsigned *>(ctx->d_fs8_best + (size_t)n_blk * 2);
    unsigned *d_flags = d_marks + (size_t)n_blk * 2;               // the flagged-survivor list lives behind the redo marks
    p.counters = ctx->d_counters;
    int redo_div = 4, col_hits = F8_COL_HITS;       // redo pass: ~4 T pixel-ops/s against the exact kernel's 12.9 -- past a quarter of the blocks the exact kernel is cheaper
    if (const char *e = getenv("MEMCI_F8_REDO_DIV")) { const int v = atoi(e); if (v >= 1) redo_div = v; }      // tuning
    if (const char *e = getenv("MEMCI_F8_COL_HITS")) { const int v = atoi(e); if (v >= 1) col_hits = v; }
    const int redo_limit = n_blk * ndir / redo_div + 64;
    CK(ctx, cudaMemsetAsync(ctx->d_counters + 8, 0, 6 * sizeof(int), ctx->stream));   // [8] overflow [9] survivors [10] float64 replays [11] redo blocks [12] job queue [13] flagged survivors
    const F8Conv cv = p.cv;
    for (int t0 = 0; t0 < n_jobs; t0 += (int)chunk_jobs) {
        p.t0 = t0; p.n_tiles = n_jobs - t0 < (int)chunk_jobs ? n_jobs - t0 : (int)chunk_jobs;
        if (t0 > 0) CK(ctx, cudaMemsetAsync(ctx->d_counters + 12, 0, sizeof(int), ctx->stream));
        int grid = p.n_tiles < sh.ctas * ctx->num_sms ? p.n_tiles : sh.ctas * ctx->num_sms;
        const int prof = memci_prof_begin(ctx, MEMCI_PROF_FS, (double)ctx->fs_last_ops * p.n_tiles / n_jobs);
        if (sh.nt == 512) rc = bs == 8 ? f8_launch<true, 512>(ctx, tm, p, grid, sh.smem) : f8_launch<false, 512>(ctx, tm, p, grid, sh.smem);
        else rc = bs == 8 ? f8_launch<true, 256>(ctx, tm, p, grid, sh.smem) : f8_launch<false, 256>(ctx, tm, p, grid, sh.smem);
        if (rc) return rc;
        memci_prof_end(ctx, prof);
        CKL(ctx);
        const int prof_sc = memci_prof_begin(ctx, getenv("MEMCI_F8_SCANKIND") ? MEMCI_PROF_HBM_SMALL : MEMCI_PROF_REFINE, 0.0);
        if (sh.nt == 512)
            fs8_scan_kernel<512><<<p.n_tiles, 128, 0, ctx->stream>>>(H, W, bs, R, nbc, n_blk, p.n_groups, T, t0, p.tiles, p.codes, reinterpret_cast<const uint4 *>(p.colinfo), p.bmin, cv,
                                                                     d_list, seg_cap, d_cursors, d_marks, ctx->d_redo_list, ctx->redo_cap, ctx->d_counters, col_hits);
        else
            fs8_scan_kernel<256><<<p.n_tiles, 64, 0, ctx->stream>>>(H, W, bs, R, nbc, n_blk, p.n_groups, T, t0, p.tiles, p.codes, reinterpret_cast<const uint4 *>(p.colinfo), p.bmin, cv,
                                                                    d_list, seg_cap, d_cursors, d_marks, ctx->d_redo_list, ctx->redo_cap, ctx->d_counters, col_hits);
        memci_prof_end(ctx, prof_sc);
        CKL(ctx);
    }
    if (getenv("MEMCI_F8_DEBUG")) {
        int hc[6];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(hc, ctx->d_counters + 8, sizeof(hc), cudaMemcpyDeviceToHost);
        fprintf(stderr, "f8 debug: err=%s counters[8..13]=%d %d %d %d %d %d  n_jobs=%d chunk=%d seg_cap=%u  shape nt=%d ctas=%d rpt=%d stages=%d nb_tile=%d box=%dx%d smem=%d\n",
                cudaGetErrorString(cudaGetLastError()), hc[0], hc[1], hc[2], hc[3], hc[4], hc[5], n_jobs, (int)chunk_jobs, seg_cap,
                sh.nt, sh.ctas, sh.rpt, sh.n_stage, sh.nb_tile, sh.box_w, sh.box_rows, sh.smem);
    }
    const int prof_rf = memci_prof_begin(ctx, MEMCI_PROF_REFINE, 0.0);
    fs8_refine_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(fs.luma, ft.luma, H, W, ctx->Wp, bs, R, nbc, n_blk,
                                                                 d_list, (int)seg_cap, d_cursors, F8_NLIST, ctx->d_counters, ctx->d_fs8_best,
                                                                 d_flags, F8_FLAG_CAP, d_marks, ctx->d_redo_list, ctx->redo_cap);
    CKL(ctx);
    fs8_replay_kernel<<<32, 256, 0, ctx->stream>>>(fs.luma, ft.luma, fs.rgb, ft.rgb, H, W, ctx->Wp, bs, R, nbc, n_blk, d_flags, F8_FLAG_CAP,
                                                   ctx->d_counters, ctx->d_fs8_best, d_marks, ctx->d_redo_list, ctx->redo_cap);
    CKL(ctx);
    const int n_out = (br1 - br0) * nbc * ndir;
    fs8_output_kernel<<<ceil_div_i(n_out, 256), 256, 0, ctx->stream>>>(ctx->d_fs8_best, d_marks, d_cursors, ctx->d_counters, redo_limit,
                                                                      H, W, bs, R, nbc, n_blk, br0 * nbc, br1 * nbc, ndir,
                                                                      out->vec, out->sad, out2 ? out2->vec : out->vec, out2 ? out2->sad : out->sad);
    memci_prof_end(ctx, prof_rf);
    CKL(ctx);
    return MEMCI_OK;
}
