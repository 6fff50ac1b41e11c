// Motion-compensated interpolation: MCI/Unidirectional.py:36-84 and MCI/Bidirectional.py:27-77.
//
// The reference is a raster-order SCATTER: every source pixel p splats forward (and backward)
// onto q = p + round(mv * dist), and what lands on q depends on the ORDER of arrival (strict /
// non-strict SAD priority, SAD-weighted blend with whatever is already there).  We restate it
// as an order-preserving GATHER over the target pixels q:
//
//   * an "event" at q is a (source pixel p, direction) pair whose splat hits q;
//   * events reach q in raster order of p, forward before backward at the same p;
//   * with e = q - p the effective displacement (wrap-around of negative targets folded in,
//     W1), raster order of p is DESCENDING lexicographic order of e -- independent of q.
//
// So each 32x8 tile of targets (one warp) collects the set of distinct (e, direction) values of the
// MV cells that can reach it (MVs are piecewise constant on cells, so the set is tiny: ~2-6
// entries for coherent motion), sorts it once, and every pixel walks that sorted list, testing
// "does the pixel at q - e really carry displacement e?" and replaying the reference's state
// machine (SAD_interpolated_frame / Interpolated_Frame) as a decision record in registers; the
// pixel values are fetched afterwards.  Holes (pixels no event reached) are recorded as runs and
// filled by a second kernel with the 21x21 masked median (sliding histogram along the run).
#include <stdlib.h>
#include <type_traits>

#include "memci_internal.cuh"

#define MCI_EMPTY 0xffffffffu
#define MCI_KOFF 8191
#define HF_WARPS 8
#define HF_MAXRUN 8                          // holes per run (power of two, <= 32)
#define HF_TW (21 + HF_MAXRUN - 1)           // staged window width of a run
#define HF_TWP (HF_TW | 1)                   // odd row stride: conflict-free column reads

struct MciParams {
    int H, W, bidir, ndir;
    int y_begin, y_end;          // target rows splatted by this launch (band + hole-fill halo)
    int fill_begin, fill_end;    // target rows whose holes this launch fills
    float dist;
    const uint8_t *src, *tgt;
    uint8_t *out;
    uint32_t *fill;              // per pixel: v0 | v1<<8 | v2<<16 | hole<<24 (holes carry the TARGET-frame pixel): hole-fill input
    int *hole_list, *counters;
    const short2 *disp[2];
    int cell[2], cols[2];
    unsigned cmul[2];            // ceil(2^32 / cell): x / cell == umulhi(x, cmul) for 0 <= x < 2^16 (0 when cell == 1)
    const float *sad[2];
    int scell[2], scols[2];
    unsigned smul[2];
    int sad_nested[2];           // scell % cell == 0: every pixel of an MV cell reads the same SAD cell
    int force_path;              // debug (MEMCI_MCI_PATH): 0 auto, 1 never the staged fast path, 2 always the dense enumeration
};

// ---- per-cell integer displacement ---------------------------------------------------------
// forward : int(round(mv * dist))        f32 product, round-half-even  (Unidirectional.py:48-49,
//                                         Bidirectional.py:34-35; W1)
// backward: int(round(mv * (1.0 - dist))) f64 product                   (Bidirectional.py:42-44; W2)
struct DispArgs { const float2 *vec[2]; short2 *disp[2]; int n[2]; };
__global__ void mci_disp_kernel(const DispArgs A, float dist, int ndir, int *__restrict__ dmax)
{
    // both directions in one launch: blockIdx.y = direction (0 forward, 1 backward)
    const int backward = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = backward ? A.n[1] : A.n[0];
    const float2 *vec = backward ? A.vec[1] : A.vec[0];
    short2 *disp = backward ? A.disp[1] : A.disp[0];
    int m = 0;
    if (backward < ndir && i < n) {
        float2 v = vec[i];
        int dr, dc;
        if (!backward) {
            dr = (int)rintf(__fmul_rn(v.x, dist));
            dc = (int)rintf(__fmul_rn(v.y, dist));
        } else {
            double bw = __dsub_rn(1.0, (double)dist);
            dr = (int)rint(__dmul_rn((double)v.x, bw));
            dc = (int)rint(__dmul_rn((double)v.y, bw));
        }
        dr = max(-32767, min(32767, dr));
        dc = max(-32767, min(32767, dc));
        disp[i] = make_short2((short)dr, (short)dc);
        m = max(abs(dr), abs(dc));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(dmax, m);
}

// ---- per-pixel state machine -----------------------------------------------------------------
struct PixState {
    int v0, v1, v2;      // Interpolated_Frame[q] (int32, -1 = hole)
    float sadi;          // SAD_interpolated_frame[q]
    bool written;
};

// x / cell for 0 <= x < 2^16 as one IMAD.HI (exact: x * cell < 2^32); mul == 0 encodes cell == 1
__device__ __forceinline__ int cell_of(int x, unsigned mul) { return mul ? (int)__umulhi((unsigned)x, mul) : x; }

// Exact mixed-precision blend of Bidirectional.py:48-49 (W3); out of line: it is the rare path of
// mci_apply_t and the splat kernel instantiates that 16 times.
__device__ __noinline__ int mci_blend_exact(float pxv, float sadi, float t, int v, double bd, double td)
{
    float a = __fdiv_rn(__fmul_rn(pxv, sadi), t);
    double c = __ddiv_rn(__dmul_rn((double)v, bd), td);
    float s = __double2float_rn(__dadd_rn((double)a, c));
    return __float2int_rz(s);
}

// Apply the event "source pixel (p_r, p_c) of direction DIR lands on this target pixel".
template <int DIR>
__device__ __forceinline__ float mci_sad_at(const MciParams &P, int p_r, int p_c)
{
    return P.sad[DIR][cell_of(p_r, P.smul[DIR]) * P.scols[DIR] + cell_of(p_c, P.smul[DIR])];
}

// Bidirectional.py:48-49: IF = tgt*SADI/t + IF*bsad/t with t = bsad + SADI.  numba types the fused
// expression float32:  f32( f64( f32(u8*SADI) / t ) + f64(IF) * f64(bsad) / f64(t) ), truncated to int32 (W3).
// A fast estimate decides the truncation unless it falls within 1e-3 of an integer (its error is < 1e-4),
// where the exact mixed-precision expression is evaluated instead (f64 divide: ~100 instructions).
__device__ __forceinline__ void mci_blend3(int &v0, int &v1, int &v2, int t0, int t1, int t2, float sad, float sadi)
{
    const float t = __fadd_rn(sad, sadi);
    const double td = (double)t, bd = (double)sad;
    int v[3] = {v0, v1, v2};
    const int px[3] = {t0, t1, t2};
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float est = __fdividef(fmaf((float)px[ch], sadi, (float)v[ch] * sad), t);
        const float fl = floorf(est), fr = est - fl;
        if (fr > 1e-3f && fr < 0.999f) v[ch] = (int)fl;
        else v[ch] = mci_blend_exact((float)px[ch], sadi, t, v[ch], bd, td);
    }
    v0 = v[0]; v1 = v[1]; v2 = v[2];
}

template <int DIR>
__device__ __forceinline__ void mci_apply_t(const MciParams &P, PixState &st, int p_r, int p_c, bool self_write, const float sad)
{
    const uint8_t *px = (DIR == 0 ? P.src : P.tgt) + ((size_t)p_r * P.W + p_c) * 3;
    if (!P.bidir) {
        if (self_write || sad <= st.sadi) {             // Unidirectional.py:52 (non-strict) / :56-58
            st.v0 = px[0]; st.v1 = px[1]; st.v2 = px[2];
            st.sadi = sad; st.written = true;
        }
    } else if (DIR == 0) {
        if (sad < st.sadi) {                            // Bidirectional.py:37 (strict)
            st.v0 = px[0]; st.v1 = px[1]; st.v2 = px[2];
            st.sadi = sad; st.written = true;
        }
    } else {
        if (sad < st.sadi) {                            // :46
            if (st.sadi != __int_as_float(0x7f800000)) {
                mci_blend3(st.v0, st.v1, st.v2, px[0], px[1], px[2], sad, st.sadi);       // :48-49
            } else {                                    // :51-53
                st.v0 = px[0]; st.v1 = px[1]; st.v2 = px[2];
            }
            st.sadi = sad; st.written = true;
        }
    }
}

__device__ __forceinline__ void mci_apply(const MciParams &P, PixState &st, int p_r, int p_c, int dir, bool self_write)
{
    if (dir == 0) mci_apply_t<0>(P, st, p_r, p_c, self_write, mci_sad_at<0>(P, p_r, p_c));
    else mci_apply_t<1>(P, st, p_r, p_c, self_write, mci_sad_at<1>(P, p_r, p_c));
}

// Generic event test (dense fallback): does the pixel at q - e carry effective displacement e?
__device__ __forceinline__ void mci_try_event(const MciParams &P, PixState &st, int q_r, int q_c,
                                              int eff_r, int eff_c, int dir)
{
    const int p_r = q_r - eff_r, p_c = q_c - eff_c;
    if (p_r < 0 || p_r >= P.H || p_c < 0 || p_c >= P.W) return;
    const short2 d = P.disp[dir][cell_of(p_r, P.cmul[dir]) * P.cols[dir] + cell_of(p_c, P.cmul[dir])];
    const int u_i = p_r + d.x, v_i = p_c + d.y;
    bool self_write = false;
    if (u_i < P.H && v_i < P.W) {                       // Unidirectional.py:51, Bidirectional.py:36/45
        const int t_r = u_i < 0 ? u_i + P.H : u_i;      // negative index wraps (W1)
        const int t_c = v_i < 0 ? v_i + P.W : v_i;
        if (t_r != q_r || t_c != q_c) return;
    } else {
        // unidir only: out-of-frame target writes the pixel onto itself, unconditionally (:56-58)
        if (P.bidir || eff_r != 0 || eff_c != 0) return;
        self_write = true;
    }
    mci_apply(P, st, p_r, p_c, dir, self_write);
}

__device__ __forceinline__ unsigned mci_key(int eff_r, int eff_c, int dir)
{
    return ((unsigned)(MCI_KOFF - eff_r) << 15) | ((unsigned)(MCI_KOFF - eff_c) << 1) | (unsigned)dir;
}

// ---- splat kernel: one WARP per 32 x SW_TH tile of target pixels -------------------------------
// Lane = column, the SW_TH rows of the lane live in registers.  The warp (no CTA barrier anywhere)
//   1. scans the MV cells whose footprint can reach the tile (un-wrapped and, at the frame's far edges,
//      wrapped neighbourhoods), de-duplicates their (effective displacement, direction) keys with
//      match.any inside a 32-cell chunk and a broadcast scan of the short list across chunks, and stages
//      the un-wrapped neighbourhood's packed displacements in shared memory;
//   2. rank-sorts the list (ascending key = raster order of the source pixel, forward before backward);
//   3. walks it: for every entry the lane tests its SW_TH pixels -- "does the cell of q - e carry exactly
//      this displacement?" is one shared-memory load and one compare -- and replays the reference's state
//      machine on a match.
// Output: the uint8 frame, the packed hole-fill image (P.fill) and the list of hole runs.
#ifndef SW_TH
#define SW_TH 8
#endif
#ifndef SW_MINB
#define SW_MINB 4
#endif
#define SW_WARPS 4
#define SW_MAXK 256          // distinct (displacement, direction) keys per tile; more -> dense fallback
#define SW_STAGE 320         // staged displacement cells per direction

struct SwShared {
    unsigned keys[SW_MAXK];
    unsigned dps[SW_MAXK];
    uint2 sorted[SW_MAXK];   // .x = key, .y = packed per-cell displacement the entry stands for
    unsigned stage[2][SW_STAGE];
    float stage_sad[2][SW_STAGE];
    unsigned rowbuf[24];     // one output row of the tile (32 px x 3 B) for 4-byte coalesced stores
    int s_r0[2], s_c0[2], s_nc[2], staged[2];   // staged neighbourhood per direction (first cell row / column, cells per row)
};

// Dense fallback (key list overflow / huge frames): enumerate every possible displacement in
// source-raster order, e in [H-dmax, H-1] (wrapped) then [dmax, -dmax] (direct), each value once.
__device__ __noinline__ void sw_dense_pixel(const MciParams &P, PixState &st, int q_r, int q_c, int dmax)
{
    const int H = P.H, W = P.W;
    for (int eff_r = q_r; eff_r >= -dmax; --eff_r) {
        if (eff_r > dmax && eff_r < H - dmax) { eff_r = dmax + 1; continue; }
        if (q_r - eff_r >= H) break;
        for (int eff_c = q_c; eff_c >= -dmax; --eff_c) {
            if (eff_c > dmax && eff_c < W - dmax) { eff_c = dmax + 1; continue; }
            if (q_c - eff_c >= W) break;
            for (int dir = 0; dir < P.ndir; ++dir) mci_try_event(P, st, q_r, q_c, eff_r, eff_c, dir);
        }
    }
}

// ---- decision state of one target pixel ----------------------------------------------------------
// The control flow of the reference's state machine depends only on SADs, never on pixel values, so
// the walk (phase A) records WHICH events fire -- the last overriding copy plus at most one backward
// blend stacked on it -- in five registers per pixel, without touching the frames; phase B then issues
// the pixel loads of the whole tile at once and evaluates the blends.
//   sadi : SAD_interpolated_frame[q]
//   base : kind << 28 | payload.  kind 0 nothing yet (hole), 1 source-frame pixel (payload = linear index),
//          2 target-frame pixel, 3 literal packed value (a longer blend chain that was evaluated early)
//   bp   : target-frame pixel of the pending backward blend (-1: none), bs / bi: its bsad and the SADI it saw
#define SW_KIND(base) ((unsigned)(base) >> 28)
#define SW_PAYLOAD(base) ((base) & 0x0fffffff)

// 24-bit RGB of pixel p_lin as two aligned word loads (frame slots carry 16 bytes of slack)
__device__ __forceinline__ void sw_px_issue(const uint8_t *frame, int p_lin, unsigned &lo, unsigned &hi, unsigned &sh)
{
    const size_t a = (size_t)frame + (size_t)p_lin * 3;
    const unsigned *w = reinterpret_cast<const unsigned *>(a & ~(size_t)3);
    lo = w[0]; hi = w[1]; sh = (unsigned)(a & 3) * 8u;
}
__device__ __forceinline__ unsigned sw_px_finish(unsigned lo, unsigned hi, unsigned sh) { return __funnelshift_r(lo, hi, sh) & 0xffffffu; }

// Bidirectional.py:48-49 on packed values: v = blend(target pixel t, v)
__device__ __noinline__ unsigned sw_blend_packed(unsigned v, unsigned t, float sad, float sadi)
{
    int v0 = (int)(v & 255u), v1 = (int)((v >> 8) & 255u), v2 = (int)((v >> 16) & 255u);
    mci_blend3(v0, v1, v2, (int)(t & 255u), (int)((t >> 8) & 255u), (int)((t >> 16) & 255u), sad, sadi);
    return (unsigned)v0 | ((unsigned)v1 << 8) | ((unsigned)v2 << 16);
}

// value of a pixel whose decision state is (base, bp, bs, bi); only called when base != hole
__device__ __noinline__ unsigned sw_eval(const MciParams *P, int base, int bp, float bs, float bi)
{
    unsigned v;
    if (SW_KIND(base) == 3u) v = (unsigned)SW_PAYLOAD(base) & 0xffffffu;
    else {
        unsigned lo, hi, sh;
        sw_px_issue(SW_KIND(base) == 1u ? P->src : P->tgt, SW_PAYLOAD(base), lo, hi, sh);
        v = sw_px_finish(lo, hi, sh);
    }
    if (bp >= 0) {
        unsigned lo, hi, sh;
        sw_px_issue(P->tgt, bp, lo, hi, sh);
        v = sw_blend_packed(v, sw_px_finish(lo, hi, sh), bs, bi);
    }
    return v;
}

// The event "pixel p_lin of direction D lands on this target pixel" (Unidirectional.py:52-58,
// Bidirectional.py:37-53) as a decision only, written with selects (the call sites are unrolled).
template <int D, bool BIDIR>
__device__ __forceinline__ void sw_fire(const MciParams &P, float &sadi, int &base, int &bp, float &bs, float &bi,
                                        int p_lin, float sad, bool self_write)
{
    if (!BIDIR) {
        const bool t = self_write || sad <= sadi;                       // :52 (non-strict) / :56-58
        base = t ? ((1 << 28) | p_lin) : base;
        sadi = t ? sad : sadi;
    } else if (D == 0) {
        const bool t = sad < sadi;                                      // :37 (strict): plain copy
        base = t ? ((1 << 28) | p_lin) : base;
        bp = t ? -1 : bp;
        sadi = t ? sad : sadi;
    } else {
        const bool t = sad < sadi;                                      // :46
        const bool occ = sadi != __int_as_float(0x7f800000);
        const bool blend = t && occ;                                    // :48-49 blend with what is there
        if (blend && bp >= 0) base = (int)((3u << 28) | sw_eval(&P, base, bp, bs, bi));   // second stacked blend (rare): evaluate the first now
        base = (t && !occ) ? ((2 << 28) | p_lin) : base;                // :51-53 empty: copy
        bp = t ? (occ ? p_lin : -1) : bp;
        bs = blend ? sad : bs;
        bi = blend ? sadi : bi;
        sadi = t ? sad : sadi;
    }
}

// Fast path: one list entry against the lane's SW_TH target pixels, everything staged in shared memory
// (un-wrapped entry, displacement + SAD of the cell staged).  p_r = ty0 + j - eff_r is the same for every
// lane, so all row arithmetic is warp-uniform; per lane and row it is one shared-memory load and a compare.
template <int D, bool BIDIR>
__device__ __forceinline__ void sw_walk_entry(const MciParams &P, const SwShared &sm, float (&sadi)[SW_TH], int (&base)[SW_TH],
                                              int (&bp)[SW_TH], float (&bs)[SW_TH], float (&bi)[SW_TH], const uint2 ent,
                                              int eff_r, int eff_c, int ty0, int q_c, bool col_in, int sr0, int sc0, int snc)
{
    const int H = P.H, W = P.W;
    const int p_c = q_c - eff_c;
    const bool col_ok = col_in && (unsigned)p_c < (unsigned)W;
    const int lc = (col_ok ? cell_of(p_c, P.cmul[D]) : sc0) - sc0;
    const int p_r0 = ty0 - eff_r;
    const unsigned want = col_ok ? ent.y : 0x80008000u;          // (-32768, -32768) is never a displacement: mci_disp_kernel clamps to +-32767
    if (P.cell[D] >= SW_TH) {                                     // warp-uniform
        // MV cells at least as tall as the tile: the tile's source rows lie in at most two cell rows, so the lane tests two
        // staged cells instead of SW_TH, and an entry no lane matches is dropped after one vote
        const int j_lo = max(0, -p_r0), j_hi = min(SW_TH, min(H - p_r0, P.y_end - ty0));      // rows [j_lo, j_hi) have a source row
        if (j_lo >= j_hi) return;
        const int cA = cell_of(p_r0 + j_lo, P.cmul[D]);
        const int j_split = (cA + 1) * P.cell[D] - p_r0;          // first tile row whose source row is in the next cell row
        const int siA = (cA - sr0) * snc + lc;
        const bool hitA = sm.stage[D][siA] == want;
        const float sadA = sm.stage_sad[D][siA];
        bool hitB = false;
        float sadB = 0.f;
        if (j_split < j_hi) { hitB = sm.stage[D][siA + snc] == want; sadB = sm.stage_sad[D][siA + snc]; }
        if (!__any_sync(0xffffffffu, hitA || hitB)) return;
#pragma unroll
        for (int j = 0; j < SW_TH; ++j) {
            if (j < j_lo || j >= j_hi) continue;                  // warp-uniform
            const bool in_b = j >= j_split;
            if (in_b ? hitB : hitA)
                sw_fire<D, BIDIR>(P, sadi[j], base[j], bp[j], bs[j], bi[j], (p_r0 + j) * W + p_c, in_b ? sadB : sadA, false);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < SW_TH; ++j) {
        const int p_r = p_r0 + j;
        if ((unsigned)p_r >= (unsigned)H || ty0 + j >= P.y_end) continue;          // warp-uniform
        const int si = (cell_of(p_r, P.cmul[D]) - sr0) * snc + lc;
        if (sm.stage[D][si] == want)
            sw_fire<D, BIDIR>(P, sadi[j], base[j], bp[j], bs[j], bi[j], p_r * W + p_c, sm.stage_sad[D][si], false);
    }
}

// Generic path: one list entry against one target pixel, any geometry (wrapped entries, neighbourhoods too
// large to stage, SAD cells that do not nest, the unidirectional self-write slot).
template <int D, bool BIDIR>
__device__ __forceinline__ void sw_test_entry(const MciParams &P, const SwShared &sm, float &sadi, int &base, int &bp, float &bs, float &bi,
                                              const uint2 ent, int eff_r, int eff_c, int p_r, int p_c, bool use_stage, int sr0, int sc0, int snc)
{
    const int cr = cell_of(p_r, P.cmul[D]), cc = cell_of(p_c, P.cmul[D]);
    const unsigned dp = use_stage ? sm.stage[D][(cr - sr0) * snc + (cc - sc0)]
                                  : reinterpret_cast<const unsigned *>(P.disp[D])[(size_t)cr * P.cols[D] + cc];
    if (dp == ent.y) {
        sw_fire<D, BIDIR>(P, sadi, base, bp, bs, bi, p_r * P.W + p_c, mci_sad_at<D>(P, p_r, p_c), false);
    } else if (D == 0 && !BIDIR && eff_r == 0 && eff_c == 0) {
        // unidir: this pixel's own vector leaves the frame -> it writes itself (:56-58)
        const int u_i = p_r + (short)(dp & 0xffffu), v_i = p_c + (short)(dp >> 16);
        if (!(u_i < P.H && v_i < P.W)) sw_fire<0, false>(P, sadi, base, bp, bs, bi, p_r * P.W + p_c, mci_sad_at<0>(P, p_r, p_c), true);
    }
}

template <bool BIDIR>
__global__ void __launch_bounds__(SW_WARPS * 32, SW_MINB)
mci_splat_kernel(const __grid_constant__ MciParams P, int tiles_x, int n_tiles)
{
    __shared__ SwShared sm_all[SW_WARPS];
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * SW_WARPS + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
#ifdef SW_PHASE_TIMING
    long long t_ph0 = clock64();
#endif
    SwShared &sm = sm_all[threadIdx.x >> 5];
    const int tyi = tile / tiles_x, txi = tile - tyi * tiles_x;
    const int tx0 = txi * 32, ty0 = P.y_begin + tyi * SW_TH;
    const int q_c = tx0 + lane;
    const int H = P.H, W = P.W;
    const bool col_in = q_c < W;
    const int dmax = P.counters[2];
    // A wrapped effective displacement lies in [H - dmax, H - 1], an un-wrapped one in [-dmax, dmax]: on a frame
    // no larger than 2 * dmax the two ranges meet and one key would stand for two different cell displacements.
    const bool keys_fit = (H + dmax <= MCI_KOFF) && (W + dmax <= MCI_KOFF) && dmax <= MCI_KOFF && H > 2 * dmax && W > 2 * dmax;
    const unsigned lt_mask = (1u << lane) - 1u;

    int n = 0;
    bool overflow = !keys_fit || P.force_path == 2;
    bool any_far = false;                                      // a wrapped entry made it into the list
    if (lane < 2) { sm.s_r0[lane] = 0; sm.s_c0[lane] = 0; sm.s_nc[lane] = 1; sm.staged[lane] = 0; }
    __syncwarp();

    if (keys_fit) {
        // ---- 1. distinct (effective displacement, direction) values that can reach this tile ----
#pragma unroll 1
        for (int dir = 0; dir < P.ndir; ++dir) {
            const int g = P.cell[dir];
            const unsigned gm = P.cmul[dir];
            const unsigned *gdisp = reinterpret_cast<const unsigned *>(P.disp[dir]);
#pragma unroll 1
            for (int wr = 0; wr < 2; ++wr) {
                int lo_r = ty0 - wr * H - dmax, hi_r = ty0 + SW_TH - 1 - wr * H + dmax;
                if (wr) hi_r = min(hi_r, dmax - 1);       // wrapping needs u + dr < 0
                lo_r = max(lo_r, 0); hi_r = min(hi_r, H - 1);
                if (lo_r > hi_r) continue;
#pragma unroll 1
                for (int wc = 0; wc < 2; ++wc) {
                    int lo_c = tx0 - wc * W - dmax, hi_c = tx0 + 31 - wc * W + dmax;
                    if (wc) hi_c = min(hi_c, dmax - 1);
                    lo_c = max(lo_c, 0); hi_c = min(hi_c, W - 1);
                    if (lo_c > hi_c) continue;
                    const int cr_lo = cell_of(lo_r, gm), cr_hi = cell_of(hi_r, gm);
                    const int cc_lo = cell_of(lo_c, gm), cc_hi = cell_of(hi_c, gm);
                    const int ncc = cc_hi - cc_lo + 1, ncell = (cr_hi - cr_lo + 1) * ncc;
                    const bool do_stage = !wr && !wc && ncell <= SW_STAGE;
                    if (do_stage && lane == 0) { sm.s_r0[dir] = cr_lo; sm.s_c0[dir] = cc_lo; sm.s_nc[dir] = ncc; sm.staged[dir] = 1; }
                    for (int base = 0; base < ncell; base += 32) {
                        const int i = base + lane;
                        bool want = false;
                        unsigned key = MCI_EMPTY, dp = 0u;
                        if (i < ncell) {
                            const int ro = (int)__fdividef((float)i + 0.5f, (float)ncc);    // exact for these small integers
                            const int cr = cr_lo + ro, cc = cc_lo + (i - ro * ncc);
                            dp = gdisp[(size_t)cr * P.cols[dir] + cc];
                            if (do_stage) {
                                sm.stage[dir][i] = dp;
                                if (P.sad_nested[dir])
                                    sm.stage_sad[dir][i] = P.sad[dir][cell_of(cr * g, P.smul[dir]) * P.scols[dir] + cell_of(cc * g, P.smul[dir])];
                            }
                            const int dx_ = (short)(dp & 0xffffu), dy_ = (short)(dp >> 16);   // short2: .x low half, .y high half
                            const int eff_r = dx_ + wr * H, eff_c = dy_ + wc * W;
                            const int pr0 = cr * g, pr1 = min(H, pr0 + g) - 1;
                            const int pc0 = cc * g, pc1 = min(W, pc0 + g) - 1;
                            // conservative footprint test (the per-pixel test in the walk is exact)
                            want = !(pr1 + eff_r < ty0 || pr0 + eff_r > ty0 + SW_TH - 1) &&
                                   !(pc1 + eff_c < tx0 || pc0 + eff_c > tx0 + 31) &&
                                   !(wr ? (pr0 + dx_ >= 0) : (pr1 + dx_ < 0)) &&
                                   !(wc ? (pc0 + dy_ >= 0) : (pc1 + dy_ < 0));
                            if (want) key = mci_key(eff_r, eff_c, dir);
                        }
                        if (!__any_sync(0xffffffffu, want)) continue;
                        const unsigned peers = __match_any_sync(0xffffffffu, key);
                        bool isnew = want && (__ffs(peers) - 1 == lane);
                        const int n_scan = min(n, SW_MAXK);
                        for (int j = 0; j < n_scan; ++j) isnew = isnew && (sm.keys[j] != key);     // broadcast reads
                        const unsigned nb = __ballot_sync(0xffffffffu, isnew);
                        if (isnew) {
                            const int pos = n + __popc(nb & lt_mask);
                            if (pos < SW_MAXK) { sm.keys[pos] = key; sm.dps[pos] = dp; }
                        }
                        n += __popc(nb);
                        if (nb && (wr || wc)) any_far = true;
                        __syncwarp();
                    }
                }
            }
        }
        // the unidirectional self-write lives at e = (0,0) and has no footprint of its own
        if (!BIDIR) {
            const unsigned key = mci_key(0, 0, 0);
            bool have = false;
            const int n_scan = min(n, SW_MAXK);
            for (int j = 0; j < n_scan; ++j) have = have || (sm.keys[j] == key);
            if (!have) {
                if (lane == 0 && n < SW_MAXK) { sm.keys[n] = key; sm.dps[n] = 0u; }
                ++n;
            }
            __syncwarp();
        }
        if (n > SW_MAXK) overflow = true;
        // ---- 2. rank-sort (ascending key = raster order of the source pixel) ----
        if (!overflow) {
            for (int i = lane; i < n; i += 32) {
                const unsigned k = sm.keys[i];
                int rank = 0;
                for (int j = 0; j < n; ++j) rank += sm.keys[j] < k;
                sm.sorted[rank] = make_uint2(k, sm.dps[i]);
            }
        }
        __syncwarp();
    }

#ifdef SW_PHASE_TIMING
    long long t_ph1 = clock64();
#endif
    // ---- 3. phase A: which events fire for the lane's SW_TH pixels (no frame access) ----
    unsigned val[SW_TH];                                       // packed result, bit 24 = hole
    const int sr0_0 = sm.s_r0[0], sc0_0 = sm.s_c0[0], snc_0 = sm.s_nc[0], sr0_1 = sm.s_r0[1], sc0_1 = sm.s_c0[1], snc_1 = sm.s_nc[1];
    const bool st0 = sm.staged[0] != 0, st1 = sm.staged[1] != 0;
    // fast path: every entry un-wrapped and its cells (displacement + SAD) staged; unidirectional tiles must
    // also be out of reach of the self-write rule (a vector leaving the frame at the bottom / right, :56-58)
    const bool fast = P.force_path == 0 && !overflow && !any_far && st0 && P.sad_nested[0] &&
                      (BIDIR ? (st1 && P.sad_nested[1]) : (ty0 + SW_TH - 1 + dmax < H && tx0 + 31 + dmax < W));
    if (fast) {
        float sadi[SW_TH], bs[SW_TH], bi[SW_TH];
        int base[SW_TH], bp[SW_TH];
#pragma unroll
        for (int j = 0; j < SW_TH; ++j) { sadi[j] = __int_as_float(0x7f800000); base[j] = 0; bp[j] = -1; bs[j] = 0.f; bi[j] = 0.f; }
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const uint2 ent = sm.sorted[i];                    // one broadcast LDS.64
            const int eff_r = MCI_KOFF - (int)(ent.x >> 15);
            const int eff_c = MCI_KOFF - (int)((ent.x >> 1) & 16383u);
            if (!BIDIR || (ent.x & 1u) == 0u) sw_walk_entry<0, BIDIR>(P, sm, sadi, base, bp, bs, bi, ent, eff_r, eff_c, ty0, q_c, col_in, sr0_0, sc0_0, snc_0);
            else sw_walk_entry<1, BIDIR>(P, sm, sadi, base, bp, bs, bi, ent, eff_r, eff_c, ty0, q_c, col_in, sr0_1, sc0_1, snc_1);
        }
        // ---- phase B: the pixel loads of four rows in flight together, then the pending blends ----
#pragma unroll
        for (int h = 0; h < SW_TH; h += 4) {
            unsigned alo[4], ahi[4], ash[4], blo[4], bhi[4], bsh[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = h + k;
                alo[k] = ahi[k] = ash[k] = blo[k] = bhi[k] = bsh[k] = 0u;
                const unsigned kind = SW_KIND(base[j]);
                if (kind == 1u || kind == 2u) sw_px_issue(kind == 1u ? P.src : P.tgt, SW_PAYLOAD(base[j]), alo[k], ahi[k], ash[k]);
                if (BIDIR && bp[j] >= 0) sw_px_issue(P.tgt, bp[j], blo[k], bhi[k], bsh[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = h + k;
                const unsigned kind = SW_KIND(base[j]);
                unsigned v = kind == 3u ? ((unsigned)SW_PAYLOAD(base[j]) & 0xffffffu) : sw_px_finish(alo[k], ahi[k], ash[k]);
                if (BIDIR && bp[j] >= 0) v = sw_blend_packed(v, sw_px_finish(blo[k], bhi[k], bsh[k]), bs[j], bi[j]);
                val[j] = kind == 0u ? 0x1ffffffu : v;         // hole: uint8(-1) per channel until the fill kernel runs
            }
        }
    } else {
        // generic path, one tile row at a time (rolled loops)
#pragma unroll 1
        for (int j = 0; j < SW_TH; ++j) {
            const int q_r = ty0 + j;
            unsigned v = 0x1ffffffu;
            if (q_r < P.y_end && col_in) {
                if (!overflow) {
                    float sadi = __int_as_float(0x7f800000), bs = 0.f, bi = 0.f;
                    int base = 0, bp = -1;
#pragma unroll 1
                    for (int i = 0; i < n; ++i) {
                        const uint2 ent = sm.sorted[i];
                        const int eff_r = MCI_KOFF - (int)(ent.x >> 15);
                        const int eff_c = MCI_KOFF - (int)((ent.x >> 1) & 16383u);
                        const int p_r = q_r - eff_r, p_c = q_c - eff_c;
                        if ((unsigned)p_r >= (unsigned)H || (unsigned)p_c >= (unsigned)W) continue;
                        // un-wrapped entries (|e| <= dmax) only ever look at cells of the staged neighbourhood
                        const bool near = eff_r <= dmax && eff_c <= dmax;
                        if (!BIDIR || (ent.x & 1u) == 0u) sw_test_entry<0, BIDIR>(P, sm, sadi, base, bp, bs, bi, ent, eff_r, eff_c, p_r, p_c, st0 && near, sr0_0, sc0_0, snc_0);
                        else sw_test_entry<1, BIDIR>(P, sm, sadi, base, bp, bs, bi, ent, eff_r, eff_c, p_r, p_c, st1 && near, sr0_1, sc0_1, snc_1);
                    }
                    if (SW_KIND(base) != 0u) v = sw_eval(&P, base, bp, bs, bi);
                } else {
                    PixState st;
                    st.v0 = st.v1 = st.v2 = -1;
                    st.sadi = __int_as_float(0x7f800000);
                    st.written = false;
                    sw_dense_pixel(P, st, q_r, q_c, dmax);
                    if (st.written) v = (unsigned)(st.v0 & 255) | ((unsigned)(st.v1 & 255) << 8) | ((unsigned)(st.v2 & 255) << 16);
                }
            }
#pragma unroll
            for (int jj = 0; jj < SW_TH; ++jj) if (jj == j) val[jj] = v;
        }
    }

#ifdef SW_PHASE_TIMING
    long long t_ph2 = clock64();
#endif
    // ---- 4. outputs: uint8 frame, packed hole-fill image, hole runs (one atomic pair per tile) ----
    // The per-row values go through shared memory (keys[] is dead after the sort) so that this loop can
    // stay rolled: the unrolled form was a quarter of the kernel's code.
    static_assert(SW_MAXK >= SW_TH * 32, "value buffer aliases keys[]");
    unsigned *vbuf = sm.keys;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < SW_TH; ++j) vbuf[j * 32 + lane] = val[j];
    __syncwarp();
    const bool words_ok = (W & 3) == 0;                       // row starts 4-byte aligned: coalesced word stores
    int n_holes = 0, n_starts = 0;
    // pass 1: stores + per-row hole / run-start ballots (kept in shared memory: dps[] is dead too)
    unsigned *balbuf = sm.dps;                                 // [0..7] hole ballots, [8..15] run-start ballots
#pragma unroll 1
    for (int j = 0; j < SW_TH; ++j) {
        const int q_r = ty0 + j;
        if (q_r >= P.y_end) { if (lane == 0) { balbuf[j] = 0u; balbuf[8 + j] = 0u; } continue; }     // warp-uniform
        const size_t q = (size_t)q_r * W + q_c;
        const unsigned vj = vbuf[j * 32 + lane];
        const unsigned rgb = vj & 0xffffffu;
        const bool is_hole = (vj >> 24) != 0u;
        if (words_ok) {
            unsigned char *rb = reinterpret_cast<unsigned char *>(sm.rowbuf);
            rb[lane * 3] = (unsigned char)rgb; rb[lane * 3 + 1] = (unsigned char)(rgb >> 8); rb[lane * 3 + 2] = (unsigned char)(rgb >> 16);
            __syncwarp();
            const int n_words = (min(32, W - tx0) * 3) >> 2;   // W % 4 == 0 and tx0 % 32 == 0: whole words
            if (lane < n_words) reinterpret_cast<unsigned *>(P.out + ((size_t)q_r * W + tx0) * 3)[lane] = sm.rowbuf[lane];
            __syncwarp();
        } else if (col_in) {
            uint8_t *o = P.out + q * 3;
            o[0] = (uint8_t)rgb; o[1] = (uint8_t)(rgb >> 8); o[2] = (uint8_t)(rgb >> 16);     // New_IF = uint8(IF)  (:65 / :57)
        }
        if (col_in) {
            unsigned word = rgb;
            if (is_hole) {                                     // :70 / :63: a hole is first replaced by the target pixel
                const uint8_t *t = P.tgt + q * 3;
                word = (unsigned)t[0] | ((unsigned)t[1] << 8) | ((unsigned)t[2] << 16) | (1u << 24);
            }
            P.fill[q] = word;
        }
        // hole runs: consecutive hole lanes become one (first pixel, length) entry, capped at HF_MAXRUN
        const bool hole = col_in && is_hole && q_r >= P.fill_begin && q_r < P.fill_end;
        const unsigned bal = __ballot_sync(0xffffffffu, hole);
        const bool start = hole && ((lane & (HF_MAXRUN - 1)) == 0 || !((bal >> (lane - 1)) & 1u));
        const unsigned sb = __ballot_sync(0xffffffffu, start);
        if (lane == 0) { balbuf[j] = bal; balbuf[8 + j] = sb; }
        n_holes += __popc(bal); n_starts += __popc(sb);
    }
    if (n_starts) {                                            // warp-uniform
        __syncwarp();
        int pos0 = 0;
        if (lane == 0) {
            pos0 = atomicAdd(&P.counters[4], n_starts);
            atomicAdd(&P.counters[1], n_holes);
        }
        pos0 = __shfl_sync(0xffffffffu, pos0, 0);
#pragma unroll 1
        for (int j = 0; j < SW_TH; ++j) {
            const unsigned bal = balbuf[j], sb = balbuf[8 + j];
            if ((sb >> lane) & 1u) {
                const unsigned inv = ~(bal >> lane);
                const int len = min(inv ? __ffs(inv) - 1 : 32, HF_MAXRUN - (lane & (HF_MAXRUN - 1)));
                const int pos = pos0 + __popc(sb & lt_mask);
                P.hole_list[2 * pos] = (ty0 + j) * W + q_c;
                P.hole_list[2 * pos + 1] = len;
            }
            pos0 += __popc(sb);
        }
    }
#ifdef SW_PHASE_TIMING
    if (lane == 0) { long long t_ph3 = clock64(); unsigned long long *tp = reinterpret_cast<unsigned long long *>(P.hole_list) + 100000; atomicAdd(tp + 0, (unsigned long long)(t_ph1 - t_ph0)); atomicAdd(tp + 1, (unsigned long long)(t_ph2 - t_ph1)); atomicAdd(tp + 2, (unsigned long long)(t_ph3 - t_ph2)); atomicAdd(tp + 3, (unsigned long long)n); atomicAdd(tp + 4, fast ? 1ull : 0ull); atomicAdd(tp + 5, 1ull); }
#endif
}

// ---- hole filling: Unidirectional.py:62-80 / Bidirectional.py:55-73 -----------------------------
// Holes are visited in raster order by the reference; each is first overwritten with the target
// frame's pixel, then gets the per-channel median of the non-hole values in its clipped 21x21
// window.  Equivalent rule for hole h (W6): a window pixel contributes its splat value if it is
// not a hole, the TARGET-frame pixel if it is a hole at raster <= h, nothing otherwise.
// np.median: middle element, or the f64 mean of the two middles truncated on the uint8 store.
//
// One warp per run of <= HF_MAXRUN horizontally adjacent holes.  The union of the run's windows
// (<= 21 x 28 packed words of P.fill) is staged in shared memory with independent coalesced loads;
// the first hole histograms its whole window, every next hole (one pixel to the right, same rows)
// slides it: the column entering on the right is added, the column leaving on the left is removed,
// and the hole itself switches from "later hole, excluded" to "replaced by the target pixel" --
// ~43 shared-memory updates instead of 441 global reads.  The three channel histograms share one
// word per bin (10 bits each, counts <= 441), so one warp prefix scan ranks all three channels.
__device__ __forceinline__ void hist_select3(const unsigned *__restrict__ hist, int lane, int k_lo, int k_hi, unsigned (&res)[3])
{
    // Each lane owns 8 bins; the three channel counts of a bin share one word (10-bit fields, counts <= 441,
    // so bit 9 of every field is free).  "prefix >= k+1" for all three channels at once is one subtraction
    // with that bit as a guard: (prefix | guard) - (k+1) keeps the guard bit exactly when there is no borrow.
    const unsigned GUARD = 0x20080200u, ONES = 0x00100401u;
    const uint4 a = *reinterpret_cast<const uint4 *>(hist + lane * 8);
    const uint4 c = *reinterpret_cast<const uint4 *>(hist + lane * 8 + 4);
    const unsigned cnt[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    unsigned local = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) local += cnt[i];          // packed add: no field exceeds 441
    unsigned incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const unsigned excl = incl - local;
    const unsigned k1_lo = (unsigned)(k_lo + 1) * ONES, k1_hi = (unsigned)(k_hi + 1) * ONES;
    unsigned run = excl + GUARD, ge_lo = 0u, ge_hi = 0u;  // per field: how many of the lane's 8 prefixes reach rank k
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        run += cnt[i];
        ge_lo += ((run - k1_lo) >> 9) & ONES;
        ge_hi += ((run - k1_hi) >> 9) & ONES;
    }
    // the lane holding rank k: exclusive prefix <= k < inclusive prefix; its bin is 8 - (#prefixes >= k+1)
    const unsigned in_lo = ((incl + GUARD - k1_lo) >> 9) & ~((excl + GUARD - k1_lo) >> 9) & ONES;
    const unsigned in_hi = ((incl + GUARD - k1_hi) >> 9) & ~((excl + GUARD - k1_hi) >> 9) & ONES;
    const unsigned bin1 = (unsigned)(lane * 8 + 9);       // bin + 1 so that 0 means "not in this lane"
    const unsigned v_lo = __reduce_or_sync(0xffffffffu, in_lo * bin1 - (ge_lo & (in_lo * 1023u)));
    const unsigned v_hi = __reduce_or_sync(0xffffffffu, in_hi * bin1 - (ge_hi & (in_hi * 1023u)));
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)                         // np.median: (a+b)/2.0, truncated on the uint8 store
        res[ch] = (((v_lo >> (10 * ch)) & 1023u) + ((v_hi >> (10 * ch)) & 1023u) - 2u) >> 1;
}

__device__ __forceinline__ void hist_add(unsigned *hist, unsigned word)
{
    atomicAdd(&hist[word & 255u], 1u);
    atomicAdd(&hist[(word >> 8) & 255u], 1u << 10);
    atomicAdd(&hist[(word >> 16) & 255u], 1u << 20);
}
__device__ __forceinline__ void hist_sub(unsigned *hist, unsigned word)
{
    atomicSub(&hist[word & 255u], 1u);
    atomicSub(&hist[(word >> 8) & 255u], 1u << 10);
    atomicSub(&hist[(word >> 16) & 255u], 1u << 20);
}

__global__ void __launch_bounds__(HF_WARPS * 32)
mci_holefill_kernel(const uint32_t *__restrict__ fill, uint8_t *__restrict__ out,
                    const int *__restrict__ hole_list, const int *__restrict__ counters, int H, int W)
{
    __shared__ __align__(16) unsigned hist_all[HF_WARPS][256];
    __shared__ unsigned win_all[HF_WARPS][21 * HF_TWP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned *hist = hist_all[wib], *win = win_all[wib];
    const int n_runs = counters[4];
    for (int ri = warp; ri < n_runs; ri += nwarps) {
        const int h0 = hole_list[2 * ri], len = hole_list[2 * ri + 1];
        const int u = h0 / W, vs = h0 - u * W;
        const int u0 = max(0, u - 10), u1 = min(H, u + 11), nrows = u1 - u0;
        const int c0 = max(0, vs - 10), c1 = min(W, vs + len + 10), tw = c1 - c0;       // union of the run's windows
        *reinterpret_cast<uint4 *>(hist + lane * 8) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(hist + lane * 8 + 4) = make_uint4(0u, 0u, 0u, 0u);
        {
            const int total = nrows * tw;                                          // <= 588
            const unsigned rcp = (65536u + (unsigned)tw - 1u) / (unsigned)tw;         // e / tw as a multiply (exact for e < 588, tw <= 28)
            for (int e = lane; e < total; e += 32) {
                const int ra = (int)(((unsigned)e * rcp) >> 16), cb = e - ra * tw;
                win[ra * HF_TWP + cb] = fill[(size_t)(u0 + ra) * W + c0 + cb];
            }
        }
        __syncwarp();
        // ---- full window of the first hole ----
        int cnt = 0;
        {
            const int ww = min(W, vs + 11) - c0, total = nrows * ww;
            const unsigned rcp = (65536u + (unsigned)ww - 1u) / (unsigned)ww;
            for (int e = lane; e < total; e += 32) {
                const int ra = (int)(((unsigned)e * rcp) >> 16), cb = e - ra * ww;
                const unsigned word = win[ra * HF_TWP + cb];
                if (!(word >> 24) || (u0 + ra) * W + c0 + cb <= h0) { hist_add(hist, word); ++cnt; }
            }
        }
        int n = __reduce_add_sync(0xffffffffu, cnt);              // >= 1: the hole itself contributes
        for (int s = 0; s < len; ++s) {
            const int v = vs + s, h = h0 + s;
            __syncwarp();
            unsigned res[3];
            hist_select3(hist, lane, (n - 1) >> 1, n >> 1, res);
            if (lane == 0) {
                uint8_t *o = out + (size_t)h * 3;
                o[0] = (uint8_t)res[0]; o[1] = (uint8_t)res[1]; o[2] = (uint8_t)res[2];
            }
            if (s + 1 == len) break;
            __syncwarp();
            int delta = 0;
            if (lane < nrows) {
                const int a = u0 + lane;
                if (v + 11 < W) {                                  // column entering on the right
                    const unsigned word = win[lane * HF_TWP + (v + 11 - c0)];
                    if (!(word >> 24) || a < u) { hist_add(hist, word); ++delta; }
                }
                if (v - 10 >= 0) {                                 // column leaving on the left
                    const unsigned word = win[lane * HF_TWP + (v - 10 - c0)];
                    if (!(word >> 24) || a <= u) { hist_sub(hist, word); --delta; }
                }
            } else if (lane == 31) {                               // the next hole itself: now a target pixel
                hist_add(hist, win[(u - u0) * HF_TWP + (v + 1 - c0)]);
                ++delta;
            }
            n += __reduce_add_sync(0xffffffffu, delta);
        }
        __syncwarp();
    }
}

int memci_mci_run(memci_ctx *ctx, int slot_src, int slot_tgt, const MVField *fwd, const MVField *bwd,
                  float dist, int out_slot, int bidir, int row0, int row1)
{
    const int H = ctx->H, W = ctx->W;
    const MVField *fl[2] = {fwd, bwd};
    const int ndir = bidir ? 2 : 1;
    int rc;
    CK(ctx, cudaMemsetAsync(ctx->d_counters + 1, 0, 4 * sizeof(int), ctx->stream));   // [1] holes [2] dmax [4] hole runs
    MciParams P;
    memset(&P, 0, sizeof(P));
    DispArgs DA;
    memset(&DA, 0, sizeof(DA));
    size_t n_max = 0;
    for (int d = 0; d < ndir; ++d) {
        const MVField *f = fl[d];
        size_t n = (size_t)f->mv_rows * f->mv_cols;
        if ((rc = memci_reserve(ctx, (void **)&ctx->d_disp[d], &ctx->disp_cap[d], n * sizeof(short2)))) return rc;
        DA.vec[d] = f->vec; DA.disp[d] = ctx->d_disp[d]; DA.n[d] = (int)n;
        n_max = n > n_max ? n : n_max;
        P.disp[d] = ctx->d_disp[d]; P.cell[d] = f->mv_cell; P.cols[d] = f->mv_cols;
        P.sad[d] = f->sad; P.scell[d] = f->sad_cell; P.scols[d] = f->sad_cols;
        P.sad_nested[d] = (f->sad_cell % f->mv_cell) == 0;
        P.cmul[d] = f->mv_cell > 1 ? (unsigned)((0x100000000ull + (unsigned)f->mv_cell - 1) / (unsigned)f->mv_cell) : 0u;
        P.smul[d] = f->sad_cell > 1 ? (unsigned)((0x100000000ull + (unsigned)f->sad_cell - 1) / (unsigned)f->sad_cell) : 0u;
    }
    mci_disp_kernel<<<dim3(ceil_div_i((int)n_max, 256), ndir), 256, 0, ctx->stream>>>(DA, dist, ndir, ctx->d_counters + 2);
    CKL(ctx);
    size_t npx = (size_t)H * W;
    if (ctx->hole_cap < npx) {
        if (ctx->d_fill) cudaFree(ctx->d_fill);
        if (ctx->d_hole_list) cudaFree(ctx->d_hole_list);
        ctx->d_fill = nullptr; ctx->d_hole_list = nullptr; ctx->hole_cap = 0;
        CK(ctx, cudaMalloc(&ctx->d_fill, npx * sizeof(uint32_t)));
        CK(ctx, cudaMalloc(&ctx->d_hole_list, npx * sizeof(int)));
        ctx->hole_cap = npx;
    }
    if (row1 < 0) { row0 = 0; row1 = H; }
    if (row0 < 0 || row1 > H || row0 >= row1) return memci_fail(ctx, MEMCI_ERR_ARG, "row range [%d, %d) outside [0, %d)", row0, row1, H);
    P.H = H; P.W = W; P.bidir = bidir; P.ndir = ndir; P.dist = dist;
    if (const char *e = getenv("MEMCI_MCI_PATH")) P.force_path = atoi(e);
    P.fill_begin = row0; P.fill_end = row1;
    P.y_begin = row0 - 10 > 0 ? row0 - 10 : 0;           // the 21x21 hole-fill window reads 10 rows of splat either side
    P.y_end = row1 + 10 < H ? row1 + 10 : H;
    P.src = ctx->frames[slot_src].rgb; P.tgt = ctx->frames[slot_tgt].rgb;
    P.out = ctx->frames[out_slot].rgb; P.fill = ctx->d_fill;
    P.hole_list = ctx->d_hole_list; P.counters = ctx->d_counters;
    const int tiles_x = ceil_div_i(W, 32), n_tiles = tiles_x * ceil_div_i(P.y_end - P.y_begin, SW_TH);
    ctx->mci_timed = 0;
    if (ctx->timing) CK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
#ifdef SW_PHASE_TIMING
    cudaMemsetAsync(reinterpret_cast<unsigned long long *>(ctx->d_hole_list) + 100000, 0, 64, ctx->stream);
#endif
    const int prof = memci_prof_begin(ctx, MEMCI_PROF_MCI, 9.0 * (double)(P.y_end - P.y_begin) * W);      // 2 frame reads + 1 write
    if (bidir) mci_splat_kernel<true><<<ceil_div_i(n_tiles, SW_WARPS), SW_WARPS * 32, 0, ctx->stream>>>(P, tiles_x, n_tiles);
    else mci_splat_kernel<false><<<ceil_div_i(n_tiles, SW_WARPS), SW_WARPS * 32, 0, ctx->stream>>>(P, tiles_x, n_tiles);
    memci_prof_end(ctx, prof);
    CKL(ctx);
    if (ctx->timing) { CK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream)); ctx->mci_timed = 1; }
#ifdef SW_PHASE_TIMING
    { unsigned long long h[6]; cudaStreamSynchronize(ctx->stream); cudaMemcpy(h, reinterpret_cast<unsigned long long *>(ctx->d_hole_list) + 100000, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "splat phases (avg cycles per tile-warp): list %.0f, walk+pixels %.0f, output %.0f; entries/tile %.2f, fast tiles %.3f, tiles %llu\n", (double)h[0] / h[5], (double)h[1] / h[5], (double)h[2] / h[5], (double)h[3] / h[5], (double)h[4] / h[5], h[5]); }
#endif
    const int prof_hf = memci_prof_begin(ctx, MEMCI_PROF_HOLEFILL, 0.0);
    mci_holefill_kernel<<<ctx->num_sms * 6, HF_WARPS * 32, 0, ctx->stream>>>(P.fill, P.out, P.hole_list, P.counters, H, W);
    memci_prof_end(ctx, prof_hf);
    CKL(ctx);
    FrameSlot &o = ctx->frames[out_slot];
    o.valid = 1; o.luma_ok = 0; o.luma8_ok = 0; o.pyr_levels = 0;
    return MEMCI_OK;
}
