// Motion-compensated interpolation: MCI/Unidirectional.py:36-84 and MCI/Bidirectional.py:27-77.
//
// The reference is a raster-order SCATTER: every source pixel p splats forward (and backward)
// onto q = p + round(mv * dist), and what lands on q depends on the ORDER of arrival (strict /
// non-strict SAD priority, SAD-weighted blend with whatever is already there).  We restate it
// as an order-preserving GATHER, one thread per target pixel q:
//
//   * an "event" at q is a (source pixel p, direction) pair whose splat hits q;
//   * events reach q in raster order of p, forward before backward at the same p;
//   * with e = q - p the effective displacement (wrap-around of negative targets folded in,
//     W1), raster order of p is DESCENDING lexicographic order of e -- independent of q.
//
// So each 32x8 tile of targets collects the set of distinct (e, direction) values of the MV
// cells that can reach it (MVs are piecewise constant on cells, so the set is tiny: ~2-6
// entries for coherent motion), sorts it once, and every pixel walks that sorted list, testing
// "does the pixel at q - e really carry displacement e?" and replaying the reference's state
// machine (SAD_interpolated_frame / Interpolated_Frame) in registers.  Holes (pixels no event
// reached) are recorded and filled by a second kernel with the 21x21 masked median.
#include <type_traits>

#include "memci_internal.cuh"

#define MCI_TW 32
#define MCI_TH 8
#define MCI_NT (MCI_TW * MCI_TH)
#define MCI_SLOTS 1024
#define MCI_MAXKEYS 768
#define MCI_EMPTY 0xffffffffu
#define MCI_KOFF 8191

struct MciParams {
    int H, W, bidir, ndir;
    int y_begin, y_end;          // target rows splatted by this launch (band + hole-fill halo)
    int fill_begin, fill_end;    // target rows whose holes this launch fills
    float dist;
    const uint8_t *src, *tgt;
    uint8_t *out, *mask;
    int *hole_list, *counters;
    const short2 *disp[2];
    int cell[2], cols[2], cshift[2];     // cshift >= 0: cell == 1 << cshift (no integer division)
    const float *sad[2];
    int scell[2], scols[2], sshift[2];
};

// ---- per-cell integer displacement ---------------------------------------------------------
// forward : int(round(mv * dist))        f32 product, round-half-even  (Unidirectional.py:48-49,
//                                         Bidirectional.py:34-35; W1)
// backward: int(round(mv * (1.0 - dist))) f64 product                   (Bidirectional.py:42-44; W2)
__global__ void mci_disp_kernel(const float2 *__restrict__ vec, int n, float dist, int backward,
                                short2 *__restrict__ disp, int *__restrict__ dmax)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int m = 0;
    if (i < n) {
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

__device__ __forceinline__ int cell_of(int x, int cell, int shift) { return shift >= 0 ? (x >> shift) : (x / cell); }

// Apply the event "source pixel (p_r, p_c) of direction DIR lands on this target pixel".
template <int DIR>
__device__ __forceinline__ void mci_apply_t(const MciParams &P, PixState &st, int p_r, int p_c, bool self_write)
{
    const float sad = P.sad[DIR][cell_of(p_r, P.scell[DIR], P.sshift[DIR]) * P.scols[DIR] + cell_of(p_c, P.scell[DIR], P.sshift[DIR])];
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
                // :48-49.  numba types the fused expression float32:
                //   f32( f64( f32(u8*SADI) / t ) + f64(IF) * f64(bsad) / f64(t) ), truncated to int32 (W3)
                const float t = __fadd_rn(sad, st.sadi);
                const double td = (double)t, bd = (double)sad;
                int v[3] = {st.v0, st.v1, st.v2};
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    // fast estimate of the blended value; it decides the truncation unless it falls
                    // within 1e-3 of an integer (its error is < 1e-4), where the exact mixed-precision
                    // expression of the reference is evaluated instead (f64 divide: ~100 instructions)
                    const float est = __fdividef(fmaf((float)px[ch], st.sadi, (float)v[ch] * sad), t);
                    const float fl = floorf(est), fr = est - fl;
                    if (fr > 1e-3f && fr < 0.999f) {
                        v[ch] = (int)fl;
                    } else {
                        float a = __fdiv_rn(__fmul_rn((float)px[ch], st.sadi), t);
                        double c = __ddiv_rn(__dmul_rn((double)v[ch], bd), td);
                        float s = __double2float_rn(__dadd_rn((double)a, c));
                        v[ch] = __float2int_rz(s);
                    }
                }
                st.v0 = v[0]; st.v1 = v[1]; st.v2 = v[2];
            } else {                                    // :51-53
                st.v0 = px[0]; st.v1 = px[1]; st.v2 = px[2];
            }
            st.sadi = sad; st.written = true;
        }
    }
}

__device__ __forceinline__ void mci_apply(const MciParams &P, PixState &st, int p_r, int p_c, int dir, bool self_write)
{
    if (dir == 0) mci_apply_t<0>(P, st, p_r, p_c, self_write);
    else mci_apply_t<1>(P, st, p_r, p_c, self_write);
}

// Generic event test (dense fallback): does the pixel at q - e carry effective displacement e?
__device__ __forceinline__ void mci_try_event(const MciParams &P, PixState &st, int q_r, int q_c,
                                              int eff_r, int eff_c, int dir)
{
    const int p_r = q_r - eff_r, p_c = q_c - eff_c;
    if (p_r < 0 || p_r >= P.H || p_c < 0 || p_c >= P.W) return;
    const short2 d = P.disp[dir][cell_of(p_r, P.cell[dir], P.cshift[dir]) * P.cols[dir] + cell_of(p_c, P.cell[dir], P.cshift[dir])];
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

// sorted list entry: effective displacement, the raw per-cell displacement it stands for, direction.
// kind 0 = ordinary event; kind 1 = the unidirectional self-write slot at e = (0,0).
struct __align__(16) MciEntry { int eff_r, eff_c; unsigned dpacked; int dir_kind; };

#define MCI_STAGE_MAX 1536      // staged displacement cells per direction (6 KB)

__global__ void __launch_bounds__(MCI_NT)
mci_splat_kernel(const MciParams P)
{
    __shared__ unsigned table[MCI_SLOTS];
    __shared__ unsigned list[MCI_MAXKEYS];
    __shared__ unsigned dlist[MCI_MAXKEYS];
    __shared__ MciEntry sorted[MCI_MAXKEYS];
    __shared__ unsigned stage[2][MCI_STAGE_MAX];
    __shared__ int n_keys, overflow;

    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * MCI_TW, ty0 = P.y_begin + blockIdx.y * MCI_TH;
    const int q_c = tx0 + (tid & 31), q_r = ty0 + (tid >> 5);
    const int H = P.H, W = P.W;
    const int dmax = P.counters[2];
    const bool keys_fit = (H + dmax <= MCI_KOFF) && (W + dmax <= MCI_KOFF) && dmax <= MCI_KOFF;

    for (int i = tid; i < MCI_SLOTS; i += MCI_NT) table[i] = MCI_EMPTY;
    if (tid == 0) { n_keys = 0; overflow = keys_fit ? 0 : 1; }
    __syncthreads();

    // un-wrapped neighbourhood of the tile, in cells, per direction (also the staging window)
    int s_r0[2], s_c0[2], s_nc[2];
    bool staged[2];
    for (int dir = 0; dir < 2; ++dir) { s_r0[dir] = s_c0[dir] = 0; s_nc[dir] = 1; staged[dir] = false; }

    if (keys_fit) {
        // ---- 1. distinct (effective displacement, direction) values that can reach this tile ----
        auto insert = [&](unsigned key, unsigned dpacked) {
            unsigned h = (key * 2654435761u) >> 22;          // 10 bits
            for (int probe = 0; probe < MCI_SLOTS; ++probe) {
                const unsigned prev = atomicCAS(&table[h], MCI_EMPTY, key);
                if (prev == key) return;
                if (prev == MCI_EMPTY) {                      // this thread owns the new entry
                    const int pos = atomicAdd(&n_keys, 1);
                    if (pos < MCI_MAXKEYS) { list[pos] = key; dlist[pos] = dpacked; } else overflow = 1;
                    return;
                }
                h = (h + 1) & (MCI_SLOTS - 1);
            }
            overflow = 1;
        };
        for (int dir = 0; dir < P.ndir; ++dir) {
            const int g = P.cell[dir], sh = P.cshift[dir];
            for (int wr = 0; wr < 2; ++wr) {
                int lo_r = ty0 - wr * H - dmax, hi_r = ty0 + MCI_TH - 1 - wr * H + dmax;
                if (wr) hi_r = min(hi_r, dmax - 1);       // wrapping needs u + dr < 0
                lo_r = max(lo_r, 0); hi_r = min(hi_r, H - 1);
                if (lo_r > hi_r) continue;
                for (int wc = 0; wc < 2; ++wc) {
                    int lo_c = tx0 - wc * W - dmax, hi_c = tx0 + MCI_TW - 1 - wc * W + dmax;
                    if (wc) hi_c = min(hi_c, dmax - 1);
                    lo_c = max(lo_c, 0); hi_c = min(hi_c, W - 1);
                    if (lo_c > hi_c) continue;
                    const int cr_lo = cell_of(lo_r, g, sh), cr_hi = cell_of(hi_r, g, sh);
                    const int cc_lo = cell_of(lo_c, g, sh), cc_hi = cell_of(hi_c, g, sh);
                    const int ncc = cc_hi - cc_lo + 1, ncell = (cr_hi - cr_lo + 1) * ncc;
                    const bool do_stage = !wr && !wc && ncell <= MCI_STAGE_MAX;
                    if (do_stage) { s_r0[dir] = cr_lo; s_c0[dir] = cc_lo; s_nc[dir] = ncc; staged[dir] = true; }
                    for (int i = tid; i < ncell; i += MCI_NT) {
                        const int cr = cr_lo + i / ncc, cc = cc_lo + i % ncc;
                        const unsigned dp = reinterpret_cast<const unsigned *>(P.disp[dir])[(size_t)cr * P.cols[dir] + cc];
                        if (do_stage) stage[dir][i] = dp;
                        const int dx_ = (short)(dp & 0xffffu), dy_ = (short)(dp >> 16);   // short2: .x low half, .y high half
                        const int eff_r = dx_ + wr * H, eff_c = dy_ + wc * W;
                        const int pr0 = cr * g, pr1 = min(H, pr0 + g) - 1;
                        const int pc0 = cc * g, pc1 = min(W, pc0 + g) - 1;
                        // conservative footprint test (the per-pixel test below is exact)
                        if (pr1 + eff_r < ty0 || pr0 + eff_r > ty0 + MCI_TH - 1) continue;
                        if (pc1 + eff_c < tx0 || pc0 + eff_c > tx0 + MCI_TW - 1) continue;
                        if (wr ? (pr0 + dx_ >= 0) : (pr1 + dx_ < 0)) continue;
                        if (wc ? (pc0 + dy_ >= 0) : (pc1 + dy_ < 0)) continue;
                        insert(mci_key(eff_r, eff_c, dir), dp);
                    }
                }
            }
        }
        // the unidirectional self-write lives at e = (0,0) and has no footprint of its own
        if (!P.bidir && tid == 0) insert(mci_key(0, 0, 0), 0u);
        __syncthreads();
        // ---- 2. rank-sort (ascending key = raster order of the source pixel) ----
        if (!overflow) {
            const int n = n_keys;
            for (int i = tid; i < n; i += MCI_NT) {
                const unsigned k = list[i];
                int rank = 0;
                for (int j = 0; j < n; ++j) rank += list[j] < k;
                MciEntry e;
                e.eff_r = MCI_KOFF - (int)(k >> 15);
                e.eff_c = MCI_KOFF - (int)((k >> 1) & 16383u);
                e.dpacked = dlist[i];
                e.dir_kind = (int)(k & 1u);
                sorted[rank] = e;
            }
        }
        __syncthreads();
    }

    // ---- 3. replay the splat for this target pixel ----
    const bool inside = q_r < P.y_end && q_c < W;
    PixState st;
    st.v0 = st.v1 = st.v2 = -1;
    st.sadi = __int_as_float(0x7f800000);
    st.written = false;
    if (inside) {
        if (!overflow) {
            const int n = n_keys;
            // An entry stands for "displacement d, wrapped or not": e = d (+H / +W).  If the pixel at
            // p = q - e carries exactly d, its target is q by construction (the reference's bounds test
            // and negative-index wrap are already folded into e), so the test is one compare.
            auto test = [&](auto dir_tag, const MciEntry &e, int p_r, int p_c, int sr0, int sc0, int snc, bool stg) {
                constexpr int D = decltype(dir_tag)::value;
                const int cr = cell_of(p_r, P.cell[D], P.cshift[D]), cc = cell_of(p_c, P.cell[D], P.cshift[D]);
                const int lr = cr - sr0, lc = cc - sc0;
                unsigned dp;
                if (stg && lr >= 0 && lc >= 0 && lc < snc && lr * snc + lc < MCI_STAGE_MAX && e.eff_r <= dmax && e.eff_c <= dmax)
                    dp = stage[D][lr * snc + lc];
                else
                    dp = reinterpret_cast<const unsigned *>(P.disp[D])[(size_t)cr * P.cols[D] + cc];
                if (dp == e.dpacked) {
                    mci_apply_t<D>(P, st, p_r, p_c, false);
                } else if (D == 0 && !P.bidir && e.eff_r == 0 && e.eff_c == 0) {
                    // unidir: this pixel's own vector leaves the frame -> it writes itself (:56-58)
                    const int u_i = p_r + (short)(dp & 0xffffu), v_i = p_c + (short)(dp >> 16);
                    if (!(u_i < H && v_i < W)) mci_apply_t<0>(P, st, p_r, p_c, true);
                }
            };
            for (int i = 0; i < n; ++i) {
                const MciEntry e = sorted[i];                     // one broadcast LDS.128
                const int p_r = q_r - e.eff_r, p_c = q_c - e.eff_c;
                if ((unsigned)p_r >= (unsigned)H || (unsigned)p_c >= (unsigned)W) continue;
                if (e.dir_kind == 0) test(std::integral_constant<int, 0>{}, e, p_r, p_c, s_r0[0], s_c0[0], s_nc[0], staged[0]);
                else test(std::integral_constant<int, 1>{}, e, p_r, p_c, s_r0[1], s_c0[1], s_nc[1], staged[1]);
            }
        } else {
            // dense fallback: enumerate every possible displacement in source-raster order
            // e in [H-dmax, H-1] (wrapped) then [dmax, -dmax] (direct), each value visited once
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
        const size_t q = (size_t)q_r * W + q_c;
        uint8_t *o = P.out + q * 3;
        o[0] = (uint8_t)st.v0; o[1] = (uint8_t)st.v1; o[2] = (uint8_t)st.v2;     // New_IF = uint8(IF)  (:65 / :57)
        P.mask[q] = st.written ? 0 : 1;
    }
    // hole list, one atomic per warp
    const bool hole = inside && !st.written && q_r >= P.fill_begin && q_r < P.fill_end;
    const unsigned bal = __ballot_sync(0xffffffffu, hole);
    if (bal) {
        const int lane = tid & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(&P.counters[1], __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (hole) P.hole_list[base + __popc(bal & ((1u << lane) - 1u))] = q_r * W + q_c;
    }
}

// ---- hole filling: Unidirectional.py:62-80 / Bidirectional.py:55-73 -----------------------------
// Holes are visited in raster order by the reference; each is first overwritten with the target
// frame's pixel, then gets the per-channel median of the non-hole values in its clipped 21x21
// window.  Equivalent rule for hole h (W6): a window pixel contributes its splat value if it is
// not a hole, the TARGET-frame pixel if it is a hole at raster <= h, nothing otherwise.
// np.median: middle element, or the f64 mean of the two middles truncated on the uint8 store.
// One warp per hole; the k-th smallest byte is found by bisection on the value.
// One warp per hole.  The warp histograms the contributing window pixels per channel in shared
// memory (3 x 256 counters), then each lane owns 8 bins: a warp prefix scan locates the lane and
// bin holding rank (n-1)/2 and rank n/2.
#define HF_WARPS 4

__device__ __forceinline__ int hist_select(const unsigned *__restrict__ hist, int lane, int k_lo, int k_hi, int &v_hi)
{
    const uint4 a = *reinterpret_cast<const uint4 *>(hist + lane * 8);
    const uint4 c = *reinterpret_cast<const uint4 *>(hist + lane * 8 + 4);
    const unsigned cnt[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    unsigned local = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) local += cnt[i];
    unsigned incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const unsigned excl = incl - local;
    int found_lo = -1, found_hi = -1;
    if ((unsigned)k_lo >= excl && (unsigned)k_lo < incl) {
        unsigned run = excl;
#pragma unroll
        for (int i = 0; i < 8; ++i) { run += cnt[i]; if (found_lo < 0 && (unsigned)k_lo < run) found_lo = lane * 8 + i; }
    }
    if ((unsigned)k_hi >= excl && (unsigned)k_hi < incl) {
        unsigned run = excl;
#pragma unroll
        for (int i = 0; i < 8; ++i) { run += cnt[i]; if (found_hi < 0 && (unsigned)k_hi < run) found_hi = lane * 8 + i; }
    }
    v_hi = __reduce_max_sync(0xffffffffu, found_hi);
    return __reduce_max_sync(0xffffffffu, found_lo);
}

__global__ void __launch_bounds__(HF_WARPS * 32)
mci_holefill_kernel(const uint8_t *__restrict__ tgt, uint8_t *out, const uint8_t *__restrict__ mask,
                    const int *__restrict__ hole_list, const int *__restrict__ counters, int H, int W)
{
    __shared__ __align__(16) unsigned hist_all[HF_WARPS][3][256];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned (*hist)[256] = hist_all[wib];
    const int n_holes = counters[1];
    for (int hi = warp; hi < n_holes; hi += nwarps) {
        const int h = hole_list[hi];
        const int u = h / W, v = h - u * W;
        const int u0 = max(0, u - 10), u1 = min(H, u + 11), v0 = max(0, v - 10), v1 = min(W, v + 11);
        const int ww = v1 - v0, total = (u1 - u0) * ww;
#pragma unroll
        for (int i = 0; i < 6; ++i)
            *reinterpret_cast<uint4 *>(&hist[0][0] + (i * 32 + lane) * 4) = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        int cnt = 0;
        for (int e = lane; e < total; e += 32) {
            const int a = u0 + e / ww, b = v0 + e % ww;
            const int idx = a * W + b;
            const uint8_t *p = nullptr;
            if (!mask[idx]) p = out + (size_t)idx * 3;            // splat value
            else if (idx <= h) p = tgt + (size_t)idx * 3;         // earlier hole: already replaced by the target pixel
            if (p) {
                atomicAdd(&hist[0][p[0]], 1u);
                atomicAdd(&hist[1][p[1]], 1u);
                atomicAdd(&hist[2][p[2]], 1u);
                ++cnt;
            }
        }
        __syncwarp();
        const int n = __reduce_add_sync(0xffffffffu, cnt);        // >= 1: the hole itself contributes
        const int k_lo = (n - 1) >> 1, k_hi = n >> 1;
        unsigned res[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            int v_hi;
            const int v_lo = hist_select(hist[ch], lane, k_lo, k_hi, v_hi);
            res[ch] = (unsigned)((v_lo + v_hi) >> 1);             // np.median: (a+b)/2.0, truncated on the uint8 store
        }
        __syncwarp();
        if (lane == 0) {
            uint8_t *o = out + (size_t)h * 3;
            o[0] = (uint8_t)res[0]; o[1] = (uint8_t)res[1]; o[2] = (uint8_t)res[2];
        }
    }
}

int memci_mci_run(memci_ctx *ctx, int slot_src, int slot_tgt, const MVField *fwd, const MVField *bwd,
                  float dist, int out_slot, int bidir, int row0, int row1)
{
    const int H = ctx->H, W = ctx->W;
    const MVField *fl[2] = {fwd, bwd};
    const int ndir = bidir ? 2 : 1;
    int rc;
    CK(ctx, cudaMemsetAsync(ctx->d_counters + 1, 0, 2 * sizeof(int), ctx->stream));
    MciParams P;
    memset(&P, 0, sizeof(P));
    for (int d = 0; d < ndir; ++d) {
        const MVField *f = fl[d];
        size_t n = (size_t)f->mv_rows * f->mv_cols;
        if ((rc = memci_reserve(ctx, (void **)&ctx->d_disp[d], &ctx->disp_cap[d], n * sizeof(short2)))) return rc;
        mci_disp_kernel<<<ceil_div_i((int)n, 256), 256, 0, ctx->stream>>>(f->vec, (int)n, dist, d, ctx->d_disp[d],
                                                                          ctx->d_counters + 2);
        CKL(ctx);
        P.disp[d] = ctx->d_disp[d]; P.cell[d] = f->mv_cell; P.cols[d] = f->mv_cols;
        P.sad[d] = f->sad; P.scell[d] = f->sad_cell; P.scols[d] = f->sad_cols;
        P.cshift[d] = (f->mv_cell & (f->mv_cell - 1)) == 0 ? __builtin_ctz(f->mv_cell) : -1;
        P.sshift[d] = (f->sad_cell & (f->sad_cell - 1)) == 0 ? __builtin_ctz(f->sad_cell) : -1;
    }
    size_t npx = (size_t)H * W;
    if (ctx->hole_cap < npx) {
        if (ctx->d_hole_mask) cudaFree(ctx->d_hole_mask);
        if (ctx->d_hole_list) cudaFree(ctx->d_hole_list);
        ctx->d_hole_mask = nullptr; ctx->d_hole_list = nullptr; ctx->hole_cap = 0;
        CK(ctx, cudaMalloc(&ctx->d_hole_mask, npx));
        CK(ctx, cudaMalloc(&ctx->d_hole_list, npx * sizeof(int)));
        ctx->hole_cap = npx;
    }
    if (row1 < 0) { row0 = 0; row1 = H; }
    if (row0 < 0 || row1 > H || row0 >= row1) return memci_fail(ctx, MEMCI_ERR_ARG, "row range [%d, %d) outside [0, %d)", row0, row1, H);
    P.H = H; P.W = W; P.bidir = bidir; P.ndir = ndir; P.dist = dist;
    P.fill_begin = row0; P.fill_end = row1;
    P.y_begin = row0 - 10 > 0 ? row0 - 10 : 0;           // the 21x21 hole-fill window reads 10 rows of splat either side
    P.y_end = row1 + 10 < H ? row1 + 10 : H;
    P.src = ctx->frames[slot_src].rgb; P.tgt = ctx->frames[slot_tgt].rgb;
    P.out = ctx->frames[out_slot].rgb; P.mask = ctx->d_hole_mask;
    P.hole_list = ctx->d_hole_list; P.counters = ctx->d_counters;
    dim3 grid(ceil_div_i(W, MCI_TW), ceil_div_i(P.y_end - P.y_begin, MCI_TH));
    ctx->mci_timed = 0;
    if (ctx->timing) CK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    const int prof = memci_prof_begin(ctx, 1);
    mci_splat_kernel<<<grid, MCI_NT, 0, ctx->stream>>>(P);
    memci_prof_end(ctx, prof);
    CKL(ctx);
    if (ctx->timing) { CK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream)); ctx->mci_timed = 1; }
    mci_holefill_kernel<<<ctx->num_sms * 8, HF_WARPS * 32, 0, ctx->stream>>>(P.tgt, P.out, P.mask, P.hole_list, P.counters, H, W);
    CKL(ctx);
    FrameSlot &o = ctx->frames[out_slot];
    o.valid = 1; o.luma_ok = 0; o.pyr_levels = 0;
    return MEMCI_OK;
}
