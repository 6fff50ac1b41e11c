// Hierarchical block matching (ME/hbma.py) on sm_100a.
//
// HBMA evaluates ~27 candidates per block per level (3 candidate vectors x 3x3 refinement),
// two orders of magnitude less SAD work than full search, so the kernels here are the simple
// warp-per-block form: lanes = candidates, block + search region staged per warp in shared memory.
//
// Cost model (get_cost, hbma.py:15-31): float32( sum_f64 |L1-L2| ), compared as float32.
// (float)((double)S1000 / 1000.0) reproduces it unless the value sits within the f64
// accumulation error of a float32 rounding boundary; those candidates, and every
// SSD candidate, are replayed in float64 in the reference's operation order.
#include <stdlib.h>

#include "memci_internal.cuh"

struct BwRes { float r, c, cost; };

__device__ __forceinline__ float hbma_cost(const int32_t *__restrict__ L1, const int32_t *__restrict__ L2,
                                           const uint8_t *__restrict__ rgb1, const uint8_t *__restrict__ rgb2,
                                           int H, int W, int Wp, int ar, int ac, int br, int bc, int bs,
                                           int cost_key, double eps)
{
    if (cost_key == 0) {
        unsigned S = sad1000_global(L1, L2, H, W, Wp, ar, ac, br, bc, bs);
        double v = (double)S / 1000.0;
        float lo = (float)(v - eps), hi = (float)(v + eps);
        if (lo == hi) return (float)v;
    }
    return (float)exact_cost_f64(rgb1, rgb2, H, W, ar, ac, br, bc, bs, cost_key);
}

// block_wise_fs (hbma.py:34-75), executed by one warp.  (ar, ac) = top-left of block1 in
// frame 1, (idx_r, idx_c) = search centre in frame 2.  All lanes return the result.
__device__ inline BwRes bwfs_warp(const int32_t *__restrict__ L1, const int32_t *__restrict__ L2,
                                  const uint8_t *__restrict__ rgb1, const uint8_t *__restrict__ rgb2,
                                  int H, int W, int Wp, int ar, int ac, int idx_r, int idx_c, int win, int bs,
                                  int cost_key, double eps, int lane)
{
    BwRes res;
    if (idx_r >= H || idx_c >= W || idx_r < 0 || idx_c < 0) {           // :36-37
        res.r = 0.f; res.c = 0.f; res.cost = __int_as_float(0x7f800000);
        return res;
    }
    int max_r = win + 1, min_r = -win, max_c = win + 1, min_c = -win;
    if (idx_r + win + bs >= H) max_r = max(1, H - idx_r - bs);           // :53-54
    if (idx_r - win < 0) min_r += win - idx_r;                           // :55-56
    if (idx_c + win + bs >= W) max_c = max(1, W - idx_c - bs);
    if (idx_c - win < 0) min_c += win - idx_c;
    const int nr = max_r - min_r, nc = max_c - min_c, n = nr * nc;
    unsigned long long best = 0xffffffffffffffffull;
    for (int c = lane; c < n; c += 32) {
        int ri = c / nc, ci = c - ri * nc;
        int r_off = min_r + ri, c_off = min_c + ci;
        float cost = hbma_cost(L1, L2, rgb1, rgb2, H, W, Wp, ar, ac, idx_r + r_off, idx_c + c_off, bs, cost_key, eps);
        unsigned d2 = (unsigned)(r_off * r_off + c_off * c_off);
        // cost >= 0: the float32 bit pattern orders like the value; ties on the f32 value fall
        // through to distance, then (r_off, c_off) scan order  (:69)
        unsigned long long k = ((unsigned long long)__float_as_uint(cost) << 32) |
                               ((unsigned long long)d2 << 16) | ((unsigned)ri << 8) | (unsigned)ci;
        best = k < best ? k : best;
    }
    best = shfl_min_u64(best);
    res.r = (float)(min_r + (int)((best >> 8) & 255));
    res.c = (float)(min_c + (int)(best & 255));
    res.cost = __uint_as_float((unsigned)(best >> 32));
    return res;
}


// float32 cost of a candidate from its integer SAD S (luma*1000 units): get_cost returns
// float32(f64 sum), and the f64 sum is within eps of S/1000.  fdiv_rn(S, 1000) is the correctly
// rounded float32 of S/1000; it is also the reference's value unless S/1000 lies within eps of the
// rounding boundary (half an ulp away from the result), in which case the caller replays in f64.
__device__ __forceinline__ bool cost_from_S(unsigned S, float eps, float &cost)
{
    if (S >= (1u << 24)) return false;
    const float Sf = (float)S;
    const float f = __fdiv_rn(Sf, 1000.0f);
    const float r = fmaf(-f, 1000.0f, Sf);                              // exact remainder S - 1000 f
    const float half_ulp = __int_as_float(max((__float_as_int(f) & 0x7f800000) - (24 << 23), 0));
    cost = f;
    return S == 0u || (half_ulp - fabsf(r) * 0.001f) > eps + half_ulp * 1e-3f;
}

// block_wise_fs (hbma.py:34-75) with the block and the search region staged in this warp's
// shared memory (zero filled outside the frame = np.pad).  Lanes = candidates; every lane returns
// the winner.  `sm` must hold bs*bs + (bs+2*win)^2 ints.
__device__ inline BwRes bwfs_warp_staged(int *sm, const int32_t *__restrict__ L1, const int32_t *__restrict__ L2,
                                         const uint8_t *__restrict__ rgb1, const uint8_t *__restrict__ rgb2,
                                         int H, int W, int Wp, int ar, int ac, int idx_r, int idx_c, int win, int bs,
                                         int cost_key, float eps, int lane, bool block_staged)
{
    BwRes res;
    if (idx_r >= H || idx_c >= W || idx_r < 0 || idx_c < 0) {           // :36-37
        res.r = 0.f; res.c = 0.f; res.cost = __int_as_float(0x7f800000);
        return res;
    }
    int max_r = win + 1, min_r = -win, max_c = win + 1, min_c = -win;
    if (idx_r + win + bs >= H) max_r = max(1, H - idx_r - bs);           // :53-54
    if (idx_r - win < 0) min_r += win - idx_r;                           // :55-56
    if (idx_c + win + bs >= W) max_c = max(1, W - idx_c - bs);
    if (idx_c - win < 0) min_c += win - idx_c;
    const int nr = max_r - min_r, nc = max_c - min_c, n = nr * nc;
    int *blk = sm, *reg = sm + bs * bs;
    const int rh = nr + bs - 1, rw = nc + bs - 1;
    if (!block_staged) {
        for (int i = lane; i < bs * bs; i += 32) {
            const int r = ar + i / bs, c = ac + i % bs;
            blk[i] = (r < H && c < W) ? L1[(size_t)r * Wp + c] : 0;
        }
    }
    __syncwarp();
    const int r0 = idx_r + min_r, c0 = idx_c + min_c;                    // >= 0 by construction
    for (int i = lane; i < rh * rw; i += 32) {
        const int r = r0 + i / rw, c = c0 + i % rw;
        reg[i] = (r < H && c < W) ? L2[(size_t)r * Wp + c] : 0;
    }
    __syncwarp();
    unsigned long long best = 0xffffffffffffffffull;
    for (int cand = lane; cand < n; cand += 32) {
        const int ri = cand / nc, ci = cand - ri * nc;
        const int r_off = min_r + ri, c_off = min_c + ci;
        float cost;
        bool ok = false;
        if (cost_key == 0) {
            unsigned S = 0;
            const int *rp = reg + ri * rw + ci;
            for (int i = 0; i < bs; ++i)
                for (int j = 0; j < bs; ++j) S = __sad(blk[i * bs + j], rp[i * rw + j], S);
            ok = cost_from_S(S, eps, cost);
        }
        if (!ok) cost = (float)exact_cost_f64(rgb1, rgb2, H, W, ar, ac, idx_r + r_off, idx_c + c_off, bs, cost_key);
        const unsigned d2 = (unsigned)(r_off * r_off + c_off * c_off);
        const unsigned long long k = ((unsigned long long)__float_as_uint(cost) << 32) |
                                     ((unsigned long long)d2 << 16) | ((unsigned)ri << 8) | (unsigned)ci;
        best = k < best ? k : best;
    }
    best = shfl_min_u64(best);
    __syncwarp();
    res.r = (float)(min_r + (int)((best >> 8) & 255));
    res.c = (float)(min_c + (int)(best & 255));
    res.cost = __uint_as_float((unsigned)(best >> 32));
    return res;
}

#define HBMA_WARPS 8
__device__ __forceinline__ BwRes bwfs_any(int *sm, int sm_ints, const int32_t *L1, const int32_t *L2, const uint8_t *rgb1,
                                          const uint8_t *rgb2, int H, int W, int Wp, int ar, int ac, int idx_r, int idx_c,
                                          int win, int bs, int cost_key, double eps, int lane, bool block_staged)
{
    const int need = bs * bs + (bs + 2 * win) * (bs + 2 * win);
    if (need <= sm_ints)
        return bwfs_warp_staged(sm, L1, L2, rgb1, rgb2, H, W, Wp, ar, ac, idx_r, idx_c, win, bs, cost_key, (float)eps, lane, block_staged);
    return bwfs_warp(L1, L2, rgb1, rgb2, H, W, Wp, ar, ac, idx_r, idx_c, win, bs, cost_key, eps, lane);
}

// full_search_jit (hbma.py:77-90): block-level field on the coarsest level.
__global__ void hbma_fs_kernel(const int32_t *__restrict__ L1, const int32_t *__restrict__ L2,
                               const uint8_t *__restrict__ rgb1, const uint8_t *__restrict__ rgb2,
                               int H, int W, int Wp, int bs, int win, int cost_key, double eps,
                               int rows, int cols, float2 *__restrict__ vec, float *__restrict__ cost, int sm_ints)
{
    extern __shared__ int hbma_sm[];
    int *sm = hbma_sm + (threadIdx.x >> 5) * sm_ints;
    int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int blk = warp; blk < rows * cols; blk += nwarps) {
        int row = blk / cols, col = blk - row * cols;
        BwRes r = bwfs_any(sm, sm_ints, L1, L2, rgb1, rgb2, H, W, Wp, row * bs, col * bs, row * bs, col * bs, win, bs,
                           cost_key, eps, lane, false);
        if (lane == 0) { vec[blk] = make_float2(r.r, r.c); cost[blk] = r.cost; }
    }
}

// increase_vec_density_jit (hbma.py:100-140): one warp per child block of the sliced output.
__global__ void hbma_densify_kernel(const int32_t *__restrict__ L1, const int32_t *__restrict__ L2,
                                    const uint8_t *__restrict__ rgb1, const uint8_t *__restrict__ rgb2,
                                    int H, int W, int Wp, int bs, int sub_win, int cost_key, double eps, int vec_scale,
                                    const float2 *__restrict__ in_vec, int in_rows, int in_cols, int in_pitch,
                                    float2 *__restrict__ out_vec, float *__restrict__ out_cost,
                                    int out_rows, int out_cols, int out_pitch, int sm_ints)
{
    extern __shared__ int hbma_sm[];
    int *sm = hbma_sm + (threadIdx.x >> 5) * sm_ints;
    int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int blk = warp; blk < out_rows * out_cols; blk += nwarps) {
        int o_r = blk / out_cols, o_c = blk - o_r * out_cols;
        int row = o_r >> 1, col = o_c >> 1;
        float2 o = make_float2(0.f, 0.f);
        float oc = 0.f;
        bool evaluated = row < in_rows && col < in_cols;
        if (evaluated) {
            // second child row/col is skipped when the parent touches the frame edge (:117-124, Q5)
            int r_max = ((row + 1) * 2 * bs >= H) ? row * 2 + 1 : (row + 1) * 2;
            int c_max = ((col + 1) * 2 * bs >= W) ? col * 2 + 1 : (col + 1) * 2;
            evaluated = o_r < r_max && o_c < c_max;
        }
        if (evaluated) {
            float2 v[3];
            v[0] = in_vec[(size_t)row * in_pitch + col];                          // parent, unscaled (:107, Q4)
            int hc = col >= 1 ? col - 1 : min(col + 1, in_cols - 1);
            int vr = row >= 1 ? row - 1 : min(row + 1, in_rows - 1);
            float2 h = in_vec[(size_t)row * in_pitch + hc];
            float2 u = in_vec[(size_t)vr * in_pitch + col];
            float sc = (float)vec_scale;
            v[1] = make_float2(__fmul_rn(h.x, sc), __fmul_rn(h.y, sc));
            v[2] = make_float2(__fmul_rn(u.x, sc), __fmul_rn(u.y, sc));
            int im_r = o_r * bs, im_c = o_c * bs;
            float lowest = __int_as_float(0x7f800000);
            if (sm_ints > 0) {                       // stage the child block once for its three searches
                for (int i = lane; i < bs * bs; i += 32) {
                    const int r = im_r + i / bs, c = im_c + i % bs;
                    sm[i] = (r < H && c < W) ? L1[(size_t)r * Wp + c] : 0;
                }
                __syncwarp();
            }
            const int rdim = bs + 2 * sub_win;
            if (sm_ints >= bs * bs + 3 * rdim * rdim) {
                // ---- fused: the three searches' candidates (<= 3 (2 sub+1)^2) share one pass over the lanes ----
                int idr[3], idc[3], mnr[3], mnc[3], nr[3], nc[3], off[4];
                off[0] = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    idr[k] = (int)((double)im_r + (double)v[k].x);                // f64 sum, trunc to int32 (:134)
                    idc[k] = (int)((double)im_c + (double)v[k].y);
                    const bool inside = !(idr[k] >= H || idc[k] >= W || idr[k] < 0 || idc[k] < 0);   // :36-37
                    int max_r = sub_win + 1, min_r = -sub_win, max_c = sub_win + 1, min_c = -sub_win;
                    if (idr[k] + sub_win + bs >= H) max_r = max(1, H - idr[k] - bs);
                    if (idr[k] - sub_win < 0) min_r += sub_win - idr[k];
                    if (idc[k] + sub_win + bs >= W) max_c = max(1, W - idc[k] - bs);
                    if (idc[k] - sub_win < 0) min_c += sub_win - idc[k];
                    mnr[k] = min_r; mnc[k] = min_c;
                    nr[k] = inside ? max_r - min_r : 0; nc[k] = inside ? max_c - min_c : 0;
                    off[k + 1] = off[k] + nr[k] * nc[k];
                    if (inside) {
                        int *reg = sm + bs * bs + k * rdim * rdim;
                        const int rh = nr[k] + bs - 1, rw = nc[k] + bs - 1;
                        const int r0 = idr[k] + min_r, c0 = idc[k] + min_c;
                        for (int i = lane; i < rh * rw; i += 32) {
                            const int r = r0 + i / rw, c = c0 + i % rw;
                            reg[i] = (r < H && c < W) ? L2[(size_t)r * Wp + c] : 0;
                        }
                    }
                }
                __syncwarp();
                unsigned long long best[3] = {~0ull, ~0ull, ~0ull};
                for (int cand = lane; cand < off[3]; cand += 32) {
                    const int k = cand >= off[2] ? 2 : (cand >= off[1] ? 1 : 0);
                    const int loc = cand - off[k];
                    const int ri = loc / nc[k], ci = loc - ri * nc[k];
                    const int r_off = mnr[k] + ri, c_off = mnc[k] + ci;
                    const int rw = nc[k] + bs - 1;
                    float cost;
                    bool ok = false;
                    if (cost_key == 0) {
                        unsigned S = 0;
                        const int *rp = sm + bs * bs + k * rdim * rdim + ri * rw + ci;
                        for (int i = 0; i < bs; ++i)
                            for (int j = 0; j < bs; ++j) S = __sad(sm[i * bs + j], rp[i * rw + j], S);
                        ok = cost_from_S(S, (float)eps, cost);
                    }
                    if (!ok) cost = (float)exact_cost_f64(rgb1, rgb2, H, W, im_r, im_c, idr[k] + r_off, idc[k] + c_off, bs, cost_key);
                    const unsigned long long key = ((unsigned long long)__float_as_uint(cost) << 32) |
                                                   ((unsigned long long)(unsigned)(r_off * r_off + c_off * c_off) << 16) |
                                                   ((unsigned)ri << 8) | (unsigned)ci;
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk) if (kk == k) best[kk] = key < best[kk] ? key : best[kk];
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const unsigned long long bk = shfl_min_u64(best[k]);
                    // out-of-frame search centre returns [0, 0, inf] and is never selected (:36-37, :135)
                    const float cost = nr[k] > 0 ? __uint_as_float((unsigned)(bk >> 32)) : __int_as_float(0x7f800000);
                    if (cost < lowest) {                                          // strict: first vector wins ties (:135)
                        lowest = cost;
                        oc = cost;
                        o = make_float2(__fadd_rn((float)(mnr[k] + (int)((bk >> 8) & 255)), v[k].x),
                                        __fadd_rn((float)(mnc[k] + (int)(bk & 255)), v[k].y));
                    }
                }
                __syncwarp();
            } else {
#pragma unroll 1
                for (int k = 0; k < 3; ++k) {
                    int ir = (int)((double)im_r + (double)v[k].x);                // f64 sum, trunc to int32 (:134)
                    int ic = (int)((double)im_c + (double)v[k].y);
                    BwRes r = bwfs_any(sm, sm_ints, L1, L2, rgb1, rgb2, H, W, Wp, im_r, im_c, ir, ic, sub_win, bs, cost_key, eps, lane, true);
                    if (r.cost < lowest) {                                        // strict: first vector wins ties (:135)
                        lowest = r.cost;
                        oc = r.cost;
                        o = make_float2(__fadd_rn(r.r, v[k].x), __fadd_rn(r.c, v[k].y));
                    }
                }
            }
        }
        if (lane == 0) {
            out_vec[(size_t)o_r * out_pitch + o_c] = o;
            out_cost[(size_t)o_r * out_pitch + o_c] = oc;
        }
    }
}

// ---- increase_vec_density, fast path (sub_win == 1, SAD, block size 4 or 8) -------------------------
// One warp per PARENT block: its four children share the three candidate vectors (parent unscaled,
// horizontal neighbour, vertical neighbour -- hbma.py:107-115), so the 2BS x 2BS source tile and one
// (2BS+2)^2 target region per vector are staged once (zero filled outside the frame = np.pad) and the
// 4 x 3 x 9 = 108 candidates run as four passes of 27 lanes.  All index arithmetic is compile-time
// (the generic kernel spends most of its ~1000 instructions per child on runtime divisions and staging).
// Same arithmetic as the generic kernel: float32 cost from the integer SAD unless it sits on a rounding
// boundary (then the f64 replay), key = (cost bits, d2, ri, ci), strict '<' across the three vectors.
template <int BS>
__global__ void __launch_bounds__(HBMA_WARPS * 32)
hbma_densify_parent_kernel(const int32_t *__restrict__ L1, const int32_t *__restrict__ L2,
                           const uint8_t *__restrict__ rgb1, const uint8_t *__restrict__ rgb2,
                           int H, int W, int Wp, double eps, int vec_scale,
                           const float2 *__restrict__ in_vec, int in_rows, int in_cols, int in_pitch,
                           float2 *__restrict__ out_vec, float *__restrict__ out_cost,
                           int out_rows, int out_cols, int out_pitch)
{
    constexpr int SB = 2 * BS, RS = 2 * BS + 2, RSP = RS + 1, SM_INTS = SB * SB + 3 * RS * RSP;
    __shared__ int sm_all[HBMA_WARPS][SM_INTS];
    int *src = sm_all[threadIdx.x >> 5], *reg = src + SB * SB;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int prow_n = (out_rows + 1) >> 1, pcol_n = (out_cols + 1) >> 1;
    const float epsf = (float)eps;
    // candidate of this lane inside a pass: vector k, refinement offset (ri, ci) of the 3x3 window
    const int k_l = lane / 9, m_l = lane - k_l * 9, ri_l = m_l / 3, ci_l = m_l - ri_l * 3;
    for (int pb = warp; pb < prow_n * pcol_n; pb += nwarps) {
        const int row = pb / pcol_n, col = pb - row * pcol_n;
        const bool parent_ok = row < in_rows && col < in_cols;
        int r_max = 0, c_max = 0;
        float2 v[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        const int im_r0 = row * SB, im_c0 = col * SB;
        int R0[3] = {0, 0, 0}, C0[3] = {0, 0, 0};
        if (parent_ok) {
            // second child row/col is skipped when the parent touches the frame edge (:117-124, Q5)
            r_max = ((row + 1) * SB >= H) ? row * 2 + 1 : (row + 1) * 2;
            c_max = ((col + 1) * SB >= W) ? col * 2 + 1 : (col + 1) * 2;
            v[0] = in_vec[(size_t)row * in_pitch + col];                          // parent, unscaled (:107, Q4)
            const int hc = col >= 1 ? col - 1 : min(col + 1, in_cols - 1);
            const int vr = row >= 1 ? row - 1 : min(row + 1, in_rows - 1);
            const float2 h = in_vec[(size_t)row * in_pitch + hc];
            const float2 u = in_vec[(size_t)vr * in_pitch + col];
            const float sc = (float)vec_scale;
            v[1] = make_float2(__fmul_rn(h.x, sc), __fmul_rn(h.y, sc));
            v[2] = make_float2(__fmul_rn(u.x, sc), __fmul_rn(u.y, sc));
            for (int i = lane; i < SB * SB; i += 32) {
                const int r = im_r0 + i / SB, c = im_c0 + i % SB;
                src[i] = (r < H && c < W) ? L1[(size_t)r * Wp + c] : 0;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                R0[k] = (int)((double)im_r0 + (double)v[k].x) - 1;                // search centre of child (0,0), minus sub_win
                C0[k] = (int)((double)im_c0 + (double)v[k].y) - 1;
                for (int i = lane; i < RS * RS; i += 32) {
                    const int rr = i / RS, cc = i - rr * RS;
                    const int r = R0[k] + rr, c = C0[k] + cc;
                    reg[k * RS * RSP + rr * RSP + cc] = (r >= 0 && c >= 0 && r < H && c < W) ? L2[(size_t)r * Wp + c] : 0;
                }
            }
        }
        __syncwarp();
#pragma unroll 1
        for (int child = 0; child < 4; ++child) {
            const int a = child >> 1, b = child & 1;
            const int o_r = row * 2 + a, o_c = col * 2 + b;
            if (o_r >= out_rows || o_c >= out_cols) continue;                     // warp-uniform
            const bool evaluated = parent_ok && o_r < r_max && o_c < c_max;
            float2 o = make_float2(0.f, 0.f);
            float oc = 0.f;
            if (evaluated) {
                const int im_r = o_r * BS, im_c = o_c * BS;
                unsigned hi = 0xffffffffu, lo = 0xffffffffu;
                int mn_r = 0, mn_c = 0;
                if (lane < 27) {
                    const float2 vk = k_l == 0 ? v[0] : (k_l == 1 ? v[1] : v[2]);
                    const int idr = (int)((double)im_r + (double)vk.x);           // f64 sum, trunc to int32 (:134)
                    const int idc = (int)((double)im_c + (double)vk.y);
                    if (!(idr >= H || idc >= W || idr < 0 || idc < 0)) {          // :36-37
                        int max_r = 2, min_r = -1, max_c = 2, min_c = -1;         // sub_win == 1
                        if (idr + 1 + BS >= H) max_r = max(1, H - idr - BS);      // :53-54
                        if (idr - 1 < 0) min_r += 1 - idr;                        // :55-56
                        if (idc + 1 + BS >= W) max_c = max(1, W - idc - BS);
                        if (idc - 1 < 0) min_c += 1 - idc;
                        mn_r = min_r; mn_c = min_c;
                        if (ri_l < max_r - min_r && ci_l < max_c - min_c) {
                            const int r_off = min_r + ri_l, c_off = min_c + ci_l;
                            const int R0k = k_l == 0 ? R0[0] : (k_l == 1 ? R0[1] : R0[2]);
                            const int C0k = k_l == 0 ? C0[0] : (k_l == 1 ? C0[1] : C0[2]);
                            const int rr0 = idr + r_off - R0k, cc0 = idc + c_off - C0k;      // inside the staged region for integral vectors
                            float cost;
                            bool ok = false;
                            if (rr0 >= 0 && cc0 >= 0 && rr0 + BS <= RS && cc0 + BS <= RS) {
                                const int *sp = src + (a * BS) * SB + b * BS;
                                const int *rp = reg + k_l * RS * RSP + rr0 * RSP + cc0;
                                unsigned S = 0;
#pragma unroll
                                for (int i = 0; i < BS; ++i)
#pragma unroll
                                    for (int j = 0; j < BS; ++j) S = __sad(sp[i * SB + j], rp[i * RSP + j], S);
                                ok = cost_from_S(S, epsf, cost);
                            }
                            if (!ok) cost = (float)exact_cost_f64(rgb1, rgb2, H, W, im_r, im_c, idr + r_off, idc + c_off, BS, 0);
                            hi = __float_as_uint(cost);
                            lo = ((unsigned)(r_off * r_off + c_off * c_off) << 16) | ((unsigned)ri_l << 8) | (unsigned)ci_l;
                        }
                    }
                }
                float lowest = __int_as_float(0x7f800000);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const bool mine = k_l == k;
                    const unsigned mh = __reduce_min_sync(0xffffffffu, mine ? hi : 0xffffffffu);
                    const unsigned ml = __reduce_min_sync(0xffffffffu, (mine && hi == mh) ? lo : 0xffffffffu);
                    // the winner's lane also knows its window origin (min_r, min_c): fetch it from that lane
                    const unsigned who = __ballot_sync(0xffffffffu, mine && hi == mh && lo == ml && hi != 0xffffffffu);
                    if (who) {                                                    // out-of-frame centre: [0,0,inf], never selected (:36-37, :135)
                        const int wl = __ffs(who) - 1;
                        const int wr = __shfl_sync(0xffffffffu, mn_r, wl), wc = __shfl_sync(0xffffffffu, mn_c, wl);
                        const float cost = __uint_as_float(mh);
                        if (cost < lowest) {                                      // strict: first vector wins ties (:135)
                            lowest = cost;
                            oc = cost;
                            o = make_float2(__fadd_rn((float)(wr + (int)((ml >> 8) & 255u)), v[k].x),
                                            __fadd_rn((float)(wc + (int)(ml & 255u)), v[k].y));
                        }
                    }
                }
            }
            if (lane == 0) {
                out_vec[(size_t)o_r * out_pitch + o_c] = o;
                out_cost[(size_t)o_r * out_pitch + o_c] = oc;
            }
        }
        __syncwarp();
    }
}

__global__ void hbma_compact_kernel(const float2 *__restrict__ in_vec, const float *__restrict__ in_cost, int pitch,
                                    int rows, int cols, float2 *__restrict__ out_vec, float *__restrict__ out_sad)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    int r = i / cols, c = i - r * cols;
    out_vec[i] = in_vec[(size_t)r * pitch + c];
    out_sad[i] = in_cost[(size_t)r * pitch + c];
}

static double hbma_eps(int bs)
{
    // bound on |f64 sequential sum - exact sum| for bs*bs terms of magnitude <= 255 (x2 margin)
    double n = (double)bs * bs;
    return 2.0 * n * (2e-13 + 2.9e-14 * n);
}

int memci_hbma_run(memci_ctx *ctx, int slot_src, int slot_tgt, int bs, int win, int sub, int steps,
                   int min_bs, int cost_key, MVField *out)
{
    if (bs <= 0 || bs > 64 || win < 0 || win > 127 || sub < 0 || sub > 127 || steps < 0 || steps > MEMCI_MAX_PYR ||
        min_bs <= 0 || (cost_key != 0 && cost_key != 1))
        return memci_fail(ctx, MEMCI_ERR_ARG, "hbma: argument out of range (bs %d win %d sub %d steps %d min_bs %d cost %d)",
                          bs, win, sub, steps, min_bs, cost_key);
    int rc;
    if ((rc = memci_ensure_luma(ctx, slot_src))) return rc;
    if ((rc = memci_ensure_luma(ctx, slot_tgt))) return rc;
    if ((rc = memci_ensure_pyramid(ctx, slot_src, steps))) return rc;
    if ((rc = memci_ensure_pyramid(ctx, slot_tgt, steps))) return rc;
    FrameSlot &f1 = ctx->frames[slot_src], &f2 = ctx->frames[slot_tgt];
    int hs[MEMCI_MAX_PYR + 1], ws[MEMCI_MAX_PYR + 1];
    hs[0] = ctx->H; ws[0] = ctx->W;
    for (int l = 1; l <= steps; ++l) { hs[l] = (hs[l - 1] + 1) / 2; ws[l] = (ws[l - 1] + 1) / 2; }
    auto LUMA = [&](FrameSlot &f, int l) { return l == 0 ? f.luma : f.pyr_luma[l]; };
    auto RGB = [&](FrameSlot &f, int l) { return l == 0 ? f.rgb : f.pyr_rgb[l]; };
    auto WP = [&](int l) { return l == 0 ? ctx->Wp : ((ws[l] + 3) & ~3); };

    // worst-case intermediate size: the last densification output at the final block size
    int fbs = bs;
    while (fbs > min_bs) fbs >>= 1;
    if (fbs <= 0) return memci_fail(ctx, MEMCI_ERR_ARG, "hbma: min_block_size %d unreachable from %d", min_bs, bs);
    size_t max_cells = 0;
    {
        int r = ceil_div_i(hs[steps], bs), c = ceil_div_i(ws[steps], bs);
        max_cells = (size_t)r * c;
        for (int l = steps - 1; l >= 0; --l) {
            size_t n = (size_t)(2 * r) * (2 * c);
            if (n > max_cells) max_cells = n;
            r = ceil_div_i(hs[l], bs); c = ceil_div_i(ws[l], bs);
        }
        int b = bs;
        while (b > min_bs) {
            b >>= 1;
            size_t n = (size_t)(2 * r) * (2 * c);
            if (n > max_cells) max_cells = n;
            r = ceil_div_i(ctx->H, b); c = ceil_div_i(ctx->W, b);
        }
    }
    for (int i = 0; i < 2; ++i)
        if ((rc = memci_mv_reserve(ctx, &ctx->tmp_mv[i], 1, 1, (int)max_cells, 1, 1, (int)max_cells))) return rc;

    const int grid = ctx->num_sms * 8, nt = HBMA_WARPS * 32;
    auto sm_ints_for = [&](int b, int w) {               // per-warp staging: block + search region, capped at 6 KB
        int need = b * b + (b + 2 * w) * (b + 2 * w);
        return need <= 1536 ? need : 0;
    };
    int cur = 0;
    int rows = ceil_div_i(hs[steps], bs), cols = ceil_div_i(ws[steps], bs), pitch = cols;
    // Algorithmic work of the whole call in cost pixel-ops (one |La - Lb| or (La - Lb)^2 accumulate), NOMINAL counts: the
    // coarse level searches (2 win + 1)^2 candidates per block, every densification 3 vectors x (2 sub + 1)^2 candidates per
    // child block (hbma.py:100-140); clipping at the frame border removes about 1 % of them and is not subtracted.
    double work = (double)rows * cols * (2 * win + 1) * (2 * win + 1) * bs * bs;
    {
        int r = rows, c = cols;
        for (int l = steps - 1; l >= 0; --l) { r = ceil_div_i(hs[l], bs); c = ceil_div_i(ws[l], bs); work += 3.0 * r * c * (2 * sub + 1) * (2 * sub + 1) * bs * bs; }
        for (int b2 = bs >> 1; b2 >= min_bs && b2 > 0; b2 >>= 1) { r = ceil_div_i(ctx->H, b2); c = ceil_div_i(ctx->W, b2); work += 3.0 * r * c * (2 * sub + 1) * (2 * sub + 1) * b2 * b2; }
    }
    const int prof = memci_prof_begin(ctx, MEMCI_PROF_HBMA, work);
    {
        const int smi = sm_ints_for(bs, win);
        hbma_fs_kernel<<<grid, nt, smi * HBMA_WARPS * sizeof(int), ctx->stream>>>(
            LUMA(f1, steps), LUMA(f2, steps), RGB(f1, steps), RGB(f2, steps), hs[steps], ws[steps], WP(steps), bs, win,
            cost_key, hbma_eps(bs), rows, cols, ctx->tmp_mv[cur].vec, ctx->tmp_mv[cur].sad, smi);
    }
    CKL(ctx);
    auto densify = [&](int lvl, int b, int vec_scale) -> int {
        int orows = ceil_div_i(hs[lvl], b), ocols = ceil_div_i(ws[lvl], b);   // the slice kept by the reference
        int opitch = cols * 2;
        if (sub == 1 && cost_key == 0 && (b == 4 || b == 8) && !getenv("MEMCI_HBMA_GENERIC")) {     // parent-level fast path
            if (b == 8)
                hbma_densify_parent_kernel<8><<<grid, nt, 0, ctx->stream>>>(
                    LUMA(f1, lvl), LUMA(f2, lvl), RGB(f1, lvl), RGB(f2, lvl), hs[lvl], ws[lvl], WP(lvl), hbma_eps(b), vec_scale,
                    ctx->tmp_mv[cur].vec, rows, cols, pitch, ctx->tmp_mv[cur ^ 1].vec, ctx->tmp_mv[cur ^ 1].sad, orows, ocols, opitch);
            else
                hbma_densify_parent_kernel<4><<<grid, nt, 0, ctx->stream>>>(
                    LUMA(f1, lvl), LUMA(f2, lvl), RGB(f1, lvl), RGB(f2, lvl), hs[lvl], ws[lvl], WP(lvl), hbma_eps(b), vec_scale,
                    ctx->tmp_mv[cur].vec, rows, cols, pitch, ctx->tmp_mv[cur ^ 1].vec, ctx->tmp_mv[cur ^ 1].sad, orows, ocols, opitch);
            CKL(ctx);
            cur ^= 1; rows = orows; cols = ocols; pitch = opitch;
            return MEMCI_OK;
        }
        int smi = b * b + 3 * (b + 2 * sub) * (b + 2 * sub);          // fused three-vector search
        if (smi > 1536) smi = sm_ints_for(b, sub);
        hbma_densify_kernel<<<grid, nt, smi * HBMA_WARPS * sizeof(int), ctx->stream>>>(
            LUMA(f1, lvl), LUMA(f2, lvl), RGB(f1, lvl), RGB(f2, lvl), hs[lvl], ws[lvl], WP(lvl), b, sub, cost_key, hbma_eps(b),
            vec_scale, ctx->tmp_mv[cur].vec, rows, cols, pitch, ctx->tmp_mv[cur ^ 1].vec, ctx->tmp_mv[cur ^ 1].sad,
            orows, ocols, opitch, smi);
        CKL(ctx);
        cur ^= 1; rows = orows; cols = ocols; pitch = opitch;
        return MEMCI_OK;
    };
    for (int lvl = steps - 1; lvl >= 0; --lvl)
        if ((rc = densify(lvl, bs, 2))) return rc;                               // hbma.py:181-182
    int b = bs;
    while (b > min_bs) {                                                         // hbma.py:184-186
        b >>= 1;
        if ((rc = densify(0, b, 1))) return rc;
    }
    memci_prof_end(ctx, prof);
    if ((rc = memci_mv_reserve(ctx, out, b, rows, cols, b, rows, cols))) return rc;
    hbma_compact_kernel<<<ceil_div_i(rows * cols, 256), 256, 0, ctx->stream>>>(ctx->tmp_mv[cur].vec, ctx->tmp_mv[cur].sad,
                                                                               pitch, rows, cols, out->vec, out->sad);
    CKL(ctx);
    out->valid = 1;
    return MEMCI_OK;
}
