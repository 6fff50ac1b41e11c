/*
 * memci_oracle.c -- CPU restatement of the MEMCI hot path of lhl2617/Interpolatory.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library,
 * and only as the checker / the reported CPU baseline.  The product path
 * (interpolatory_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  Every function here was checked bit-for-bit against
 * the unmodified reference executed in the build container (numba 0.65.0,
 * NumPy 2.3.5, SciPy 1.18.1) by oracle/gen_golden.py, and the resulting
 * input/output vectors are committed under tests/golden/ (the reference has
 * no tests or golden vectors of its own -- SURVEY.md section 4).
 *
 * All file:line citations are relative to
 *   /root/reference/Simulator/python/src/Interpolators/MEMCI/   (unless noted).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: the reference (numba, no fastmath) never fuses a
 * multiply and an add, and the float64 luma/SAD below must round identically.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* helpers                                                                   */
/* ------------------------------------------------------------------------ */

/* np.pad(frame, ((0,bs),(0,bs),(0,0))) -- ME/full_search.py:66-67, ME/hbma.py:93-94,144-145 */
static uint8_t *pad_frame(const uint8_t *im, int H, int W, int bs)
{
    int Hp = H + bs, Wp = W + bs;
    uint8_t *p = (uint8_t *)calloc((size_t)Hp * Wp * 3, 1);
    if (!p) return NULL;
    for (int r = 0; r < H; ++r)
        memcpy(p + (size_t)r * Wp * 3, im + (size_t)r * W * 3, (size_t)W * 3);
    return p;
}

/* Luma as numba evaluates `0.299*x[:,:,0] + 0.587*x[:,:,1] + 0.114*x[:,:,2]` on a
 * float32 array: scalar f64 * f32 array promotes to f64, separate mul and add,
 * left-to-right.  ME/full_search.py:13-16, ME/hbma.py:18-21. */
static inline double luma64(const uint8_t *px)
{
    double r = (double)(float)px[0], g = (double)(float)px[1], b = (double)(float)px[2];
    double t0 = 0.299 * r;
    double t1 = 0.587 * g;
    double t2 = 0.114 * b;
    double s = t0 + t1;
    return s + t2;
}

/* Sequential row-major f64 sum of |L1 - L2| over a bs x bs block
 * (np.sum in numba = flat loop from 0.0).  ME/full_search.py:17, ME/hbma.py:22.
 * a and b point at the top-left pixel of each block inside padded frames of
 * row pitch (in pixels) pa / pb. */
static double sad_f64(const uint8_t *a, int pa, const uint8_t *b, int pb, int bs)
{
    double acc = 0.0;
    for (int i = 0; i < bs; ++i) {
        const uint8_t *ra = a + (size_t)i * pa * 3;
        const uint8_t *rb = b + (size_t)i * pb * 3;
        for (int j = 0; j < bs; ++j) {
            double d = luma64(ra + 3 * j) - luma64(rb + 3 * j);
            acc += fabs(d);
        }
    }
    return acc;
}

/* SSD variant (cost_key 1), ME/hbma.py:23-28. */
static double ssd_f64(const uint8_t *a, int pa, const uint8_t *b, int pb, int bs)
{
    double acc = 0.0;
    for (int i = 0; i < bs; ++i) {
        const uint8_t *ra = a + (size_t)i * pa * 3;
        const uint8_t *rb = b + (size_t)i * pb * 3;
        for (int j = 0; j < bs; ++j) {
            double d = luma64(ra + 3 * j) - luma64(rb + 3 * j);
            acc += d * d;
        }
    }
    return acc;
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

/* ------------------------------------------------------------------------ */
/* Full search  (ME/full_search.py:19-55 helper, :61-74 get_motion_vectors)  */
/* ------------------------------------------------------------------------ */

/* Candidate window of block (s_row, s_col); reproduces the row/col bug at :37
 * (column bound tested against frame_shape[0] = H) and the exclusive bottom
 * bound at :36/:42. */
static void fs_bounds(int H, int W, int bs, int R, int s_row, int s_col,
                      int *r0, int *r1, int *c0, int *c1)
{
    int tmr = (s_row + bs >= H) ? s_row + 1 : H - bs;
    int tmc = (s_col + bs >= H) ? s_col + 1 : W - bs;   /* sic: H, not W */
    *r0 = imax(0, s_row - R);
    *r1 = imin(tmr, s_row + R + 1);
    *c0 = imax(0, s_col - R);
    *c1 = imin(tmc, s_col + R + 1);
}

/* Block-level output: out[(br*nbc + bc)*3 + {0,1,2}] = (dy, dx, sad) as float32,
 * nbr = ceil(H/bs), nbc = ceil(W/bs).  The reference broadcasts each triple over
 * the block's pixels (:51-53); orc_expand_dense() does that. */
ORC_API int orc_fs_blocks(const uint8_t *im1, const uint8_t *im2, int H, int W,
                          int bs, int R, float *out)
{
    if (H <= 0 || W <= 0 || bs <= 0 || R < 0) return -1;
    uint8_t *sp = pad_frame(im1, H, W, bs), *tp = pad_frame(im2, H, W, bs);
    if (!sp || !tp) { free(sp); free(tp); return -2; }
    int Wp = W + bs;
    int nbr = ceil_div(H, bs), nbc = ceil_div(W, bs);
#pragma omp parallel for schedule(dynamic, 1)
    for (int br = 0; br < nbr; ++br) {
        int s_row = br * bs;
        for (int bc = 0; bc < nbc; ++bc) {
            int s_col = bc * bs;
            int r0, r1, c0, c1;
            fs_bounds(H, W, bs, R, s_row, s_col, &r0, &r1, &c0, &c1);
            double lowest_sad = INFINITY, lowest_dist = INFINITY;
            int ti_r = 0, ti_c = 0;                       /* target_index = (0, 0)  :31 */
            const uint8_t *sb = sp + ((size_t)s_row * Wp + s_col) * 3;
            for (int t_row = r0; t_row < r1; ++t_row) {
                for (int t_col = c0; t_col < c1; ++t_col) {
                    const uint8_t *tb = tp + ((size_t)t_row * Wp + t_col) * 3;
                    /* uint32(f64 sum) = truncation  :10 (N2) */
                    uint32_t sad = (uint32_t)sad_f64(sb, Wp, tb, Wp, bs);
                    long long dr = s_row - t_row, dc = s_col - t_col;
                    double distance = (double)(dr * dr + dc * dc);
                    if ((double)sad < lowest_sad ||
                        ((double)sad == lowest_sad && distance < lowest_dist)) {   /* :47 */
                        lowest_dist = distance;
                        lowest_sad = (double)sad;
                        ti_r = t_row; ti_c = t_col;
                    }
                }
            }
            float *o = out + ((size_t)br * nbc + bc) * 3;
            o[0] = (float)(ti_r - s_row);
            o[1] = (float)(ti_c - s_col);
            o[2] = (float)lowest_sad;                      /* inf if the window was empty */
        }
    }
    free(sp); free(tp);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Three-step search  (ME/tss.py:18-61 helper, :63-69 get_motion_vectors)     */
/* ------------------------------------------------------------------------ */
ORC_API int orc_tss_blocks(const uint8_t *im1, const uint8_t *im2, int H, int W,
                           int bs, int steps, float *out)
{
    static const int prec[9] = {1, 2, 1, 2, 3, 2, 1, 2, 1};                /* :21 */
    if (H <= 0 || W <= 0 || bs <= 0 || steps < 0 || steps > 20) return -1;
    uint8_t *sp = pad_frame(im1, H, W, bs), *tp = pad_frame(im2, H, W, bs);
    if (!sp || !tp) { free(sp); free(tp); return -2; }
    int Wp = W + bs;
    int nbr = ceil_div(H, bs), nbc = ceil_div(W, bs);
#pragma omp parallel for schedule(dynamic, 1)
    for (int br = 0; br < nbr; ++br)
        for (int bc = 0; bc < nbc; ++bc) {
            int s_row = br * bs, s_col = bc * bs;
            const uint8_t *sb = sp + ((size_t)s_row * Wp + s_col) * 3;
            int have_center = 0;                                            /* center_sad is None  :26 */
            double lowest_sad = INFINITY;
            int li_r = 0, li_c = 0, lowest_i = 0;
            int curr_row = s_row, curr_col = s_col;
            int tmr = (s_row + bs >= H) ? s_row + 1 : H - bs;               /* :36-39 */
            int tmc = (s_col + bs >= H) ? s_col + 1 : W - bs;               /* :40-43, sic: H */
            for (int step = steps; step > 0; --step) {
                int S = 1 << (step - 1);
                int i = -1;
                for (int t_row = curr_row - S; t_row < curr_row + S + 1; t_row += S)
                    for (int t_col = curr_col - S; t_col < curr_col + S + 1; t_col += S) {
                        ++i;
                        if ((t_row != curr_row || t_col != curr_col || !have_center) &&
                            (t_row >= 0 && t_row <= tmr && t_col >= 0 && t_col <= tmc)) {       /* :48 inclusive */
                            const uint8_t *tb = tp + ((size_t)t_row * Wp + t_col) * 3;
                            double sad = (double)(uint32_t)sad_f64(sb, Wp, tb, Wp, bs);
                            if (sad < lowest_sad || (sad == lowest_sad && prec[i] > prec[lowest_i])) {   /* :51 */
                                lowest_sad = sad; li_r = t_row; li_c = t_col; lowest_i = i;
                            }
                        }
                    }
                curr_row = li_r; curr_col = li_c;                           /* :55-57 */
                have_center = 1;
            }
            float *o = out + ((size_t)br * nbc + bc) * 3;
            o[0] = (float)(li_r - s_row); o[1] = (float)(li_c - s_col); o[2] = (float)lowest_sad;
        }
    free(sp); free(tp);
    return 0;
}

/* Exact algorithmic work of one full-search direction: number of candidates and
 * SAD pixel-ops (candidates * bs^2), from the reference's own bounds. */
ORC_API int orc_fs_work(int H, int W, int bs, int R, long long *n_cand, long long *px_ops)
{
    long long n = 0;
    for (int s_row = 0; s_row < H; s_row += bs)
        for (int s_col = 0; s_col < W; s_col += bs) {
            int r0, r1, c0, c1;
            fs_bounds(H, W, bs, R, s_row, s_col, &r0, &r1, &c0, &c1);
            if (r1 > r0 && c1 > c0) n += (long long)(r1 - r0) * (c1 - c0);
        }
    if (n_cand) *n_cand = n;
    if (px_ops) *px_ops = n * bs * bs;
    return 0;
}

/* Block -> dense broadcast: full_search.py:51-53, hbma.py:151-159 (upscale_mvs). */
ORC_API int orc_expand_dense(const float *blocks, int nbr, int nbc, int bs,
                             int H, int W, float *dense)
{
    memset(dense, 0, (size_t)H * W * 3 * sizeof(float));
    for (int br = 0; br < nbr; ++br)
        for (int bc = 0; bc < nbc; ++bc) {
            const float *b = blocks + ((size_t)br * nbc + bc) * 3;
            int r1 = imin(H, (br + 1) * bs), c1 = imin(W, (bc + 1) * bs);
            for (int r = br * bs; r < r1; ++r)
                for (int c = bc * bs; c < c1; ++c) {
                    float *d = dense + ((size_t)r * W + c) * 3;
                    d[0] = b[0]; d[1] = b[1]; d[2] = b[2];
                }
        }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* HBMA  (ME/hbma.py)                                                        */
/* ------------------------------------------------------------------------ */

/* One pyramid level: hbma.py:172-176.
 *   convolve(im/255.0, w[:,:,None], mode='constant')[::2, ::2] * 255.0 -> astype(uint8)
 * scipy.ndimage.convolve visits the 3x3 taps in raster order of the image
 * offsets (dy,dx) = (-1,-1) ... (+1,+1), accumulating in f64 from 0 (N5).
 * out is ceil(H/2) x ceil(W/2) x 3. */
ORC_API int orc_pyr_down(const uint8_t *im, int H, int W, uint8_t *out)
{
    static const double w[3][3] = {{0.0625, 0.125, 0.0625},
                                   {0.125, 0.25, 0.125},
                                   {0.0625, 0.125, 0.0625}};
    int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
#pragma omp parallel for schedule(static)
    for (int oy = 0; oy < Ho; ++oy) {
        int y = 2 * oy;
        for (int ox = 0; ox < Wo; ++ox) {
            int x = 2 * ox;
            for (int ch = 0; ch < 3; ++ch) {
                double acc = 0.0;
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        int yy = y + dy, xx = x + dx;
                        double v = 0.0;
                        if (yy >= 0 && yy < H && xx >= 0 && xx < W)
                            v = (double)im[((size_t)yy * W + xx) * 3 + ch] / 255.0;
                        acc += v * w[dy + 1][dx + 1];
                    }
                out[((size_t)oy * Wo + ox) * 3 + ch] = (uint8_t)(acc * 255.0);
            }
        }
    }
    return 0;
}

typedef struct { float r, c, cost; } vec3f;

/* hbma.py:34-75 block_wise_fs.  block1 points into padded frame 1 (pitch Wp px),
 * imp is padded frame 2.  (H, W) is the UNPADDED shape. */
static vec3f block_wise_fs(int cost_key, const uint8_t *block1, int Wp, const uint8_t *imp,
                           int idx_r, int idx_c, int win, int bs, int H, int W)
{
    vec3f res;
    if (idx_r >= H || idx_c >= W || idx_r < 0 || idx_c < 0) {          /* :36-37 */
        res.r = 0.f; res.c = 0.f; res.cost = INFINITY;
        return res;
    }
    double lowest_cost = INFINITY, lowest_dist = INFINITY;
    int lv_r = 0, lv_c = 0;
    int max_r_off = win + 1, min_r_off = -win, max_c_off = win + 1, min_c_off = -win;
    if (idx_r + win + bs >= H) max_r_off = imax(1, H - idx_r - bs);    /* :53-54 */
    if (idx_r - win < 0) min_r_off += win - idx_r;                     /* :55-56 */
    if (idx_c + win + bs >= W) max_c_off = imax(1, W - idx_c - bs);
    if (idx_c - win < 0) min_c_off += win - idx_c;
    for (int r_off = min_r_off; r_off < max_r_off; ++r_off) {
        int rr = idx_r + r_off;
        for (int c_off = min_c_off; c_off < max_c_off; ++c_off) {
            int cc = idx_c + c_off;
            const uint8_t *b2 = imp + ((size_t)rr * Wp + cc) * 3;
            double s = cost_key == 0 ? sad_f64(block1, Wp, b2, Wp, bs)
                                     : ssd_f64(block1, Wp, b2, Wp, bs);
            float cost = (float)s;                                      /* float32 return, :15 (N3) */
            double distance = (double)(r_off * r_off + c_off * c_off);
            if ((double)cost < lowest_cost ||
                ((double)cost == lowest_cost && distance < lowest_dist)) {   /* :69 */
                lowest_cost = (double)cost;
                lowest_dist = distance;
                lv_r = r_off; lv_c = c_off;
            }
        }
    }
    res.r = (float)lv_r; res.c = (float)lv_c; res.cost = (float)lowest_cost;
    return res;
}

/* hbma.py:77-98 full_search (+_jit): block-level output [ceil(H/bs)][ceil(W/bs)][3]. */
static float *hbma_full_search(int cost_key, int bs, int win, const uint8_t *im1,
                               const uint8_t *im2, int H, int W, int *rows, int *cols)
{
    uint8_t *p1 = pad_frame(im1, H, W, bs), *p2 = pad_frame(im2, H, W, bs);
    int Wp = W + bs;
    int nr = ceil_div(H, bs), nc = ceil_div(W, bs);
    float *mvs = (float *)calloc((size_t)nr * nc * 3, sizeof(float));
#pragma omp parallel for schedule(dynamic, 1)
    for (int row = 0; row < nr; ++row)
        for (int col = 0; col < nc; ++col) {
            int im_r = row * bs, im_c = col * bs;
            vec3f v = block_wise_fs(cost_key, p1 + ((size_t)im_r * Wp + im_c) * 3, Wp, p2,
                                    im_r, im_c, win, bs, H, W);
            float *o = mvs + ((size_t)row * nc + col) * 3;
            o[0] = v.r; o[1] = v.c; o[2] = v.cost;
        }
    free(p1); free(p2);
    *rows = nr; *cols = nc;
    return mvs;
}

/* hbma.py:100-149 increase_vec_density(+_jit).  mvs is [rows][cols][3] with row
 * pitch `pitch_cols` (the reference passes a sliced, non-contiguous view).
 * Returns a fresh [2*rows][2*cols][3] array. */
static float *hbma_increase(int cost_key, const float *mvs, int rows, int cols, int pitch_cols,
                            int bs, int sub_win, const uint8_t *im1, const uint8_t *im2,
                            int H, int W, int vec_scale)
{
    uint8_t *p1 = pad_frame(im1, H, W, bs), *p2 = pad_frame(im2, H, W, bs);
    int Wp = W + bs;
    int orows = rows << 1, ocols = cols << 1;
    float *out = (float *)calloc((size_t)orows * ocols * 3, sizeof(float));
#define MV(r, c) (mvs + ((size_t)(r) * pitch_cols + (c)) * 3)
#pragma omp parallel for schedule(dynamic, 1)
    for (int row = 0; row < rows; ++row)
        for (int col = 0; col < cols; ++col) {
            float vecs[3][2];
            vecs[0][0] = MV(row, col)[0];                  /* parent, NOT scaled  :107 (Q4) */
            vecs[0][1] = MV(row, col)[1];
            /* cols == 1 / rows == 1 is out-of-bounds (undefined) in the reference; we clamp. */
            const float *h = (col >= 1) ? MV(row, col - 1) : MV(row, imin(col + 1, cols - 1));   /* :108-111 */
            vecs[1][0] = h[0] * (float)vec_scale; vecs[1][1] = h[1] * (float)vec_scale;
            const float *v = (row >= 1) ? MV(row - 1, col) : MV(imin(row + 1, rows - 1), col);   /* :112-115 */
            vecs[2][0] = v[0] * (float)vec_scale; vecs[2][1] = v[1] * (float)vec_scale;
            int r_max = ((row + 1) * 2 * bs >= H) ? row * 2 + 1 : (row + 1) * 2;  /* :117-120 (Q5) */
            int c_max = ((col + 1) * 2 * bs >= W) ? col * 2 + 1 : (col + 1) * 2;  /* :121-124 */
            for (int o_r = row * 2; o_r < r_max; ++o_r) {
                int im_r = o_r * bs;
                for (int o_c = col * 2; o_c < c_max; ++o_c) {
                    int im_c = o_c * bs;
                    const uint8_t *block = p1 + ((size_t)im_r * Wp + im_c) * 3;
                    double lowest_cost = INFINITY;
                    float *o = out + ((size_t)o_r * ocols + o_c) * 3;
                    for (int k = 0; k < 3; ++k) {
                        /* (im_r + vec[0]) is float64, cast to int32 by the callee's
                         * signature -> truncation toward zero  :134 */
                        int ir = (int)((double)im_r + (double)vecs[k][0]);
                        int ic = (int)((double)im_c + (double)vecs[k][1]);
                        vec3f res = block_wise_fs(cost_key, block, Wp, p2, ir, ic, sub_win, bs, H, W);
                        if ((double)res.cost < lowest_cost) {                     /* :135 strict */
                            lowest_cost = (double)res.cost;
                            o[2] = (float)lowest_cost;
                            o[0] = res.r + vecs[k][0];
                            o[1] = res.c + vecs[k][1];
                        }
                    }
                }
            }
        }
#undef MV
    free(p1); free(p2);
    return out;
}

/* hbma.py:161-190 get_motion_vectors with upscale=False semantics: returns the
 * block-level field; *out_rows x *out_cols blocks of size *out_bs.  `out` must hold
 * ceil(H/min_bs')*ceil(W/min_bs')*3 floats where min_bs' is the final block size
 * (caller can size it with bs=1 worst case or call orc_hbma_shape first). */
ORC_API int orc_hbma_shape(int H, int W, int bs, int min_bs, int *rows, int *cols, int *fbs)
{
    while (bs > min_bs) bs >>= 1;
    if (bs <= 0) return -1;
    *rows = ceil_div(H, bs); *cols = ceil_div(W, bs); *fbs = bs;
    return 0;
}

ORC_API int orc_hbma_blocks(const uint8_t *im1, const uint8_t *im2, int H, int W, int bs,
                            int win, int sub_win, int steps, int min_bs, int cost_key,
                            float *out)
{
    if (steps < 0 || steps > 16 || bs <= 0) return -1;
    uint8_t *l1[17], *l2[17];
    int hs[17], ws[17];
    l1[0] = (uint8_t *)im1; l2[0] = (uint8_t *)im2; hs[0] = H; ws[0] = W;
    for (int i = 1; i <= steps; ++i) {                                  /* :171-177 */
        hs[i] = (hs[i - 1] + 1) / 2; ws[i] = (ws[i - 1] + 1) / 2;
        l1[i] = (uint8_t *)malloc((size_t)hs[i] * ws[i] * 3);
        l2[i] = (uint8_t *)malloc((size_t)hs[i] * ws[i] * 3);
        orc_pyr_down(l1[i - 1], hs[i - 1], ws[i - 1], l1[i]);
        orc_pyr_down(l2[i - 1], hs[i - 1], ws[i - 1], l2[i]);
    }
    int rows, cols;
    float *mvs = hbma_full_search(cost_key, bs, win, l1[steps], l2[steps], hs[steps], ws[steps],
                                  &rows, &cols);                       /* :179 */
    int pitch = cols;
    for (int lvl = steps - 1; lvl >= 0; --lvl) {                        /* :181-182 */
        float *nx = hbma_increase(cost_key, mvs, rows, cols, pitch, bs, sub_win, l1[lvl], l2[lvl],
                                  hs[lvl], ws[lvl], 2);
        free(mvs); mvs = nx;
        pitch = cols << 1;
        rows = ceil_div(hs[lvl], bs); cols = ceil_div(ws[lvl], bs);      /* slice */
    }
    while (bs > min_bs) {                                               /* :184-186 */
        bs >>= 1;
        if (bs == 0) { free(mvs); return -3; }
        float *nx = hbma_increase(cost_key, mvs, rows, cols, pitch, bs, sub_win, im1, im2, H, W, 1);
        free(mvs); mvs = nx;
        pitch = cols << 1;
        rows = ceil_div(H, bs); cols = ceil_div(W, bs);
    }
    for (int r = 0; r < rows; ++r)
        memcpy(out + (size_t)r * cols * 3, mvs + (size_t)r * pitch * 3, (size_t)cols * 3 * sizeof(float));
    free(mvs);
    for (int i = 1; i <= steps; ++i) { free(l1[i]); free(l2[i]); }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* MV smoothing  (smoothing/threeXthree_mv_smoothing.py:14-27)               */
/* ------------------------------------------------------------------------ */

static int cmp_float(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* mode: 1 = mean (mean_filter.py:2-4), 2 = median (median_filter.py:3-11),
 *       3 = weighted (weighted_mean_filter.py:4-15).  Dense f32 [H][W][3] in/out. */
ORC_API int orc_smooth_dense(const float *mv, int H, int W, int mode, int fs, float *out)
{
    static const float wts[9] = {1, 2, 1, 2, 6, 2, 1, 2, 1};
    if (fs <= 0) return -1;
    memcpy(out, mv, (size_t)H * W * 3 * sizeof(float));                /* np.copy  :17 */
    if (mode == 0) return 0;
    for (int row = 0; row < H; row += fs)
        for (int col = 0; col < W; col += fs) {
            int low_r = imax(0, row - fs), high_r = imin(row + fs, H) + 1;       /* :20-23 */
            int low_c = imax(0, col - fs), high_c = imin(col + fs, W) + 1;
            if (high_r > H) high_r = H;                                 /* slice clamps */
            if (high_c > W) high_c = W;
            float vr[9], vc[9];
            int n = 0;
            for (int r = low_r; r < high_r; r += fs)
                for (int c = low_c; c < high_c; c += fs) {
                    const float *p = mv + ((size_t)r * W + c) * 3;
                    vr[n] = p[0]; vc[n] = p[1]; ++n;
                }
            float o_r, o_c;
            if (mode == 1) {
                /* np.mean over f32: f32 accumulate (exact for integer MVs), one f32 divide */
                float sr = 0.f, sc = 0.f;
                for (int i = 0; i < n; ++i) { sr += vr[i]; sc += vc[i]; }
                o_r = sr / (float)n; o_c = sc / (float)n;
            } else if (mode == 2) {
                qsort(vr, n, sizeof(float), cmp_float);
                qsort(vc, n, sizeof(float), cmp_float);
                o_r = vr[n / 2]; o_c = vc[n / 2];                       /* upper median */
            } else {
                double tr = 0., tc = 0., tw = 0.;
                for (int i = 0; i < n; ++i) {
                    tr += (double)(vr[i] * wts[i]);                     /* f32 product, f64 sum */
                    tc += (double)(vc[i] * wts[i]);
                    tw += (double)wts[i];
                }
                o_r = (float)(int32_t)floor(tr / tw);                   /* `//` then int32 */
                o_c = (float)(int32_t)floor(tc / tw);
            }
            int r1 = imin(H, row + fs), c1 = imin(W, col + fs);
            for (int r = row; r < r1; ++r)
                for (int c = col; c < c1; ++c) {
                    float *p = out + ((size_t)r * W + c) * 3;
                    p[0] = o_r; p[1] = o_c;                             /* SAD channel untouched */
                }
        }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* MCI  (MCI/Unidirectional.py:36-84, MCI/Bidirectional.py:27-77)            */
/* ------------------------------------------------------------------------ */

/* Negative indices wrap (numba wraparound), SURVEY W1. */
static inline int wrap_idx(int i, int n) { return i < 0 ? i + n : i; }

static int cmp_u8(const void *a, const void *b)
{
    return (int)*(const uint8_t *)a - (int)*(const uint8_t *)b;
}

/* Median hole fill shared by both helpers (Unidirectional.py:62-80,
 * Bidirectional.py:55-73).  IF is the int32 frame with -1 holes (modified in place
 * exactly as the reference does), out is New_Interpolated_Frame. */
static void hole_fill(int32_t *IF, const uint8_t *tgt, int H, int W, uint8_t *out)
{
    const int k = 10;
    uint8_t *buf = (uint8_t *)malloc(21 * 21);
    for (int u = 0; u < H; ++u)
        for (int v = 0; v < W; ++v) {
            int32_t *p = IF + ((size_t)u * W + v) * 3;
            if (p[0] != -1) continue;
            const uint8_t *t = tgt + ((size_t)u * W + v) * 3;
            p[0] = t[0]; p[1] = t[1]; p[2] = t[2];
            int u0 = imax(0, u - k), u1 = imin(H, u + k + 1);
            int v0 = imax(0, v - k), v1 = imin(W, v + k + 1);
            for (int ch = 0; ch < 3; ++ch) {
                int n = 0;
                for (int a = u0; a < u1; ++a)
                    for (int b = v0; b < v1; ++b) {
                        int32_t x = IF[((size_t)a * W + b) * 3 + ch];
                        if (x != -1) buf[n++] = (uint8_t)x;              /* filter_out_neg1 */
                    }
                qsort(buf, n, 1, cmp_u8);
                /* np.median: odd -> middle; even -> (a+b)/2 in f64; uint8 store truncates */
                double med = (n & 1) ? (double)buf[n >> 1]
                                     : ((double)buf[(n >> 1) - 1] + (double)buf[n >> 1]) / 2.0;
                out[((size_t)u * W + v) * 3 + ch] = (uint8_t)med;
            }
        }
    free(buf);
}

/* Unidirectional.helper: dense MV field f32 [H][W][3], dist f32. */
ORC_API int orc_mci_unidir(const uint8_t *src, const uint8_t *tgt, const float *mv,
                           float dist, int H, int W, uint8_t *out)
{
    size_t npx = (size_t)H * W;
    int32_t *IF = (int32_t *)malloc(npx * 3 * sizeof(int32_t));
    float *SADI = (float *)malloc(npx * sizeof(float));
    if (!IF || !SADI) { free(IF); free(SADI); return -2; }
    for (size_t i = 0; i < npx * 3; ++i) IF[i] = -1;
    for (size_t i = 0; i < npx; ++i) SADI[i] = INFINITY;
    for (int u = 0; u < H; ++u)
        for (int v = 0; v < W; ++v) {
            const float *m = mv + ((size_t)u * W + v) * 3;
            const uint8_t *s = src + ((size_t)u * W + v) * 3;
            float pr = m[0] * dist, pc = m[1] * dist;                    /* f32 product (W1) */
            long long u_i = (long long)u + (long long)rintf(pr);         /* round half even  :48-49 */
            long long v_i = (long long)v + (long long)rintf(pc);
            if (u_i < H && v_i < W) {                                   /* :51 */
                /* below -H / -W numba's unchecked wraparound writes outside the array: undefined in the reference */
                if (u_i < -(long long)H || v_i < -(long long)W) { free(IF); free(SADI); return -3; }
                size_t q = (size_t)wrap_idx((int)u_i, H) * W + wrap_idx((int)v_i, W);
                if (m[2] <= SADI[q]) {                                  /* :52 */
                    IF[q * 3 + 0] = s[0]; IF[q * 3 + 1] = s[1]; IF[q * 3 + 2] = s[2];
                    SADI[q] = m[2];
                }
            } else {                                                    /* :56-58 */
                size_t q = (size_t)u * W + v;
                IF[q * 3 + 0] = s[0]; IF[q * 3 + 1] = s[1]; IF[q * 3 + 2] = s[2];
                SADI[q] = m[2];
            }
        }
    for (size_t i = 0; i < npx * 3; ++i) out[i] = (uint8_t)IF[i];       /* :65 (holes -> 255) */
    hole_fill(IF, tgt, H, W, out);
    free(IF); free(SADI);
    return 0;
}

/* Bidirectional.helper. */
ORC_API int orc_mci_bidir(const uint8_t *src, const uint8_t *tgt, const float *fwd,
                          const float *bwd, float dist, int H, int W, uint8_t *out)
{
    size_t npx = (size_t)H * W;
    int32_t *IF = (int32_t *)malloc(npx * 3 * sizeof(int32_t));
    float *SADI = (float *)malloc(npx * sizeof(float));
    if (!IF || !SADI) { free(IF); free(SADI); return -2; }
    for (size_t i = 0; i < npx * 3; ++i) IF[i] = -1;
    for (size_t i = 0; i < npx; ++i) SADI[i] = INFINITY;
    double bwdist = 1.0 - (double)dist;                                 /* f64  :42 (W2) */
    for (int u = 0; u < H; ++u)
        for (int v = 0; v < W; ++v) {
            size_t p = (size_t)u * W + v;
            const float *f = fwd + p * 3, *b = bwd + p * 3;
            const uint8_t *s = src + p * 3, *t = tgt + p * 3;
            {
                float pr = f[0] * dist, pc = f[1] * dist;
                long long u_i = (long long)u + (long long)rintf(pr);    /* :34-35 */
                long long v_i = (long long)v + (long long)rintf(pc);
                if (u_i < H && v_i < W) {
                    if (u_i < -(long long)H || v_i < -(long long)W) { free(IF); free(SADI); return -3; }   /* undefined in the reference */
                    size_t q = (size_t)wrap_idx((int)u_i, H) * W + wrap_idx((int)v_i, W);
                    if (f[2] < SADI[q]) {                               /* :37 strict */
                        IF[q * 3 + 0] = s[0]; IF[q * 3 + 1] = s[1]; IF[q * 3 + 2] = s[2];
                        SADI[q] = f[2];
                    }
                }
            }
            {
                double pr = (double)b[0] * bwdist, pc = (double)b[1] * bwdist;
                long long u_i = (long long)u + (long long)rint(pr);     /* :43-44 */
                long long v_i = (long long)v + (long long)rint(pc);
                if (u_i < H && v_i < W) {
                    if (u_i < -(long long)H || v_i < -(long long)W) { free(IF); free(SADI); return -3; }   /* undefined in the reference */
                    size_t q = (size_t)wrap_idx((int)u_i, H) * W + wrap_idx((int)v_i, W);
                    float bs = b[2], si = SADI[q];
                    if (bs < si) {                                      /* :46 */
                        if (si != INFINITY) {                           /* :47-50 blend (W3) */
                            float tt = bs + si;
                            for (int ch = 0; ch < 3; ++ch) {
                                /* uint8*f32 -> f32, /f32 -> f32; int32*f32 -> f64, /f32 -> f64 */
                                float a = ((float)t[ch] * si) / tt;
                                double c = ((double)IF[q * 3 + ch] * (double)bs) / (double)tt;
                                /* the fused array expression is typed float32 by numba:
                                 * the f64 sum is rounded to f32 before the int32 store */
                                float sum = (float)((double)a + c);
                                IF[q * 3 + ch] = (int32_t)sum;
                            }
                            SADI[q] = bs;
                        } else {                                        /* :51-53 */
                            IF[q * 3 + 0] = t[0]; IF[q * 3 + 1] = t[1]; IF[q * 3 + 2] = t[2];
                            SADI[q] = bs;
                        }
                    }
                }
            }
        }
    for (size_t i = 0; i < npx * 3; ++i) out[i] = (uint8_t)IF[i];       /* :57 */
    hole_fill(IF, tgt, H, W, out);
    free(IF); free(SADI);
    return 0;
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ======================================================================== */
/* bidir2: MCI/Bidirectional_2.py (Wang et al., IEWMC + error-adaptive       */
/* combination + block-wise directional hole interpolation)                  */
/* ======================================================================== */

/* IEWMC_helper, MCI/Bidirectional_2.py:49-127.  `mv` is the field exactly as the reference
 * indexes it (:63-67): block level [ceil(H/mbs)][ceil(W/mbs)][3] when upscale == 0 (hbma), dense
 * [H][W][3] sampled at the block origin otherwise.  `w` = get_weight_kern(2*mbs)[:,:,0] as float32
 * (:33-47; the three channels carry the same weights).  Outputs F, E: float32 [H+2p][W+2p][3].
 *
 * numba's typing, reproduced term by term (probed against the reference):
 *   TXb = int(dist * Xb)            float32 * int64 -> float64, truncated                 (:71-72)
 *   SAD = MV_i[2] / SAD_norm        float32 / int64 -> float64, compared with 250         (:73,:87)
 *   t_a = (1 - dist) * t_1          int64 - float32 -> float64 scalar, * uint8 -> float64 (:89)
 *   t_b = dist * t_2                float32 * uint8 array -> float32                      (:92)
 *   acc += w * (t_a + t_b)          float64, rounded to float32 on the store into acc     (:94-96)
 *   acc += w * F1 (occluded)        float32 * uint8 -> float32                            (:98)
 *   WM  = (w * |F1 - F2|).astype(f32)   float32 * int32 -> float64, then rounded          (:105)
 * and the three accumulators are float32 sums in block-raster order (:100,:106,:108).
 * Returns 0, or -1 if a shifted block leaves the padded arrays (the reference's slices would clip and
 * its in-place adds would fail to broadcast). */
ORC_API int orc_iewmc(const uint8_t *F1, const uint8_t *F2, const float *mv, int upscale, int H, int W,
                      int mbs, int pad, float dist, const float *w, float *Fout, float *Eout)
{
    const int Hp = H + 2 * pad, Wp = W + 2 * pad, eb = 2 * mbs, half = mbs / 2;
    const size_t npx = (size_t)Hp * Wp;
    float *accum = (float *)calloc(npx * 3, sizeof(float));
    float *wmcd = (float *)calloc(npx * 3, sizeof(float));
    float *wsum = (float *)calloc(npx * 3, sizeof(float));
    uint8_t *P1 = (uint8_t *)calloc(npx * 3, 1), *P2 = (uint8_t *)calloc(npx * 3, 1);
    if (!accum || !wmcd || !wsum || !P1 || !P2) { free(accum); free(wmcd); free(wsum); free(P1); free(P2); return -2; }
    for (int r = 0; r < H; ++r) {
        memcpy(P1 + ((size_t)(r + pad) * Wp + pad) * 3, F1 + (size_t)r * W * 3, (size_t)W * 3);
        memcpy(P2 + ((size_t)(r + pad) * Wp + pad) * 3, F2 + (size_t)r * W * 3, (size_t)W * 3);
    }
    const int mv_cols = upscale ? W : ceil_div(W, mbs);
    const double one_minus = 1.0 - (double)dist;              /* (1-dist): int64 - float32 -> float64 */
    const long long sad_norm = (long long)mbs * mbs * 3;
    int rc = 0;
    for (int br = 0; br < H && rc == 0; br += mbs) {
        for (int bc = 0; bc < W; bc += mbs) {
            const float *m = upscale ? mv + ((size_t)br * mv_cols + bc) * 3 : mv + ((size_t)(br / mbs) * mv_cols + bc / mbs) * 3;
            const long long Xb = (long long)m[0], Yb = (long long)m[1];
            const long long TXb = (long long)((double)dist * (double)Xb), TYb = (long long)((double)dist * (double)Yb);
            const double sad = (double)m[2] / (double)sad_norm;
            const int occluded = !(sad < 250.0);
            const long long r0 = br + pad - half, c0 = bc + pad - half;
            if (r0 + Xb < 0 || r0 + Xb + eb > Hp || c0 + Yb < 0 || c0 + Yb + eb > Wp ||
                r0 + TXb < 0 || r0 + TXb + eb > Hp || c0 + TYb < 0 || c0 + TYb + eb > Wp) { rc = -1; break; }
            for (int i = 0; i < eb; ++i) {
                for (int j = 0; j < eb; ++j) {
                    const float wk = w[i * eb + j];
                    const size_t s1 = ((size_t)(r0 + i) * Wp + (c0 + j)) * 3;
                    const size_t s2 = ((size_t)(r0 + Xb + i) * Wp + (c0 + Yb + j)) * 3;
                    const size_t t = ((size_t)(r0 + TXb + i) * Wp + (c0 + TYb + j)) * 3;
                    for (int ch = 0; ch < 3; ++ch) {
                        const uint8_t a = P1[s1 + ch], b = P2[s2 + ch];
                        float av;
                        if (!occluded) {
                            const double ta = one_minus * (double)a;
                            const float tb = dist * (float)b;
                            const double tc = ta + (double)tb;
                            av = (float)((double)wk * tc);
                        } else {
                            av = wk * (float)a;
                        }
                        accum[t + ch] = accum[t + ch] + av;
                        const int d = (int)a - (int)b;
                        const float wm = (float)((double)wk * (double)(d < 0 ? -d : d));
                        wmcd[t + ch] = wmcd[t + ch] + wm;
                        wsum[t + ch] = wsum[t + ch] + wk;
                    }
                }
            }
        }
    }
    if (rc == 0) {
        for (size_t p = 0; p < npx; ++p) {                    /* normalisation, :113-122 */
            const float ws = wsum[p * 3];
            for (int ch = 0; ch < 3; ++ch) {
                if (ws <= 0.f) { Fout[p * 3 + ch] = -1.f; Eout[p * 3 + ch] = -1.f; }
                else { Fout[p * 3 + ch] = accum[p * 3 + ch] / ws; Eout[p * 3 + ch] = wmcd[p * 3 + ch] / ws; }
            }
        }
    }
    free(accum); free(wmcd); free(wsum); free(P1); free(P2);
    return rc;
}

/* Error_adaptive_combination, MCI/Bidirectional_2.py:147-166.  alpha = 1 is an int64 literal, so
 * (E + alpha) and the whole quotient are float64 (float32 array + int64 -> float64); Ef + Eb is a
 * float32 sum first.  The result is rounded to float32 on the store into Fc. */
ORC_API void orc_eac(const float *Ff, const float *Ef, const float *Fb, const float *Eb, size_t npx, float *Fc)
{
    for (size_t p = 0; p < npx; ++p) {
        const int hf = Ff[p * 3] != -1.f, hb = Fb[p * 3] != -1.f;
        for (int ch = 0; ch < 3; ++ch) {
            const size_t k = p * 3 + ch;
            if (hf && hb) {
                const double num = ((double)Eb[k] + 1.0) * (double)Ff[k] + ((double)Ef[k] + 1.0) * (double)Fb[k];
                const float es = Ef[k] + Eb[k];
                Fc[k] = (float)(num / ((double)es + 2.0));
            } else if (hf) Fc[k] = Ff[k];
            else if (hb) Fc[k] = Fb[k];
            else Fc[k] = -1.f;
        }
    }
}

/* fill_holes, MCI/Bidirectional_2.py:168-262, on one n x n block (row pitch `pitch` floats of 3 channels).
 *   - ah/av/ad/ar: float32 running sums started at 0.1 in (i, j) raster order over i, j < n-1; the
 *     counter of the anti-diagonal direction is never incremented (:193-194), so ar is divided by 0.1;
 *   - p_array = np.full((4,3), -1) is an INT64 array: every directional estimate is truncated toward
 *     zero when it is stored (:236-244); distances are float64 hypot;
 *   - `A += 1 / a` and `pixel_interp += p / a` are in-place adds on float32 arrays: numba rounds the
 *     right-hand side to float32 first and adds in float32 (probed: adding in float64 is 1 ulp off in
 *     ~95 % of the blocks). */
static void fill_holes_block(const float *blk, int pitch, int n, float *out, int opitch)
{
    float a[4][4];
    for (int d = 0; d < 4; ++d) for (int k = 0; k < 4; ++k) a[d][k] = 0.1f;
    for (int i = 0; i < n - 1; ++i) {
        for (int j = 0; j < n - 1; ++j) {
            const float *pi = blk + ((size_t)i * pitch + j) * 3;
            if (pi[0] == -1.f) continue;
            const float *q;
            q = blk + ((size_t)(i + 1) * pitch + j) * 3;
            if (q[0] != -1.f) { for (int k = 0; k < 3; ++k) a[0][k] = a[0][k] + fabsf(pi[k] - q[k]); a[0][3] = a[0][3] + 1.f; }
            q = blk + ((size_t)i * pitch + j + 1) * 3;
            if (q[0] != -1.f) { for (int k = 0; k < 3; ++k) a[1][k] = a[1][k] + fabsf(pi[k] - q[k]); a[1][3] = a[1][3] + 1.f; }
            q = blk + ((size_t)(i + 1) * pitch + j + 1) * 3;
            if (q[0] != -1.f) { for (int k = 0; k < 3; ++k) a[2][k] = a[2][k] + fabsf(pi[k] - q[k]); a[2][3] = a[2][3] + 1.f; }
            if (i >= 1) {
                q = blk + ((size_t)(i - 1) * pitch + j + 1) * 3;
                if (q[0] != -1.f) { for (int k = 0; k < 3; ++k) a[3][k] = a[3][k] + fabsf(pi[k] - q[k]); }
            }
        }
    }
    float aa[4][3];
    for (int d = 0; d < 4; ++d) for (int k = 0; k < 3; ++k) aa[d][k] = a[d][k] / a[d][3];
    static const int dir[4][2] = {{1, 0}, {0, 1}, {1, 1}, {-1, 1}};
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            float *o = out + ((size_t)i * opitch + j) * 3;
            const float *self = blk + ((size_t)i * pitch + j) * 3;
            if (self[0] != -1.f) { o[0] = self[0]; o[1] = self[1]; o[2] = self[2]; continue; }
            long long parr[4][3];
            for (int d = 0; d < 4; ++d) {
                parr[d][0] = parr[d][1] = parr[d][2] = -1;
                const float *p1 = NULL, *p2 = NULL;
                double d1 = 0.0, d2 = 0.0;
                int n1 = i, m1 = j, n2 = i, m2 = j;
                while (!p1 && n1 >= 0 && n1 < n && m1 >= 0 && m1 < n) {
                    const float *q = blk + ((size_t)n1 * pitch + m1) * 3;
                    if (q[0] != -1.f) { p1 = q; d1 = hypot((double)(i - n1), (double)(j - m1)); }
                    n1 -= dir[d][0]; m1 -= dir[d][1];
                }
                while (!p2 && n2 >= 0 && n2 < n && m2 >= 0 && m2 < n) {
                    const float *q = blk + ((size_t)n2 * pitch + m2) * 3;
                    if (q[0] != -1.f) { p2 = q; d2 = hypot((double)(i - n2), (double)(j - m2)); }
                    n2 += dir[d][0]; m2 += dir[d][1];
                }
                for (int k = 0; k < 3; ++k) {
                    if (p1 && p2) parr[d][k] = (long long)(((d2 * (double)p1[k]) + (d1 * (double)p2[k])) / (d1 + d2));
                    else if (p1) parr[d][k] = (long long)p1[k];
                    else if (p2) parr[d][k] = (long long)p2[k];
                }
            }
            float A[3] = {0.f, 0.f, 0.f}, pix[3] = {0.f, 0.f, 0.f};
            for (int d = 0; d < 4; ++d) {
                if (parr[d][0] == -1) continue;
                for (int k = 0; k < 3; ++k) {
                    A[k] = A[k] + 1.f / aa[d][k];
                    pix[k] = pix[k] + (float)((double)parr[d][k] / (double)aa[d][k]);
                }
            }
            for (int k = 0; k < 3; ++k) o[k] = pix[k] / A[k];
        }
    }
}

/* BDHI + unpad_and_downconvert, MCI/Bidirectional_2.py:265-293: 2*mbs blocks from (pad, pad); a block
 * with a hole (channel 0 == -1) is replaced by fill_holes(block); then [pad:-pad] as uint8. */
ORC_API void orc_bdhi_unpad(const float *Fc, int H, int W, int mbs, int pad, uint8_t *out)
{
    const int Hp = H + 2 * pad, Wp = W + 2 * pad, fb = 2 * mbs;
    float *Fo = (float *)malloc((size_t)Hp * Wp * 3 * sizeof(float));
    memcpy(Fo, Fc, (size_t)Hp * Wp * 3 * sizeof(float));
    for (int r = pad; r < Hp - pad; r += fb) {
        for (int c = pad; c < Wp - pad; c += fb) {
            const int nr = imin(fb, Hp - r), nc = imin(fb, Wp - c);
            int hole = 0;
            for (int i = 0; i < nr && !hole; ++i)
                for (int j = 0; j < nc; ++j)
                    if (Fc[((size_t)(r + i) * Wp + c + j) * 3] == -1.f) { hole = 1; break; }
            if (hole && nr == fb && nc == fb)
                fill_holes_block(Fc + ((size_t)r * Wp + c) * 3, Wp, fb, Fo + ((size_t)r * Wp + c) * 3, Wp);
        }
    }
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c)
            for (int k = 0; k < 3; ++k) {
                const float v = Fo[((size_t)(r + pad) * Wp + c + pad) * 3 + k];
                out[((size_t)r * W + c) * 3 + k] = (uint8_t)(int)v;      /* float32 -> uint8 as numba lowers it */
            }
    free(Fo);
}

/* fill_holes on one contiguous n x n x 3 block (test hook for the validation script) */
ORC_API void orc_fill_holes(const float *blk, int n, float *out) { fill_holes_block(blk, n, n, out, n); }
