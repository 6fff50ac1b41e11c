/*
 * memci_b200.h -- C ABI of the B200-native MEMCI hot path (libmemci_b200.so).
 *
 * Drop-in boundary for lhl2617/Interpolatory's MEMCI interpolators.  Every entry
 * point names the reference interface it replaces; paths are relative to
 *   Simulator/python/src/Interpolators/MEMCI/   of the reference.
 *
 * Conventions
 *   - plain C: opaque context pointer, raw pointers and sizes, no torch / C++ types;
 *   - every function returns 0 on success or a negative memci_status; the message of
 *     the last failure is available from memci_last_error();
 *   - frames are C-contiguous uint8 [H][W][3] RGB, exactly what
 *     VideoStream.get_frame() delivers in the reference (src/VideoStream.py:61);
 *   - the caller owns all host buffers; the library owns the device "slots" inside
 *     a context.  Nothing allocated by the library is ever returned to the caller;
 *   - all work is enqueued on the context's CUDA stream; host<->device copies are
 *     asynchronous when the host buffer is pinned.  memci_sync() waits for the stream.
 *     A context is not thread-safe; contexts are independent (one per GPU / stream).
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     MEMCI_ERR_CUDA.
 */
#ifndef MEMCI_B200_H
#define MEMCI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MEMCI_API __attribute__((visibility("default")))
#else
#define MEMCI_API
#endif

typedef struct memci_ctx memci_ctx;

typedef enum memci_status {
    MEMCI_OK = 0,
    MEMCI_ERR_ARG = -1,      /* bad argument (slot index, sizes, modes)          */
    MEMCI_ERR_CUDA = -2,     /* CUDA runtime / driver error (sticky per context)  */
    MEMCI_ERR_STATE = -3,    /* slot empty / shape mismatch                       */
    MEMCI_ERR_NOMEM = -4
} memci_status;

/* smoothing modes: smoothing/threeXthree_mv_smoothing.py:14 `filter_func`,
 * MCI/MEMCIBaseInterpolator.py:23-29 smoothing_dict */
enum { MEMCI_FILTER_NONE = 0, MEMCI_FILTER_MEAN = 1, MEMCI_FILTER_MEDIAN = 2, MEMCI_FILTER_WEIGHTED = 3 };
/* ME/hbma.py:10-13 cost_key_map */
enum { MEMCI_COST_SAD = 0, MEMCI_COST_SSD = 1 };

#define MEMCI_NUM_FRAME_SLOTS 8
#define MEMCI_NUM_MV_SLOTS 8

/* ---- library / context ------------------------------------------------------ */

MEMCI_API int memci_version(void);
/* Number of CUDA devices visible; <0 on error (no driver / no device). */
MEMCI_API int memci_device_count(void);

/* One context = one GPU + one stream + device slots for frames of height H, width W.
 * `stream` is an existing cudaStream_t (as void*) to enqueue on, or NULL to let the
 * library create its own non-blocking stream.
 * Replaces: the implicit per-interpolator state of MEMCIBaseInterpolator
 * (MCI/MEMCIBaseInterpolator.py:30-44: MV_field_idx, fwr/bwr_MV_field). */
MEMCI_API int memci_ctx_create(int device, int H, int W, void *stream, memci_ctx **out);
MEMCI_API int memci_ctx_destroy(memci_ctx *ctx);
/* ctx may be NULL: returns the last error of a failed memci_ctx_create. */
MEMCI_API const char *memci_last_error(memci_ctx *ctx);
MEMCI_API int memci_sync(memci_ctx *ctx);
/* The cudaStream_t the context enqueues on (for CUDA-event timing by the caller). */
MEMCI_API void *memci_stream(memci_ctx *ctx);

/* Page-locked host staging buffers (the decode-to-device staging seam:
 * src/Interpolators/base.py:46-64 load_input_video / src/VideoStream.py:61 get_frame). */
MEMCI_API int memci_host_alloc(size_t bytes, void **out);
MEMCI_API int memci_host_free(void *p);

/* ---- frames ----------------------------------------------------------------- */

/* Host uint8[H][W][3] -> frame slot (async H2D on the context stream). */
MEMCI_API int memci_upload_frame(memci_ctx *ctx, int slot, const uint8_t *host_hwc);
/* Device uint8[H][W][3] -> frame slot (D2D; for inputs already resident in HBM). */
MEMCI_API int memci_set_frame_device(memci_ctx *ctx, int slot, const void *dev_hwc);
/* Frame slot -> host uint8[H][W][3] (async D2H; memci_sync() before reading). */
MEMCI_API int memci_download_frame(memci_ctx *ctx, int slot, uint8_t *host_hwc);
/* Device pointer of a frame slot's uint8[H][W][3] buffer (NULL on bad slot). */
MEMCI_API void *memci_frame_device_ptr(memci_ctx *ctx, int slot);

/* ---- frame staging (decode-to-device path) -------------------------------------------------
 * Replaces the host-side hand-over of decoded frames, VideoStream.get_frame -> interpolator
 * (src/VideoStream.py:61, base.py:103), for streaming use: every source frame crosses PCIe ONCE on the
 * stager's own H2D stream and is handed to any context with a device-to-device copy that waits on the
 * upload's event; output frames drain through the stager's D2H stream while the context's stream goes on
 * with the next pair.  Host buffers must be page-locked (memci_host_alloc) for the copies to overlap. */
typedef struct memci_stager memci_stager;
MEMCI_API int memci_stager_create(int device, int H, int W, int n_slots, memci_stager **out);
MEMCI_API int memci_stager_destroy(memci_stager *st);
/* Async H2D of one uint8[H][W][3] frame into ring slot `slot` (waits until every context that took the
 * slot's previous frame has copied it). */
MEMCI_API int memci_stager_upload(memci_stager *st, int slot, const uint8_t *host_hwc);
/* Ring slot -> frame slot of `ctx` (D2D on the context's stream, after the upload; at most 4 takers per upload). */
MEMCI_API int memci_frame_from_stager(memci_ctx *ctx, int frame_slot, memci_stager *st, int slot);
/* Frame slot -> host through the stager's D2H stream; the slot is rewritten only after the copy (the MCI
 * entry points wait for it).  memci_stager_sync() before reading the host buffer. */
MEMCI_API int memci_download_frame_via(memci_ctx *ctx, int slot, uint8_t *host_hwc, memci_stager *st);
MEMCI_API int memci_stager_sync(memci_stager *st);

/* ---- motion estimation ------------------------------------------------------ */

/* Replaces ME/full_search.py:61 get_motion_vectors(block_size, target_region, im1, im2)
 * (helper :19-55, get_sad :10-17).  Result (block level) goes to MV slot `mv_slot`. */
MEMCI_API int memci_me_fs(memci_ctx *ctx, int slot_src, int slot_tgt, int block_size,
                          int target_region, int mv_slot);

/* Both directions of one frame pair in ONE persistent launch (the two get_motion_vectors calls of
 * Bidirectional.py:138-143: fwd = fs(bs, R, a, b), bwd = fs(bs, R, b, a)); same results as two
 * memci_me_fs calls, one kernel start-up / tail and one exact-redo pass instead of two. */
MEMCI_API int memci_me_fs_pair(memci_ctx *ctx, int slot_a, int slot_b, int block_size, int target_region,
                               int mv_fwd, int mv_bwd);

/* Replaces ME/tss.py:63 get_motion_vectors(block_size, steps, im1, im2) (helper :18-61). */
MEMCI_API int memci_me_tss(memci_ctx *ctx, int slot_src, int slot_tgt, int block_size, int steps, int mv_slot);

/* Replaces ME/hbma.py:161 get_motion_vectors(block_size, win_size, sub_win_size, steps,
 * min_block_size, im1, im2, cost_str) (pyramid :171-177, full_search :77-98,
 * increase_vec_density :100-149).  Result at the final block size goes to `mv_slot`. */
MEMCI_API int memci_me_hbma(memci_ctx *ctx, int slot_src, int slot_tgt, int block_size,
                            int win_size, int sub_win_size, int steps, int min_block_size,
                            int cost_key, int mv_slot);

/* One pyramid level of ME/hbma.py:172-176 on frame slot `slot`, level >= 1, copied to
 * host uint8[ceil(H/2^l)][ceil(W/2^l)][3] (test / inspection hook). */
MEMCI_API int memci_download_pyramid_level(memci_ctx *ctx, int slot, int level, uint8_t *host);

/* Replaces smoothing/threeXthree_mv_smoothing.py:14 smooth(filter_func, mv_field, block_size)
 * with mean_filter.py:2, median_filter.py:3-15, weighted_mean_filter.py:4-19.
 * mv_in and mv_out may be the same slot. */
MEMCI_API int memci_smooth(memci_ctx *ctx, int mv_in, int filter_mode, int filter_size, int mv_out);

/* Geometry of an MV slot: the (dy,dx) grid and the SAD grid (they differ after
 * smoothing: vectors live on filter_size cells, SAD stays on ME blocks). */
MEMCI_API int memci_mv_info(memci_ctx *ctx, int mv_slot, int *mv_cell, int *mv_rows, int *mv_cols,
                            int *sad_cell, int *sad_rows, int *sad_cols);
/* Block-level download: vec[mv_rows][mv_cols][2] (dy,dx) and sad[sad_rows][sad_cols].
 * Either pointer may be NULL. */
MEMCI_API int memci_download_mv_blocks(memci_ctx *ctx, int mv_slot, float *vec, float *sad);
/* Dense float32[H][W][3] (dy,dx,sad): the array the reference's ME functions return
 * (full_search.py:51-53, hbma.py:151-159 upscale_mvs). */
MEMCI_API int memci_download_mv_dense(memci_ctx *ctx, int mv_slot, float *host_hw3);
/* Arbitrary dense float32[H][W][3] field -> MV slot (for callers of the reference's
 * function-level helper(source, target, MV_field, dist)). */
MEMCI_API int memci_upload_mv_dense(memci_ctx *ctx, int mv_slot, const float *host_hw3);
/* Block-level upload: vec[rows][cols][2] on `cell`-sized cells and sad on its own grid. */
MEMCI_API int memci_upload_mv_blocks(memci_ctx *ctx, int mv_slot, int mv_cell, const float *vec,
                                     int sad_cell, const float *sad);

/* ---- motion-compensated interpolation --------------------------------------- */

/* Replaces MCI/Unidirectional.py:36 helper(source_frame, target_frame, MV_field, dist). */
MEMCI_API int memci_mci_unidir(memci_ctx *ctx, int slot_src, int slot_tgt, int mv_slot,
                               float dist, int out_slot);
/* Replaces MCI/Bidirectional.py:27 helper(source, target, fwr_MV_field, bwr_MV_field, dist). */
MEMCI_API int memci_mci_bidir(memci_ctx *ctx, int slot_src, int slot_tgt, int mv_fwd,
                              int mv_bwd, float dist, int out_slot);

/* Replaces the MCI half of BiDir2Interpolator.get_interpolated_frame (MCI/Bidirectional_2.py:330-345):
 *   Ff, Ef = IEWMC(fwr_MV_field, source, target, dist, pad_size, min_block_size, upscale_MV)      (:49-144)
 *   Fb, Eb = IEWMC(bwr_MV_field, target, source, 1 - dist, ...)
 *   Out    = unpad_and_downconvert(BDHI(Error_adaptive_combination(Ff, Ef, Fb, Eb)))              (:147-290)
 * dist_fwd / dist_bwd are the two float32 values the reference's njit boundary sees (dist and 1 - dist,
 * each rounded from the Python float); weights = get_weight_kern(2 * min_block_size)[:, :, 0] (:33-47),
 * 64 float32, row-major, copied before the call returns.  The MV slots may be block level (hbma,
 * upscale=False) or coarser ME blocks (fs / tss: the dense field sampled at the 4x4 block origins). */
MEMCI_API int memci_mci_bidir2(memci_ctx *ctx, int slot_src, int slot_tgt, int mv_fwd, int mv_bwd, float dist_fwd,
                               float dist_bwd, int min_block_size, int pad_size, const float *weights, int out_slot);
/* After memci_mci_bidir2: largest |int(dist * mv)| and the number of blocks whose shifted slices leave
 * the padded arrays (the reference's numpy slices would clip and its in-place adds fail; we zero-fill). */
MEMCI_API int memci_bidir2_stats(memci_ctx *ctx, int *max_shift, int *n_out_of_range);
/* Blocks shifted outside the padded frame summed over EVERY memci_mci_bidir2 call since the last reset (synchronises
 * the stream): the throughput path (pipeline.PairPipeline) enqueues many frames before it looks, and must still raise
 * where the reference's numpy slicing fails (MCI/Bidirectional_2.py:98-106 with vectors beyond pad_size). */
MEMCI_API int memci_bidir2_out_of_range_total(memci_ctx *ctx, int *n_out_of_range, int reset);

/* ---- frame quality (Middlebury harness) --------------------------------------------------- */
/* Replaces the two skimage calls of the reference's benchmark harness (src/Benchmark.py:46-47 in benchmark(),
 * :105-106 in get_middle_frame()) for two uint8 RGB frames resident in slots:
 *   *mse  = skimage.metrics.mean_squared_error(a, b)   (exact: integer sum / (3 H W));
 *           peak_signal_noise_ratio(a, b, data_range=255) = 10 log10(255^2 / mse) is left to the caller
 *   *ssim = skimage.metrics.structural_similarity(a, b, data_range=255, multichannel=True): win_size 7, uniform
 *           window, sample covariance, K1 0.01, K2 0.03, mean over the frame cropped by 3 px and the 3 channels.
 * Either pointer may be NULL.  Synchronises the context's stream.  Frames smaller than 7 x 7: MEMCI_ERR_ARG
 * (skimage raises ValueError: win_size exceeds image extent). */
MEMCI_API int memci_quality(memci_ctx *ctx, int slot_a, int slot_b, double *mse, double *ssim);

/* ---- spatial tiling of one frame pair into horizontal bands (8K mode) -------------
 * A band = block rows [block_row0, block_row1).  Each GPU holds full-size slots but only fills the
 * rows it owns plus the halo rows it received; these entry points restrict the work to a band. */

/* Full search for block rows [block_row0, block_row1) only; other rows of the MV slot are left
 * untouched (they are filled by memci_mv_device_ptrs + a peer copy). */
MEMCI_API int memci_me_fs_rows(memci_ctx *ctx, int slot_src, int slot_tgt, int block_size, int target_region,
                               int mv_slot, int block_row0, int block_row1);
/* MCI (bidir != 0: Bidirectional.helper, else Unidirectional.helper) for output rows [row0, row1):
 * the splat is evaluated for the rows and 10 rows of hole-fill halo either side; only holes in
 * [row0, row1) are filled.  mv_bwd is ignored when bidir == 0. */
MEMCI_API int memci_mci_rows(memci_ctx *ctx, int slot_src, int slot_tgt, int mv_fwd, int mv_bwd, float dist,
                             int out_slot, int bidir, int row0, int row1);
/* Rows [row0, row1) of a frame slot <-> the same rows of a full-size host frame. */
MEMCI_API int memci_upload_frame_rows(memci_ctx *ctx, int slot, const uint8_t *host_hwc, int row0, int row1);
MEMCI_API int memci_download_frame_rows(memci_ctx *ctx, int slot, uint8_t *host_hwc, int row0, int row1);
/* Declare a frame slot valid / modified after its device buffer (memci_frame_device_ptr) was written
 * from outside (peer copy, NCCL recv): cached luma planes and pyramids are dropped. */
MEMCI_API int memci_frame_mark(memci_ctx *ctx, int slot);
/* Allocate a full-size MV slot on `cell`-sized blocks without computing it, and expose the device
 * arrays (float2 vec[rows][cols], float sad[rows][cols]) for halo exchange. */
MEMCI_API int memci_mv_alloc(memci_ctx *ctx, int mv_slot, int cell);
MEMCI_API int memci_mv_device_ptrs(memci_ctx *ctx, int mv_slot, void **vec, void **sad);

/* ---- fused convenience (one call per output frame of the class API) --------- */

typedef struct memci_pair_params {
    int me_mode;        /* 0 = fs, 1 = hbma, 2 = tss (region = steps)          */
    int block_size;     /* MEMCIBaseInterpolator.py:34                        */
    int region;         /* target_region / win_size  :35                      */
    int sub_region;     /* :40                                                */
    int steps;          /* hbma pyramid depth (caller applies the auto rule)  */
    int min_block_size; /* :42                                                */
    int filter_mode;    /* MEMCI_FILTER_*                                     */
    int filter_size;    /* :39                                                */
    int bidir;          /* 0 = unidir, 1 = bidir                              */
} memci_pair_params;

/* ME (+smoothing) for the pair in frame slots (slot_src, slot_tgt) into MV slots
 * mv_fwd (and mv_bwd when p->bidir).  Mirrors the cache-miss branch of
 * get_interpolated_frame (Unidirectional.py:109-125, Bidirectional.py:138-155). */
MEMCI_API int memci_estimate_pair(memci_ctx *ctx, const memci_pair_params *p, int slot_src,
                                  int slot_tgt, int mv_fwd, int mv_bwd);

/* ---- instrumentation ---------------------------------------------------------- */

/* Kernels launched by this context since creation (or since the last reset). */
MEMCI_API long long memci_launch_count(memci_ctx *ctx);
MEMCI_API int memci_reset_launch_count(memci_ctx *ctx);
/* Statistics of the last memci_me_fs call: candidates evaluated, SAD pixel-ops
 * (candidates * block_size^2) and blocks re-resolved by the exact float64 path. */
MEMCI_API int memci_fs_stats(memci_ctx *ctx, long long *n_candidates, long long *px_ops,
                             int *n_redo_blocks);
/* After a full search that took the two-phase path (8-bit prefilter + exact refinement, csrc/fs8.cu): how many
 * candidates survived the prefilter and were evaluated exactly, and whether the survivor list overflowed (the exact
 * kernel then searched the pair instead).  Both 0 if the last search did not use the prefilter. */
MEMCI_API int memci_fs_prefilter_stats(memci_ctx *ctx, int *n_survivors, int *overflowed);
/* Number of holes filled by the last MCI call (synchronises the stream). */
MEMCI_API int memci_mci_stats(memci_ctx *ctx, int *n_holes);
/* Device-side time of the last memci_me_fs main kernel / MCI splat kernel, measured
 * with CUDA events on the context stream when timing is enabled (ms, <0 if none). */
MEMCI_API int memci_enable_kernel_timing(memci_ctx *ctx, int on);
MEMCI_API int memci_last_kernel_ms(memci_ctx *ctx, float *fs_ms, float *mci_ms);

/* Per-launch device timing of the two hot kernels inside a caller's timed region: after
 * memci_profile_start(capacity) every full-search main-kernel launch and every MCI splat launch
 * is bracketed by a CUDA event pair on the context stream (no synchronisation); collect sums
 * the elapsed times (ms) and counts, then stops recording. */
MEMCI_API int memci_profile_start(memci_ctx *ctx, int capacity);
MEMCI_API int memci_profile_collect(memci_ctx *ctx, double *fs_ms_sum, int *fs_launches,
                                    double *mci_ms_sum, int *mci_launches);
/* The same recording, broken down by kernel family.  Each of the three arrays has MEMCI_PROF_KINDS entries
 * (any may be NULL): summed device time (ms), launches, and the ALGORITHMIC work of those launches as the
 * library counts it -- SAD pixel-ops for the search kernels (the reference's own candidate bounds),
 * bytes for the HBM-side kernels (frames read + written; 9 B/px for an interpolated frame).  These are the
 * numerators of bench.py's per-workload roofline lines; no reference counterpart (the reference only has
 * wall-clock prints, src/Interpolators/base.py:84,108). */
#define MEMCI_PROF_FS        0   /* dominant full-search kernel: fs8_prefilter_kernel, or fs_tiled_kernel when the prefilter is off */
#define MEMCI_PROF_MCI       1   /* mci_splat_kernel (uni / bidir warp + blend) */
#define MEMCI_PROF_HBMA      2   /* hbma_fs_kernel + hbma_densify*_kernel (coarse search and every densification) */
#define MEMCI_PROF_HBM_SMALL 3   /* pyr_down_kernel + luma_kernel of the pyramid levels + smooth_kernel */
#define MEMCI_PROF_REFINE    4   /* two-phase search after the prefilter: fs8_scan_kernel, fs8_refine_kernel + fs8_replay_kernel + fs8_output_kernel,
                                    the gated fs_tiled_kernel launch + fs_redo_kernel (one entry each); prefilter off: fs_redo_kernel */
#define MEMCI_PROF_HOLEFILL  5   /* mci_holefill_kernel */
#define MEMCI_PROF_PREP      6   /* luma_kernel / luma8_kernel of a full-size frame */
#define MEMCI_PROF_BIDIR2    7   /* b2_blocks_kernel + b2_frame_kernel */
#define MEMCI_PROF_KINDS     8
MEMCI_API int memci_profile_collect_kinds(memci_ctx *ctx, double *ms_sum, int *launches, double *work);

/* Issue-rate microbenchmark for the SAD kernel's roofline denominator:
 * mode 0 = integer |a-b|+c (VABSDIFF), 1 = the fp32 two-FADD form, 2/3 = both pipes mixed.
 * Returns the device time of one launch and the SAD pixel-ops it performed. */
MEMCI_API int memci_microbench(memci_ctx *ctx, int mode, int iters, float *ms, double *px_ops);

#ifdef __cplusplus
}
#endif
#endif /* MEMCI_B200_H */
