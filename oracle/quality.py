"""TEST INFRASTRUCTURE ONLY (see oracle/README or DESIGN.md section 6): CPU restatement of the two scikit-image
metrics the reference's Middlebury harness calls (src/Benchmark.py:46-47, :105-106).

The algorithm lives in a third-party dependency that is absent from /root/reference and from this image:
scikit-image, required as `scikit-image>=0.16.2` by the reference's requirements.txt (no pinned version).  This
file restates its published algorithm --
    skimage/metrics/simple_metrics.py      mean_squared_error, peak_signal_noise_ratio
    skimage/metrics/_structural_similarity.py   structural_similarity (Wang et al. 2004, uniform 7x7 window)
-- with the same building blocks (float64 images, scipy.ndimage.uniform_filter, sample covariance, crop by
(win_size - 1) // 2, per-channel mean then mean over channels).  PARITY UNPINNED for this row: skimage cannot be
run here, so the restatement is checked against a direct evaluation of the SSIM definition (tests) and known
answers (identical frames -> SSIM 1, PSNR inf; constant offset -> PSNR in closed form), not against skimage.
"""
import numpy as np
from scipy.ndimage import uniform_filter


def mean_squared_error(image0, image1):
    a, b = np.asarray(image0, np.float64), np.asarray(image1, np.float64)     # _as_floats
    return np.mean((a - b) ** 2, dtype=np.float64)


def peak_signal_noise_ratio(image_true, image_test, data_range=255):
    err = mean_squared_error(image_true, image_test)
    with np.errstate(divide='ignore'):
        return float(10 * np.log10((data_range ** 2) / err))


def _ssim_channel(X, Y, data_range, win_size=7, K1=0.01, K2=0.03):
    X = X.astype(np.float64); Y = Y.astype(np.float64)
    NP = win_size ** X.ndim
    cov_norm = NP / (NP - 1)                       # use_sample_covariance=True
    ux = uniform_filter(X, size=win_size); uy = uniform_filter(Y, size=win_size)
    uxx = uniform_filter(X * X, size=win_size); uyy = uniform_filter(Y * Y, size=win_size); uxy = uniform_filter(X * Y, size=win_size)
    vx = cov_norm * (uxx - ux * ux); vy = cov_norm * (uyy - uy * uy); vxy = cov_norm * (uxy - ux * uy)
    C1 = (K1 * data_range) ** 2; C2 = (K2 * data_range) ** 2
    A1, A2, B1, B2 = 2 * ux * uy + C1, 2 * vxy + C2, ux ** 2 + uy ** 2 + C1, vx + vy + C2
    S = (A1 * A2) / (B1 * B2)
    pad = (win_size - 1) // 2
    return S[pad:S.shape[0] - pad, pad:S.shape[1] - pad].mean(dtype=np.float64)


def structural_similarity(im1, im2, data_range=255, multichannel=True):
    im1, im2 = np.asarray(im1), np.asarray(im2)
    if im1.shape != im2.shape:
        raise ValueError('Input images must have the same dimensions.')
    if min(im1.shape[:2]) < 7:
        raise ValueError('win_size exceeds image extent.')
    if not multichannel:
        return float(_ssim_channel(im1, im2, data_range))
    return float(np.mean([_ssim_channel(im1[..., ch], im2[..., ch], data_range) for ch in range(im1.shape[-1])]))


def ssim_by_definition(im1, im2, data_range=255):
    """The SSIM definition evaluated window by window (slow; small frames only): the check on the restatement."""
    a, b = np.asarray(im1, np.float64), np.asarray(im2, np.float64)
    H, W, C = a.shape
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    tot = 0.0
    for ch in range(C):
        for r in range(3, H - 3):
            for c in range(3, W - 3):
                x = a[r - 3:r + 4, c - 3:c + 4, ch].ravel(); y = b[r - 3:r + 4, c - 3:c + 4, ch].ravel()
                mx, my = x.mean(), y.mean()
                vx, vy = x.var(ddof=1), y.var(ddof=1)
                cxy = ((x - mx) * (y - my)).sum() / 48.0
                tot += ((2 * mx * my + C1) * (2 * cxy + C2)) / ((mx * mx + my * my + C1) * (vx + vy + C2))
    return tot / (C * (H - 6) * (W - 6))
