"""PSNR / SSIM on the GPU for the Middlebury harness: the two skimage.metrics functions the reference's
src/Benchmark.py calls (:46-47, :105-106), same names and keyword arguments, for the one case it uses --
uint8 RGB frames, data_range=255, multichannel=True, default 7x7 uniform window.  Anything else is refused
loudly: there is no CPU implementation behind these functions."""
import numpy as np

from .device import get_context


def _pair(image_true, image_test):
    a, b = np.asarray(image_true), np.asarray(image_test)
    if a.shape != b.shape:
        raise ValueError('Input images must have the same dimensions.')
    if a.dtype != np.uint8 or b.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise NotImplementedError(f'GPU metrics take uint8 [H, W, 3] frames, got {a.dtype}{a.shape} / {b.dtype}{b.shape}')
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def _check_range(data_range):
    if data_range is not None and data_range != 255:
        raise NotImplementedError(f'GPU metrics are built for data_range=255 (uint8 frames), got {data_range}')


def frame_quality(image_true, image_test, device=0):
    """(mean squared error, mean SSIM) of two uint8 RGB frames, one upload and one kernel pass for both."""
    a, b = _pair(image_true, image_test)
    H, W, _ = a.shape
    if H < 7 or W < 7:
        raise ValueError('win_size exceeds image extent.')
    ctx = get_context(H, W, device)
    ka = ctx.upload_frame(5, a)
    kb = ctx.upload_frame(6, b)
    out = ctx.quality(5, 6)          # synchronises
    del ka, kb
    return out


def psnr_from_mse(mse, data_range=255):
    """skimage.metrics.peak_signal_noise_ratio's last line: identical frames give inf, like the reference run does."""
    with np.errstate(divide='ignore'):
        return float(10 * np.log10((data_range ** 2) / np.float64(mse)))


def mean_squared_error(image0, image1, device=0):
    return frame_quality(image0, image1, device)[0]


def peak_signal_noise_ratio(image_true, image_test, data_range=None, device=0):
    _check_range(data_range)
    return psnr_from_mse(frame_quality(image_true, image_test, device)[0])


def structural_similarity(im1, im2, win_size=None, gradient=False, data_range=None, multichannel=False,
                          gaussian_weights=False, full=False, device=0, **kwargs):
    _check_range(data_range)
    if not (multichannel or kwargs.pop('channel_axis', None) in (2, -1)):
        raise NotImplementedError('GPU SSIM is the multichannel (RGB) form the harness calls')
    if (win_size not in (None, 7)) or gradient or gaussian_weights or full or kwargs:
        raise NotImplementedError('GPU SSIM implements the defaults the harness uses: 7x7 uniform window, '
                                  'sample covariance, K1=0.01, K2=0.03, no gradient / full map')
    return frame_quality(im1, im2, device)[1]
