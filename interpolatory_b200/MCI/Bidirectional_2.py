"""bidir2 MEMCI on the GPU (reference: MCI/Bidirectional_2.py, after Wang et al., "Motion-Compensated
Frame Rate Up-Conversion -- Part II"): irregular-grid expanded-block weighted motion compensation in both
directions, error-adaptive combination, block-wise directional hole interpolation."""
import math

import numpy as np

from ..device import get_context
from ..util import get_first_frame_idx_and_ratio
from .MEMCIBaseInterpolator import MEMCIBaseInterpolator, _MV_FWD, _MV_BWD, _OUT_SLOT


def get_weight_kern(kernlen=8):
    """Same values as Bidirectional_2.py:33-47 (one channel of the reference's (k, k, 3) kernel; the three
    channels are copies).  Host side, NumPy, exactly the reference's expression."""
    x = np.linspace(1, 1.1, int(kernlen / 2))
    x = np.append(x, np.flip(x))
    xx, yy = np.meshgrid(x, x)
    weights = 6 * np.square(5 * (np.log2(xx) + np.log2(yy)))
    return np.ascontiguousarray(weights.astype(np.float32))


def _check_shifts(dev):
    _, n_bad = dev.bidir2_stats()
    if n_bad:
        # the reference's numpy slices clip at the padded array's border and its in-place adds then fail
        raise Exception(f'bidir2: {n_bad} block(s) shifted outside the padded frame (pad_size too small for these motion vectors)')


def interpolate(source_frame, target_frame, fwr_MV_field, bwr_MV_field, dist, pad_size=16, min_block_size=4,
                upscale_MV=True, device=0):
    """Function-level entry: IEWMC both ways, Error_adaptive_combination, BDHI, unpad_and_downconvert
    (Bidirectional_2.py:330-345) for MV fields given as arrays -- dense [H][W][3] when upscale_MV, block
    level [ceil(H/mbs)][ceil(W/mbs)][3] otherwise, as the reference indexes them (:63-67)."""
    src = np.asarray(source_frame)
    H, W, _ = src.shape
    ctx = get_context(H, W, device)
    a = ctx.upload_frame(0, src)
    b = ctx.upload_frame(1, target_frame)
    for slot, f in ((0, fwr_MV_field), (1, bwr_MV_field)):
        f = np.ascontiguousarray(f, np.float32)
        if upscale_MV:
            ctx.upload_mv_dense(slot, f)
        else:
            ctx.upload_mv_blocks(slot, min_block_size, f[:, :, :2], min_block_size, f[:, :, 2])
    ctx.mci_bidir2(0, 1, 0, 1, np.float32(dist), np.float32(1 - dist), get_weight_kern(2 * min_block_size), 7,
                   min_block_size, pad_size)
    out = ctx.download_frame(7)
    _check_shifts(ctx)
    del a, b
    return out


class BiDir2Interpolator(MEMCIBaseInterpolator):
    def get_interpolated_frame(self, idx):
        source_frame_idx, inv_dist = get_first_frame_idx_and_ratio(idx, self.rate_ratio)
        dist = 1. - inv_dist
        source_frame = self.video_stream.get_frame(source_frame_idx)
        if math.isclose(dist, 0.):
            return source_frame
        target_frame = self.video_stream.get_frame(source_frame_idx + 1)
        dev = self._device(np.asarray(source_frame).shape)
        s = self._frame_slot(dev, source_frame_idx, source_frame)
        t = self._frame_slot(dev, source_frame_idx + 1, target_frame, keep=(source_frame_idx,))
        if not self.MV_field_idx < idx / self.rate_ratio < self.MV_field_idx + 1:      # :308
            # hbma: auto depth (:311-317) and block-level vectors (upscale=False); no smoothing in bidir2
            dev.estimate_pair(self._pair_params(np.asarray(source_frame).shape, True), s, t, _MV_FWD, _MV_BWD)
            self.MV_field_idx = source_frame_idx
        # the two njit calls see float32(dist) and float32(1 - dist), each rounded from the Python float (:331-332)
        dev.mci_bidir2(s, t, _MV_FWD, _MV_BWD, np.float32(dist), np.float32(1 - dist),
                       get_weight_kern(2 * self.min_block_size), _OUT_SLOT, self.min_block_size, self.pad_size)
        out = dev.download_frame(_OUT_SLOT)
        _check_shifts(dev)
        return out

    @property
    def fwr_MV_field_dense(self):
        return self._mv_dense(_MV_FWD)

    @property
    def bwr_MV_field_dense(self):
        return self._mv_dense(_MV_BWD)

    def __str__(self):
        return 'Uni_2'
