"""Bidirectional MEMCI on the GPU (reference: MCI/Bidirectional.py)."""
import math

import numpy as np

from ..device import get_context
from ..util import get_first_frame_idx_and_ratio
from .MEMCIBaseInterpolator import MEMCIBaseInterpolator, _MV_FWD, _MV_BWD, _OUT_SLOT


def helper(source_frame, target_frame, fwr_MV_field, bwr_MV_field, dist, device=0):
    """Same call as Bidirectional.py:27 helper(source, target, fwr_MV_field, bwr_MV_field, dist):
    interleaved forward/backward splat in source raster order, strict SAD priority, SAD-weighted
    blend where a backward vector lands on an occupied pixel, then the masked-median hole fill."""
    src = np.asarray(source_frame)
    H, W, _ = src.shape
    ctx = get_context(H, W, device)
    a = ctx.upload_frame(0, src)
    b = ctx.upload_frame(1, target_frame)
    ctx.upload_mv_dense(0, fwr_MV_field)
    ctx.upload_mv_dense(1, bwr_MV_field)
    ctx.mci_bidir(0, 1, 0, 1, np.float32(dist), 7)
    out = ctx.download_frame(7)
    del a, b
    return out


class BiDirInterpolator(MEMCIBaseInterpolator):
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
        if not self.MV_field_idx < idx / self.rate_ratio < self.MV_field_idx + 1:      # :138
            dev.estimate_pair(self._pair_params(np.asarray(source_frame).shape, True), s, t, _MV_FWD, _MV_BWD)
            self.MV_field_idx = source_frame_idx
        dev.mci_bidir(s, t, _MV_FWD, _MV_BWD, np.float32(dist), _OUT_SLOT)
        return dev.download_frame(_OUT_SLOT)

    @property
    def fwr_MV_field_dense(self):
        return self._mv_dense(_MV_FWD)

    @property
    def bwr_MV_field_dense(self):
        return self._mv_dense(_MV_BWD)

    def __str__(self):
        return 'bwMEMCI'
