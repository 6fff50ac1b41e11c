"""Unidirectional MEMCI on the GPU (reference: MCI/Unidirectional.py)."""
import math

import numpy as np

from ..device import get_context
from ..util import get_first_frame_idx_and_ratio
from .MEMCIBaseInterpolator import MEMCIBaseInterpolator, _MV_FWD, _MV_BWD, _OUT_SLOT


def helper(source_frame, target_frame, MV_field, dist, device=0):
    """Same call as Unidirectional.py:36 helper(source_frame, target_frame, MV_field, dist):
    forward splat with `<=` SAD priority and self-write of out-of-frame targets, then the 21x21
    masked-median hole fill.  MV_field is any dense float32 [H, W, 3] field."""
    src = np.asarray(source_frame)
    H, W, _ = src.shape
    ctx = get_context(H, W, device)
    a = ctx.upload_frame(0, src)
    b = ctx.upload_frame(1, target_frame)
    ctx.upload_mv_dense(0, MV_field)
    ctx.mci_unidir(0, 1, 0, np.float32(dist), 7)
    out = ctx.download_frame(7)
    del a, b
    return out


class UniDirInterpolator(MEMCIBaseInterpolator):
    def get_interpolated_frame(self, idx):
        source_frame_idx, inv_dist = get_first_frame_idx_and_ratio(idx, self.rate_ratio)
        dist = 1. - inv_dist
        source_frame = self.video_stream.get_frame(source_frame_idx)
        if math.isclose(dist, 0.):                       # output frame coincides with a source frame
            return source_frame
        target_frame = self.video_stream.get_frame(source_frame_idx + 1)
        dev = self._device(np.asarray(source_frame).shape)
        s = self._frame_slot(dev, source_frame_idx, source_frame)
        t = self._frame_slot(dev, source_frame_idx + 1, target_frame, keep=(source_frame_idx,))
        # motion field cache: valid while idx stays strictly between its two source frames (:109)
        if not self.MV_field_idx < idx / self.rate_ratio < self.MV_field_idx + 1:
            dev.estimate_pair(self._pair_params(np.asarray(source_frame).shape, False), s, t, _MV_FWD, _MV_BWD)
            self.MV_field_idx = source_frame_idx
        dev.mci_unidir(s, t, _MV_FWD, np.float32(dist), _OUT_SLOT)
        return dev.download_frame(_OUT_SLOT)

    @property
    def MV_field_dense(self):
        """The dense float32 field the reference keeps in self.MV_field (materialised on demand)."""
        return self._mv_dense(_MV_FWD)

    def __str__(self):
        return 'uniMEMCI'
