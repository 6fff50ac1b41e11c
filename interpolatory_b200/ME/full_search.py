"""Full-search block matching on the GPU; same call as the reference's
ME/full_search.py:61 get_motion_vectors(block_size, target_region, im1, im2)."""
import numpy as np

from ..device import get_context


def get_motion_vectors(block_size, target_region, im1, im2, device=0):
    """Dense float32 [H, W, 3] field of (dy, dx, sad), bit-identical to the reference
    (luma SAD in float64 truncated to uint32, distance then scan-order tie-break, zero
    padding by block_size, and the row/col bound quirks of full_search.py:33-43)."""
    im1 = np.asarray(im1)
    H, W, _ = im1.shape
    ctx = get_context(H, W, device)
    a = ctx.upload_frame(0, im1)
    b = ctx.upload_frame(1, im2)
    ctx.me_fs(0, 1, block_size, target_region, 0)
    out = ctx.download_mv_dense(0)
    del a, b
    return out
