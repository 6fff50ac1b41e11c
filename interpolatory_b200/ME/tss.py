"""Three-step search on the GPU; same call as the reference's ME/tss.py:63
get_motion_vectors(block_size, steps, im1, im2)."""
import numpy as np

from ..device import get_context


def get_motion_vectors(block_size, steps, im1, im2, device=0):
    """Dense float32 [H, W, 3] field of (dy, dx, sad): 3x3 pattern at spacing 2^(step-1) per step,
    precedence-table tie-break against the running best (tss.py:21,51), inclusive target bounds."""
    im1 = np.asarray(im1)
    H, W, _ = im1.shape
    ctx = get_context(H, W, device)
    a = ctx.upload_frame(0, im1)
    b = ctx.upload_frame(1, im2)
    ctx.me_tss(0, 1, block_size, steps, 0)
    out = ctx.download_mv_dense(0)
    del a, b
    return out
