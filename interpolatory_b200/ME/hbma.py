"""Hierarchical block matching on the GPU; same call as the reference's ME/hbma.py:161
get_motion_vectors(block_size, win_size, sub_win_size, steps, min_block_size, im1, im2,
cost_str='sad', upscale=True)."""
import numpy as np

from ..device import get_context
from ..util import warn_if_hbma_reference_undefined

cost_key_map = {'sad': 0, 'ssd': 1}


def get_motion_vectors(block_size, win_size, sub_win_size, steps, min_block_size, im1, im2,
                       cost_str='sad', upscale=True, device=0):
    """upscale=True: dense float32 [H, W, 3]; upscale=False: block-level
    [ceil(H/b), ceil(W/b), 3] at the final block size b (hbma.py:184-190)."""
    cost_key = cost_key_map[cost_str]
    im1 = np.asarray(im1)
    H, W, _ = im1.shape
    warn_if_hbma_reference_undefined(H, W, block_size, steps, min_block_size)
    ctx = get_context(H, W, device)
    a = ctx.upload_frame(0, im1)
    b = ctx.upload_frame(1, im2)
    ctx.me_hbma(0, 1, block_size, win_size, sub_win_size, steps, min_block_size, 0, cost_key)
    if upscale:
        out = ctx.download_mv_dense(0)
    else:
        _, vec, _, sad = ctx.download_mv_blocks(0)
        out = np.concatenate([vec, sad[:, :, None]], axis=2)
    del a, b
    return out
