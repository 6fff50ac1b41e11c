"""3x3 motion-vector smoothing on the GPU (reference: smoothing/threeXthree_mv_smoothing.py:14
smooth(filter_func, mv_field, block_size) with mean_filter / median_filter / weighted_mean_filter).

The three filter "functions" are tokens that select the device kernel's mode, so the reference's
dispatch table (MCI/MEMCIBaseInterpolator.py:23-29) and call sites keep their shape."""
import numpy as np

from ..device import FILTER_MODES, get_context


class _Filter:
    def __init__(self, mode):
        self.mode = mode
        self.__name__ = f"{mode}_filter"

    def __repr__(self):
        return f"<memci filter {self.mode}>"

    def __call__(self, block):
        raise RuntimeError("filter tokens select a GPU kernel mode; call smooth(filter, mv_field, size)")


mean_filter = _Filter("mean")
median_filter = _Filter("median")
weighted_mean_filter = _Filter("weighted")


def filter_mode_of(filter_func):
    """Accept our tokens, mode strings, None, or the reference's own filter functions."""
    if filter_func is None:
        return "none"
    if isinstance(filter_func, str):
        if filter_func not in FILTER_MODES:
            raise KeyError(filter_func)
        return filter_func
    if isinstance(filter_func, _Filter):
        return filter_func.mode
    name = getattr(filter_func, "__name__", "")
    for mode, fn in (("weighted", "weighted_mean_filter"), ("mean", "mean_filter"), ("median", "median_filter")):
        if name == fn:
            return mode
    raise ValueError(f"unknown smoothing filter {filter_func!r}")


def smooth(filter_func, mv_field, block_size, device=0):
    """Dense float32 [H, W, 3] in, dense out; (dy, dx) rewritten per block_size cell from the 3x3
    strided neighbourhood, SAD untouched.  filter None returns the input object, like the reference."""
    mode = filter_mode_of(filter_func)
    if mode == "none":
        return mv_field
    mv = np.asarray(mv_field)
    H, W, _ = mv.shape
    ctx = get_context(H, W, device)
    ctx.upload_mv_dense(0, mv)
    ctx.smooth(0, mode, block_size)
    return ctx.download_mv_dense(0)
