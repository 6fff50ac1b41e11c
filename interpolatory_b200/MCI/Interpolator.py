"""MEMCI factory: the plug-in point InterpolatorDictionary['MEMCI'] of the reference
(src/Interpolator.py:15 -> MCI/Interpolator.py:10)."""
import math

from .Bidirectional import BiDirInterpolator
from .Unidirectional import UniDirInterpolator


def MEMCI(target_fps, video_in_path=None, video_out_path=None, max_out_frames=math.inf, max_cache_size=2, **args):
    mci_mode = args.get('mci_mode', 'unidir')
    if mci_mode == 'unidir':
        return UniDirInterpolator(target_fps, video_in_path, video_out_path, max_out_frames, max_cache_size, **args)
    if mci_mode == 'bidir':
        return BiDirInterpolator(target_fps, video_in_path, video_out_path, max_out_frames, max_cache_size, **args)
    if mci_mode == 'bidir2':
        raise NotImplementedError("mci_mode 'bidir2' (Bidirectional_2.py) is the next row of the hot-path scope "
                                  "table (SURVEY 8f #1) and is not built yet")
    raise Exception(f'Unknown mci_mode: {mci_mode}')
