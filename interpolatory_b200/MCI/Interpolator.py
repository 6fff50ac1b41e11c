"""MEMCI factory: the plug-in point InterpolatorDictionary['MEMCI'] of the reference
(src/Interpolator.py:15 -> MCI/Interpolator.py:10)."""
import math

from .Bidirectional import BiDirInterpolator
from .Bidirectional_2 import BiDir2Interpolator
from .Unidirectional import UniDirInterpolator


def MEMCI(target_fps, video_in_path=None, video_out_path=None, max_out_frames=math.inf, max_cache_size=2, **args):
    mci_mode = args.get('mci_mode', 'unidir')
    if mci_mode == 'unidir':
        return UniDirInterpolator(target_fps, video_in_path, video_out_path, max_out_frames, max_cache_size, **args)
    if mci_mode == 'bidir':
        return BiDirInterpolator(target_fps, video_in_path, video_out_path, max_out_frames, max_cache_size, **args)
    if mci_mode == 'bidir2':
        args['filter_mode'] = 'none'                                        # no smoothing in bidir2 (Interpolator.py:27)
        return BiDir2Interpolator(target_fps, video_in_path, video_out_path, max_out_frames, max_cache_size, **args)
    raise Exception(f'Unknown mci_mode: {mci_mode}')
