"""B200-native MEMCI hot path of lhl2617/Interpolatory: block-matching motion estimation
(full search, HBMA), MV smoothing and motion-compensated interpolation behind the
reference's own interpolator API.  CUDA (sm_100a) through a C-ABI shared library; no CPU
fallback."""
from .MCI.Interpolator import MEMCI
from .MCI.Unidirectional import UniDirInterpolator
from .MCI.Bidirectional import BiDirInterpolator
from .VideoStream import ArrayVideoStream, BaseVideoStream, BenchmarkVideoStream
from .device import DeviceContext, PinnedBuffer

InterpolatorDictionary = {'MEMCI': MEMCI}

__all__ = ["MEMCI", "UniDirInterpolator", "BiDirInterpolator", "InterpolatorDictionary", "DeviceContext",
           "PinnedBuffer", "ArrayVideoStream", "BaseVideoStream", "BenchmarkVideoStream"]
