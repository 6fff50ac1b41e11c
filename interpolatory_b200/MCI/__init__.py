from .Interpolator import MEMCI
from .Unidirectional import UniDirInterpolator
from .Bidirectional import BiDirInterpolator

__all__ = ["MEMCI", "UniDirInterpolator", "BiDirInterpolator"]
