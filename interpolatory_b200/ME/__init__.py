from .full_search import get_motion_vectors as fs
from .hbma import get_motion_vectors as hbma
from .tss import get_motion_vectors as tss

__all__ = ["fs", "hbma", "tss"]
