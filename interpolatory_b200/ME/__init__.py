from .full_search import get_motion_vectors as fs
from .hbma import get_motion_vectors as hbma

__all__ = ["fs", "hbma"]
