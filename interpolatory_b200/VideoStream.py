"""Video streams (reference: src/VideoStream.py).  A stream only has to provide
nframes / fps / duration and get_frame(idx) -> uint8 [H, W, 3]."""
import numpy as np


class BaseVideoStream:
    def __init__(self):
        self.nframes = None
        self.fps = None
        self.size = None
        self.duration = None

    def get_frame(self, idx):
        raise NotImplementedError('get_frame not implemented in VideoStream!')


class BenchmarkVideoStream(BaseVideoStream):
    """Two frames posing as a 2 s, 1 fps clip (src/VideoStream.py:109-129)."""

    def __init__(self, frame_1, frame_2):
        self.frame_1 = frame_1
        self.frame_2 = frame_2
        self.nframes = 2
        self.fps = 1
        self.duration = 2

    def get_frame(self, idx):
        if idx == 0:
            return self.frame_1
        if idx == 1:
            return self.frame_2
        raise Exception(f'Tried to get invalid frame idx {idx}')


class ArrayVideoStream(BaseVideoStream):
    """In-memory frames (decoded elsewhere, or synthetic) at a nominal frame rate."""

    def __init__(self, frames, fps):
        self.frames = frames
        self.nframes = len(frames)
        self.fps = fps
        self.duration = self.nframes / fps
        f0 = np.asarray(frames[0])
        self.size = (f0.shape[1], f0.shape[0])

    def get_frame(self, idx):
        return self.frames[idx]
