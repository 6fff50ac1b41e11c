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


class ReaderVideoStream(BaseVideoStream):
    """Frames decoded on demand from a reader object -- anything with get_meta_data() and get_data(idx), i.e. what
    imageio.get_reader(..., 'ffmpeg') returns -- behind a sliding window of at most `max_cache_size` decoded frames
    (reference: VideoStream, src/VideoStream.py:29-107).  The window only moves forward, one frame at a time; a frame
    before it, or more than one past its end, is decoded straight from the reader without touching the window, and
    max_cache_size == 0 disables the window.  Host memory is bounded by the window, not by the length of the clip."""

    def __init__(self, reader, max_cache_size, preload=True):
        if max_cache_size < 0:
            raise Exception('max_cache_size needs to be at least 0')
        self.video = reader
        meta = reader.get_meta_data()
        self.video_metadata = meta
        self.fps = meta['fps']
        self.size = meta.get('size')
        self.duration = meta['duration']                  # what max_frames_possible is derived from (base.py:60-61)
        n = meta.get('nframes')
        if n is None or n == float('inf'):                # imageio-ffmpeg reports inf until the stream has been counted
            n = reader.count_frames() if hasattr(reader, 'count_frames') else int(round(self.duration * self.fps))
        self.nframes = int(n)
        self.max_cache_size = int(max_cache_size)
        self._window = []                                 # decoded frames first .. first + len - 1
        self._first = 0
        if preload:
            for i in range(min(self.max_cache_size, self.nframes)):
                self.get_frame(i)

    def get_frame(self, idx):
        if self.max_cache_size == 0:
            return self.video.get_data(idx)
        end = self._first + len(self._window)            # index the window would take next
        if self._first <= idx < end:
            return self._window[idx - self._first]
        if idx != end:                                    # behind the window, or a jump ahead: no caching
            return self.video.get_data(idx)
        if len(self._window) == self.max_cache_size:
            self._window.pop(0)
            self._first += 1
        frame = self.video.get_data(idx)
        self._window.append(frame)
        return frame


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

    def __init__(self, frames, fps, page_locked=False):
        self.frames = frames
        self.page_locked = page_locked                    # frames live in page-locked memory (device.PinnedBuffer): staged in place
        self.nframes = len(frames)
        self.fps = fps
        self.duration = self.nframes / fps
        f0 = np.asarray(frames[0])
        self.size = (f0.shape[1], f0.shape[0])

    def get_frame(self, idx):
        return self.frames[idx]
