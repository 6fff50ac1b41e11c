"""Interpolator base class: the caller side of the drop-in boundary (reference:
src/Interpolators/base.py:21-161).  Only what the MEMCI path needs is kept: construction,
the output-frame loop, and the two-frame benchmark entry point."""
import math
import time

from .VideoStream import ArrayVideoStream, BenchmarkVideoStream, ReaderVideoStream
from . import util


class BaseInterpolator(object):
    def __init__(self, target_fps, video_in_path=None, video_out_path=None, max_out_frames=math.inf,
                 max_cache_size=0, **args):
        self.target_fps = target_fps
        self.video_in_path = video_in_path
        self.video_out_path = video_out_path
        self.max_out_frames = max_out_frames
        self.max_cache_size = max_cache_size
        self.video_stream = None
        self.video_out_writer = None
        self.max_frames_possible = None
        self.rate_ratio = None
        for arg, value in args.items():
            setattr(self, arg, value)
        if video_in_path is not None:
            self.load_input_video()
        if video_out_path is not None:
            self.get_video_out_writer()

    # -- I/O seam.  Decode/encode is out of scope: video files need imageio-ffmpeg.
    def load_input_video(self):
        if self.video_in_path is None:
            raise Exception('video input path not given')
        try:
            import imageio
        except ImportError as e:
            raise Exception('reading video files needs imageio + imageio-ffmpeg; '
                            'use set_video_stream() with decoded frames instead') from e
        # loop=True: reading one frame past the end wraps around, as in the reference (base.py:52-53).  Frames are decoded
        # on demand behind a window of max_cache_size frames -- never the whole clip in host memory.
        vid = imageio.get_reader(self.video_in_path, 'ffmpeg', loop=True)
        self.set_video_stream(ReaderVideoStream(vid, self.max_cache_size))

    def set_video_stream(self, stream):
        self.video_stream = stream
        self.max_frames_possible = math.floor(stream.duration * self.target_fps)
        self.rate_ratio = self.target_fps / stream.fps

    def get_video_out_writer(self):
        if self.video_out_path is None:
            raise Exception('video output path not given')
        try:
            import imageio
        except ImportError as e:
            raise Exception('writing video files needs imageio + imageio-ffmpeg') from e
        self.video_out_writer = imageio.get_writer(self.video_out_path, fps=self.target_fps,
                                                   macro_block_size=0, quality=6)

    def interpolate_video(self, sink=None):
        """Produce every output frame in order (base.py:78-115), with the reference's progress reporting (:91-101,
        :110-114: a PROGRESS line once per output second, to the progress file the GUI polls and, when the debug flag is
        set, to stdout).  Frames go to the video writer, or to `sink(frame)` when given (e.g. a list's append)."""
        copy = True
        if sink is None:
            if self.video_in_path is None or self.video_out_path is None:
                raise Exception('Cannot interpolate video, no input video path or output video path')
            sink = self.video_out_writer.append_data
            copy = False                                  # the writer encodes a frame before it returns
        start = int(round(time.time()))
        target_nframes = int(min(self.max_frames_possible, self.max_out_frames))
        for i, frame in enumerate(self._output_frames(target_nframes, copy=copy and not getattr(self, 'sink_keeps_no_reference', False))):
            if (i % self.target_fps) == 0:
                pct = str(math.floor(100 * (i + 1) / target_nframes)).rjust(3, ' ')
                elapsed_seconds = int(int(round(time.time())) - start)
                progStr = (f'PROGRESS::{pct}%::Frame {i}/{target_nframes} | Time elapsed: {util.sToMMSS(elapsed_seconds)} | '
                           f'Estimated Time Left: {util.getETA(elapsed_seconds, i, target_nframes)}')
                util.signal_progress(progStr)
                if util.debug_flags['debug_progress']:
                    print(progStr)
            sink(frame)
        if self.video_out_writer is not None:
            self.video_out_writer.close()
        end = int(round(time.time()))
        progStr = f'PROGRESS::100%::Complete | Time taken: {util.sToMMSS(end - start)}'
        util.signal_progress(progStr)
        if util.debug_flags['debug_progress']:
            print(progStr)
        return end - start

    def _output_frames(self, n, copy=True):
        """Output frames 0 .. n-1 in order.  The base class asks for them one by one; the MEMCI interpolators override this
        with the staged, batched device path (same frames, byte for byte)."""
        for i in range(n):
            yield self.get_interpolated_frame(i)

    def get_interpolated_frame(self, idx):
        raise NotImplementedError('To be implemented by derived classes')

    def __str__(self):
        raise NotImplementedError('To be implemented by derived classes')

    def get_benchmark_frame(self, image_1, image_2):
        """Middle frame of (image_1, image_2): a 2 s / 1 fps stream interpolated to 2 fps,
        frame index 1 => dist 0.5 (base.py:136-157).  Like the reference, the interpolator is
        left pointing at the two-frame stream afterwards (its `self = backup` restores nothing)."""
        self.MV_field_idx = -1
        self.video_stream = BenchmarkVideoStream(image_1, image_2)
        self.target_fps = 2
        self.max_out_frames = 2
        self.rate_ratio = 2.
        self.max_frames_possible = 2
        self.clear_cache()
        return self.get_interpolated_frame(1)

    def clear_cache(self):
        pass
