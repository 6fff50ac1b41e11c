"""Settings, dispatch tables and device plumbing shared by the MEMCI interpolators
(reference: MCI/MEMCIBaseInterpolator.py:14-98).  Settings arrive as strings, exactly as the
CLI's `MEMCI:key=val.key=val` syntax delivers them."""
import math

import numpy as np

from ..base import BaseInterpolator
from ..device import DeviceContext, FILTER_MODES
from .._native import PairParams
from ..ME.full_search import get_motion_vectors as fs
from ..ME.hbma import get_motion_vectors as hbma
from ..ME.tss import get_motion_vectors as tss
from ..smoothing import mean_filter, median_filter, weighted_mean_filter, filter_mode_of
from ..util import hbma_auto_steps, get_first_frame_idx_and_ratio, warn_if_hbma_reference_undefined

# device slot plan of one interpolator instance
_FRAME_SLOTS = (0, 1, 2)      # resident source frames (LRU)
_OUT_SLOT = 7
_MV_FWD, _MV_BWD = 0, 1


class MEMCIBaseInterpolator(BaseInterpolator):
    def __init__(self, target_fps, video_in_path=None, video_out_path=None, max_out_frames=math.inf,
                 max_cache_size=2, **args):
        super().__init__(target_fps, video_in_path, video_out_path, max_out_frames, max_cache_size, **args)
        ME_dict = {"fs": fs, "tss": tss, "hbma": hbma}
        smoothing_dict = {"mean": mean_filter, "median": median_filter, "weighted": weighted_mean_filter, "none": None}
        self.MV_field_idx = -1
        self.fwr_MV_field = []
        self.bwr_MV_field = []
        self.MV_field = []
        self.block_size = 8
        self.region = 7
        self.me_mode = ME_dict["hbma"]
        self.filter_mode = smoothing_dict["none"]
        self.filter_size = 4
        self.sub_region = 1
        self.steps = 3
        self.min_block_size = 4
        self.upscale_MV = True

        if 'block_size' in args:
            self.block_size = int(args['block_size'])
        if 'target_region' in args:
            self.region = int(args['target_region'])
        if 'me_mode' in args:
            self.me_mode = ME_dict[args['me_mode']]             # KeyError on unknown modes, like the reference
            if self.me_mode is tss:
                self.region = self.steps
        if 'filter_mode' in args:
            self.filter_mode = smoothing_dict[args['filter_mode']]
        if 'filter_size' in args:
            self.filter_size = int(args['filter_size'])

        self.pad_size = 4 * self.min_block_size
        if self.me_mode is fs:
            self.ME_args = {"block_size": self.block_size, "target_region": self.region}
        elif self.me_mode is tss:
            self.ME_args = {"block_size": self.block_size, "steps": self.steps}
        elif self.me_mode is hbma:
            self.upscale_MV = False
            self.ME_args = {"block_size": self.block_size, "win_size": self.region,
                            "sub_win_size": self.sub_region, "steps": self.steps,
                            "min_block_size": self.min_block_size, "cost_str": "sad",
                            "upscale": self.upscale_MV}
        self.device_index = int(args.get('device', 0))
        # the throughput path behind interpolate_video(): the same settings strings, optionally several GPUs
        # (devices='0,1,...': pairs are partitioned by index, multigpu.MultiGpuPipeline) and a window of source frames
        # per batch (stream_window, default max_cache_size)
        self._settings = {k: args[k] for k in ('me_mode', 'block_size', 'target_region', 'mci_mode', 'filter_mode', 'filter_size') if k in args}
        self._devices = [int(d) for d in str(args['devices']).split(',')] if 'devices' in args else [self.device_index]
        self._stream_window = int(args['stream_window']) if 'stream_window' in args else None
        self._pipe = None
        self._pin = None
        self._dev = None
        self._resident = {}
        self._resident_stream = None

    # ---- device plumbing ------------------------------------------------------------------
    def _device(self, shape):
        H, W = int(shape[0]), int(shape[1])
        if self._dev is None or (self._dev.H, self._dev.W) != (H, W):
            if self._dev is not None:
                self._dev.close()
            self._dev = DeviceContext(H, W, self.device_index)
            self._resident = {}
        return self._dev

    def _frame_slot(self, dev, frame_idx, frame, keep=()):
        """Device slot holding source frame `frame_idx` (uploaded once per stream position).  Least recently
        used eviction; the frames named in `keep` (the other frame of the pair being interpolated) are never
        evicted, so random access -- get_interpolated_frame(idx) with idx going backwards, which the reference's
        VideoStream supports (src/VideoStream.py:61-96) -- cannot make both frames of a pair share one slot."""
        if self._resident_stream is not self.video_stream:
            self._resident = {}
            self._resident_stream = self.video_stream
        if frame_idx in self._resident:
            slot = self._resident.pop(frame_idx)                 # re-insert: dict order = recency
            self._resident[frame_idx] = slot
            return slot
        used = set(self._resident.values())
        free = [s for s in _FRAME_SLOTS if s not in used]
        if free:
            slot = free[0]
        else:
            victim = next(k for k in self._resident if k not in keep)     # least recently used, never a pinned frame
            slot = self._resident.pop(victim)
        dev.upload_frame(slot, frame)
        dev.sync()                                               # pageable host memory: copy is done
        self._resident[frame_idx] = slot
        return slot

    def _pair_params(self, shape, bidir):
        p = PairParams()
        p.me_mode = 0 if self.me_mode is fs else (2 if self.me_mode is tss else 1)
        p.block_size = self.block_size
        p.region = self.region
        p.sub_region = self.sub_region
        if self.me_mode is hbma:
            self.steps = hbma_auto_steps(shape)                  # Unidirectional.py:116-121 / Bidirectional.py:144-149
            warn_if_hbma_reference_undefined(shape[0], shape[1], self.block_size, self.steps, self.min_block_size)
        p.steps = self.steps
        p.min_block_size = self.min_block_size
        p.filter_mode = FILTER_MODES[filter_mode_of(self.filter_mode)]
        p.filter_size = self.filter_size
        p.bidir = 1 if bidir else 0
        return p

    def clear_cache(self):
        self._resident = {}

    def close(self):
        if self._dev is not None:
            self._dev.close()
            self._dev = None
        if self._pipe is not None:
            for p_ in self._pipe:
                p_.close()
            self._pipe = None
        if self._pin is not None:
            for b in self._pin:
                b.close()
            self._pin = None

    def _mv_dense(self, slot):
        return self._dev.download_mv_dense(slot)

    # ---- interpolate_video(): the drop-in call is the fast path ---------------------------------------------
    def _output_frames(self, n, copy=True):
        """The frames get_interpolated_frame(0 .. n-1) would return, in order and byte for byte, produced in windows of
        source frames: a window is copied into page-locked memory as the stream delivers it (or used in place when the
        stream says its frames are page-locked already), its frame pairs go through the staged multi-stream pipeline
        (pipeline.PairPipeline, or multigpu.MultiGpuPipeline when several devices were given: each source frame crosses
        PCIe once, uploads / kernels / downloads overlap), pass-through frames are copied on the host.  Two pipelines
        alternate, so window w + 1 is decoded, staged and computed while window w's frames are handed to the sink.
        Host and device memory are bounded by the window (max_cache_size source frames, as in the reference's
        VideoStream; `stream_window` overrides), not by the clip.  Output frames the reference would build from a frame
        past the end of the stream (its reader wraps around, base.py:52-53) are left to the per-frame call.
        copy=False yields views into the staging buffers, valid until the next frame is requested two windows later
        (what a video writer needs); the default hands out independent arrays."""
        import numpy as np
        from ..device import PinnedBuffer
        from ..pipeline import PairPipeline
        from ..multigpu import MultiGpuPipeline
        vs = self.video_stream
        n_src = int(vs.nframes)
        in_place = bool(getattr(vs, 'page_locked', False))
        plan = []                                        # (source frame index, dist) exactly as get_interpolated_frame derives them
        for i in range(n):
            first, inv = get_first_frame_idx_and_ratio(i, self.rate_ratio)
            dist = 1. - inv
            plan.append((first, 0.0 if math.isclose(dist, 0.) else float(dist)))
        win = max(2, self._stream_window if self._stream_window is not None else int(self.max_cache_size))
        n_out_max = (win - 1) * (int(math.ceil(self.rate_ratio)) + 1) + 2

        def ensure(H, W):
            if self._pipe is None or (self._pipe[0].H, self._pipe[0].W) != (H, W):
                self.close()
                cls = MultiGpuPipeline if len(self._devices) > 1 else PairPipeline
                kw = dict(devices=self._devices) if len(self._devices) > 1 else dict(device=self._devices[0])
                self._pipe = [cls(H, W, n_contexts=3, **kw, **self._settings) for _ in range(2)]
            if self._pin is None or self._pin[0].shape[0] < win or self._pin[1].shape[0] < n_out_max:
                if self._pin is not None:
                    for b in self._pin:
                        b.close()
                self._pin = [PinnedBuffer((win, H, W, 3)), PinnedBuffer((n_out_max, H, W, 3)),
                             PinnedBuffer((win, H, W, 3)), PinnedBuffer((n_out_max, H, W, 3))]

        def enqueue(slot, i, j, w0, w1):
            """outputs i .. j-1 from source frames w0 .. w1 on pipeline `slot`; returns what drain() needs"""
            f0 = np.asarray(vs.get_frame(w0))
            ensure(int(f0.shape[0]), int(f0.shape[1]))
            out = self._pin[2 * slot + 1].array
            if in_place:
                src = [f0] + [np.asarray(vs.get_frame(k)) for k in range(w0 + 1, w1 + 1)]
            else:
                src = self._pin[2 * slot].array
                src[0] = f0
                for k in range(w0 + 1, w1 + 1):
                    src[k - w0] = vs.get_frame(k)
            per_pair = {}
            for o in range(i, j):
                k, d = plan[o]
                if d != 0.0:
                    per_pair.setdefault(k - w0, []).append((o - i, d))
            pipe = self._pipe[slot]
            if isinstance(pipe, MultiGpuPipeline):
                pipe._dispatch(src, per_pair, out)
            else:
                pipe._run_pairs(src, per_pair, out)
            return slot, i, j, w0, src, out

        def drain(job):
            slot, i, j, w0, src, out = job
            self._pipe[slot].sync()
            for o in range(i, j):
                k, d = plan[o]
                f = src[k - w0] if d == 0.0 else out[o - i]
                yield f.copy() if copy else f

        pending, slot, i = None, 0, 0
        while i < n:
            w0 = plan[i][0]
            w1 = min(w0 + win - 1, n_src - 1)             # source frames w0 .. w1 form this window
            j = i                                        # outputs i .. j-1 live entirely inside it
            while j < n and plan[j][0] >= w0 and (plan[j][0] + 1 <= w1 if plan[j][1] != 0.0 else plan[j][0] <= w1):
                j += 1
            if j == i:                                   # needs a frame past the end of the stream: whatever the stream does there
                if pending is not None:
                    yield from drain(pending)
                    pending = None
                yield self.get_interpolated_frame(i)
                i += 1
                continue
            job = enqueue(slot, i, j, w0, w1)
            if pending is not None:
                yield from drain(pending)
            pending, slot, i = job, slot ^ 1, j
        if pending is not None:
            yield from drain(pending)
