"""Throughput path behind the same settings as the MEMCI classes: frame pairs streamed through
several device contexts (one CUDA stream each) so the host->device copy of the next source frame
and the device->host copy of the previous output overlap the kernels of the current pair.

Frame pairs are independent: a context takes a contiguous run of pairs, so each source
frame crosses PCIe once (twice at the run boundaries); nothing is exchanged between contexts.
"""
import math

import numpy as np

from ._native import PairParams
from .device import DeviceContext, FILTER_MODES, FrameStager
from .util import get_first_frame_idx_and_ratio, hbma_auto_steps


def pair_params_from_settings(shape, **st):
    """String settings of the MEMCI plug-in (MEMCIBaseInterpolator.py:48-78) -> device parameters."""
    me = st.get('me_mode', 'hbma')
    if me not in ('fs', 'hbma', 'tss'):
        raise NotImplementedError(f"me_mode {me!r} is not part of the B200 hot path")
    mci = st.get('mci_mode', 'unidir')
    if mci not in ('unidir', 'bidir', 'bidir2'):
        raise Exception(f'Unknown mci_mode: {mci}')
    p = PairParams()
    p.me_mode = {'fs': 0, 'hbma': 1, 'tss': 2}[me]
    p.block_size = int(st.get('block_size', 8))
    p.region = 3 if me == 'tss' else int(st.get('target_region', 7))      # tss: region = steps (MEMCIBaseInterpolator.py:63-64)
    p.sub_region = 1
    p.steps = hbma_auto_steps(shape) if me == 'hbma' else 3
    p.min_block_size = 4
    p.filter_mode = FILTER_MODES['none' if mci == 'bidir2' else st.get('filter_mode', 'none')]      # no smoothing in bidir2
    p.filter_size = int(st.get('filter_size', 4))
    p.bidir = 1 if mci in ('bidir', 'bidir2') else 0
    return p


def interpolation_dists(rate_ratio):
    """dist values of the interpolated frames between two source frames, in output order, exactly as the
    class API derives them (util.get_first_frame_idx_and_ratio, W0): Python floats `1. - ratio` (float64
    arithmetic on the float32 ratio).  The device calls round them to float32 at the boundary like the
    reference's njit signatures do; bidir2 also needs float32(1 - dist) of the UNROUNDED value
    (Bidirectional_2.py:331-332), which is why the float64 is kept here."""
    out, idx = [], 1
    while True:
        first, ratio = get_first_frame_idx_and_ratio(idx, rate_ratio)
        if first >= 1:
            break
        dist = 1. - ratio
        if dist != 0.:
            out.append(float(dist))
        idx += 1
    return out


class PairPipeline:
    """Throughput path: the source frames of a batch are uploaded ONCE by a FrameStager (its own H2D stream, a
    ring of device buffers, in the order the pairs need them) and handed device-to-device to the contexts; each
    context (one CUDA stream) takes a contiguous run of pairs, so a frame's luma plane is computed once per
    context that uses it; interpolated frames drain through the stager's D2H stream.  Compute streams never wait
    for PCIe except for data they actually need, and successive batches pipeline (a ring slot is rewritten as
    soon as the previous batch's takers have copied it)."""

    def __init__(self, H, W, device=0, n_contexts=3, **settings):
        self.H, self.W, self.device = int(H), int(W), device
        self.params = pair_params_from_settings((H, W), **settings)
        self.bidir = bool(self.params.bidir)
        self.bidir2 = settings.get('mci_mode', 'unidir') == 'bidir2'
        if self.bidir2:
            from .MCI.Bidirectional_2 import get_weight_kern
            self._w = get_weight_kern(2 * self.params.min_block_size)
        self.ctxs = [DeviceContext(H, W, device) for _ in range(max(1, int(n_contexts)))]
        self.stager = None
        self._ring_cursor = 0
        self.uploads = 0                                  # source frames copied host -> device so far

    def close(self):
        if self.ctxs:
            self.sync()                                  # the D2H stream may still be reading a context's output slot
        for c in self.ctxs:
            c.close()
        self.ctxs = []
        if self.stager is not None:
            self.stager.close()
            self.stager = None

    def launch_count(self):
        return sum(c.launch_count() for c in self.ctxs)

    def sync(self):
        for c in self.ctxs:
            c.sync()
        if self.stager is not None:
            self.stager.sync()
        if self.bidir2:
            # like the class API (MCI/Bidirectional_2._check_shifts): the reference's numpy slices fail when a block is shifted
            # outside the padded frame (vectors beyond pad_size); the frames enqueued since the last sync are then invalid
            n_bad = sum(c.bidir2_out_of_range_total() for c in self.ctxs)
            if n_bad:
                raise Exception(f'bidir2: {n_bad} block(s) shifted outside the padded frame (pad_size too small for these motion vectors)')

    def plan_clip(self, n_frames, fps_in, fps_out):
        """Output schedule of a whole clip, exactly as BaseInterpolator.interpolate_video walks it (base.py:78-115):
        one entry per output frame, (source frame index, dist) with dist == 0 for a pass-through source frame."""
        rate_ratio = fps_out / fps_in
        n_out = int(math.floor((n_frames / fps_in) * fps_out))
        plan = []
        for idx in range(n_out):
            first, ratio = get_first_frame_idx_and_ratio(idx, rate_ratio)
            dist = 1. - ratio
            if math.isclose(dist, 0.):
                plan.append((first, 0.0))
            elif first + 1 < n_frames:
                plan.append((first, float(dist)))
            else:
                break                                     # the reference would read past the last frame here
        return plan

    def run_clip(self, frames, fps_in, fps_out, out, sync=True):
        """Any rate ratio (e.g. 24 -> 60, where the dists differ from pair to pair): fills out[i] for every output
        frame i of plan_clip(); pass-through frames are copied on the host, interpolated ones go through the same
        staged, multi-context path as run().  Returns the plan."""
        plan = self.plan_clip(len(frames), fps_in, fps_out)
        per_pair = {}
        for i, (k, d) in enumerate(plan):
            if d == 0.0:
                out[i] = frames[k]
            else:
                per_pair.setdefault(k, []).append((i, d))
        if per_pair:
            self._run_pairs(frames, per_pair, out)
        if sync:
            self.sync()
        return plan

    def run(self, frames, dists, out, sync=True):
        """frames: uint8 [F, H, W, 3] (page-locked for real overlap); dists: dist of each interpolated frame
        inside a pair (Python floats, see interpolation_dists); out: uint8 [(F-1)*len(dists), H, W, 3], filled
        in pair order.  Returns after everything has been enqueued AND (sync=True) completed; with sync=False
        the caller streams several batches back to back and calls sync() once."""
        nd = len(dists)
        per_pair = {k: [(k * nd + j, d) for j, d in enumerate(dists)] for k in range(len(frames) - 1)}
        self._run_pairs(frames, per_pair, out)
        if sync:
            self.sync()
        return out

    def _run_pairs(self, frames, per_pair, out):
        """per_pair: {pair index k: [(output index, dist), ...]}; pairs without work are skipped.  Everything is
        only enqueued here; sync() waits.  The stager is a RING of 2 C + 4 device buffers: a source frame is uploaded
        right before the first context that needs it asks for it (in enqueue order), the H2D stream runs ahead as far
        as ring slots are free, and a slot is rewritten once its previous tenant has been copied out (the stager
        waits for those copies itself) -- so device memory does not grow with the length of the clip."""
        pairs = sorted(per_pair)
        n_pairs = len(pairs)
        if n_pairs == 0:
            return
        C = min(len(self.ctxs), n_pairs)
        n_ring = 2 * C + 4
        if self.stager is None or self.stager.n < n_ring:
            if self.stager is not None:
                self.sync()
                self.stager.close()
            self.stager = FrameStager(self.H, self.W, n_ring, self.device)
        st = self.stager
        n_ring = st.n
        bounds = [round(i * n_pairs / C) for i in range(C + 1)]
        # enqueue order: the runs advance in lock step
        order = [(c, bounds[c] + r) for r in range(max(bounds[c + 1] - bounds[c] for c in range(C))) for c in range(C)
                 if bounds[c] + r < bounds[c + 1]]
        ring_of = {}                                     # frame index -> ring slot of its current upload
        n_up = self._ring_cursor

        def staged(f):
            nonlocal n_up
            if f not in ring_of:
                slot = n_up % n_ring
                for g in [g for g, s_ in ring_of.items() if s_ == slot]:
                    del ring_of[g]                       # that frame's tenancy ends; a later taker uploads it again
                st.upload(slot, frames[f])
                ring_of[f] = slot
                n_up += 1
                self.uploads += 1
            return ring_of[f]

        prev = {}                                        # context -> pair it processed last
        for c, r in order:
            k = pairs[r]
            ctx = self.ctxs[c]
            a, b = k % 3, (k + 1) % 3
            if prev.get(c) != k - 1:
                ctx.frame_from_stager(a, st, staged(k))  # not the successor of this context's last pair: both frames are new to it
            ctx.frame_from_stager(b, st, staged(k + 1))
            prev[c] = k
            ctx.estimate_pair(self.params, a, b, 0, 1)
            for j, (oi, d) in enumerate(per_pair[k]):
                o = 6 + (j & 1)
                if self.bidir2:
                    # the class passes the python floats dist and 1 - dist (Bidirectional_2.py:331-332)
                    ctx.mci_bidir2(a, b, 0, 1, np.float32(d), np.float32(1 - d), self._w, o,
                                   self.params.min_block_size, 4 * self.params.min_block_size)
                elif self.bidir:
                    ctx.mci_bidir(a, b, 0, 1, d, o)
                else:
                    ctx.mci_unidir(a, b, 0, d, o)
                ctx.download_frame_via(o, out[oi], st)
        self._ring_cursor = n_up % n_ring
