"""Python handle on one libmemci_b200 context (one GPU, one stream, device slots)."""
import ctypes as C

import numpy as np

from . import _native as N

FILTER_MODES = {None: 0, "none": 0, "mean": 1, "median": 2, "weighted": 3}
NUM_FRAME_SLOTS = 8
NUM_MV_SLOTS = 8


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PinnedBuffer:
    """Page-locked host array (the decode-to-device staging buffer)."""

    def __init__(self, shape, dtype=np.uint8):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        N.check(None, N.lib().memci_host_alloc(nbytes, C.byref(p)))
        self._p = p
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def close(self):
        if self._p is not None:
            self.array = None
            N.lib().memci_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceContext:
    def __init__(self, H, W, device=0, stream=None):
        self.H, self.W, self.device = int(H), int(W), int(device)
        self._ctx = C.c_void_p()
        L = N.lib()
        n = L.memci_device_count()
        if n <= 0:
            raise N.MemciError("no CUDA device visible; the MEMCI path has no CPU fallback "
                               f"({(L.memci_last_error(None) or b'').decode()})")
        N.check(None, L.memci_ctx_create(self.device, self.H, self.W, C.c_void_p(stream or 0), C.byref(self._ctx)))
        self.L = L

    # -- lifetime
    def close(self):
        if self._ctx:
            self.L.memci_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        N.check(self._ctx, rc)

    def sync(self):
        self._ck(self.L.memci_sync(self._ctx))

    @property
    def stream(self):
        return self.L.memci_stream(self._ctx)

    # -- frames
    def _frame(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        if a.shape != (self.H, self.W, 3):
            raise ValueError(f"frame shape {a.shape} != {(self.H, self.W, 3)}")
        return a

    def upload_frame(self, slot, frame):
        a = self._frame(frame)
        self._ck(self.L.memci_upload_frame(self._ctx, slot, _ptr(a)))
        return a            # keep alive until the caller syncs

    def set_frame_device(self, slot, dev_ptr):
        self._ck(self.L.memci_set_frame_device(self._ctx, slot, C.c_void_p(dev_ptr)))

    def download_frame(self, slot, out=None, sync=True):
        if out is None:
            out = np.empty((self.H, self.W, 3), np.uint8)
        self._ck(self.L.memci_download_frame(self._ctx, slot, _ptr(out)))
        if sync:
            self.sync()
        return out

    def frame_from_stager(self, slot, stager, ring_slot):
        self._ck(self.L.memci_frame_from_stager(self._ctx, slot, stager._st, int(ring_slot)))

    def download_frame_via(self, slot, out, stager):
        """Async D2H through the stager's copy stream; stager.sync() before reading `out`."""
        self._ck(self.L.memci_download_frame_via(self._ctx, slot, _ptr(out), stager._st))

    def frame_ptr(self, slot):
        return self.L.memci_frame_device_ptr(self._ctx, slot)

    # -- motion estimation
    def me_fs(self, src, tgt, block_size, target_region, mv_slot):
        self._ck(self.L.memci_me_fs(self._ctx, src, tgt, int(block_size), int(target_region), mv_slot))

    def me_fs_pair(self, a, b, block_size, target_region, mv_fwd, mv_bwd):
        """fwd = fs(a -> b) and bwd = fs(b -> a) in one launch (same results as two me_fs calls)."""
        self._ck(self.L.memci_me_fs_pair(self._ctx, a, b, int(block_size), int(target_region), mv_fwd, mv_bwd))

    def me_tss(self, src, tgt, block_size, steps, mv_slot):
        self._ck(self.L.memci_me_tss(self._ctx, src, tgt, int(block_size), int(steps), mv_slot))

    def me_hbma(self, src, tgt, block_size, win, sub, steps, min_bs, mv_slot, cost_key=0):
        self._ck(self.L.memci_me_hbma(self._ctx, src, tgt, int(block_size), int(win), int(sub), int(steps),
                                      int(min_bs), int(cost_key), mv_slot))

    def pyramid_level(self, slot, level):
        h, w = self.H, self.W
        for _ in range(level):
            h, w = (h + 1) // 2, (w + 1) // 2
        out = np.empty((h, w, 3), np.uint8)
        self._ck(self.L.memci_download_pyramid_level(self._ctx, slot, level, _ptr(out)))
        self.sync()
        return out

    def smooth(self, mv_in, mode, filter_size, mv_out=None):
        self._ck(self.L.memci_smooth(self._ctx, mv_in, FILTER_MODES[mode] if not isinstance(mode, int) else mode,
                                     int(filter_size), mv_in if mv_out is None else mv_out))

    def estimate_pair(self, params, src, tgt, mv_fwd, mv_bwd):
        self._ck(self.L.memci_estimate_pair(self._ctx, C.byref(params), src, tgt, mv_fwd, mv_bwd))

    def mv_info(self, slot):
        v = [C.c_int() for _ in range(6)]
        self._ck(self.L.memci_mv_info(self._ctx, slot, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def download_mv_blocks(self, slot):
        mc, mr, mcol, sc, sr, scol = self.mv_info(slot)
        vec = np.empty((mr, mcol, 2), np.float32)
        sad = np.empty((sr, scol), np.float32)
        self._ck(self.L.memci_download_mv_blocks(self._ctx, slot, _ptr(vec), _ptr(sad)))
        self.sync()
        return mc, vec, sc, sad

    def download_mv_dense(self, slot):
        out = np.empty((self.H, self.W, 3), np.float32)
        self._ck(self.L.memci_download_mv_dense(self._ctx, slot, _ptr(out)))
        self.sync()
        return out

    def upload_mv_dense(self, slot, mv):
        a = np.ascontiguousarray(mv, dtype=np.float32)
        if a.shape != (self.H, self.W, 3):
            raise ValueError(f"MV field shape {a.shape} != {(self.H, self.W, 3)}")
        self._ck(self.L.memci_upload_mv_dense(self._ctx, slot, _ptr(a)))
        self.sync()

    def upload_mv_blocks(self, slot, mv_cell, vec, sad_cell, sad):
        v = np.ascontiguousarray(vec, dtype=np.float32)
        s = np.ascontiguousarray(sad, dtype=np.float32)
        self._ck(self.L.memci_upload_mv_blocks(self._ctx, slot, int(mv_cell), _ptr(v), int(sad_cell), _ptr(s)))
        self.sync()

    # -- MCI
    def mci_unidir(self, src, tgt, mv, dist, out_slot):
        self._ck(self.L.memci_mci_unidir(self._ctx, src, tgt, mv, C.c_float(float(np.float32(dist))), out_slot))

    def mci_bidir(self, src, tgt, mv_fwd, mv_bwd, dist, out_slot):
        self._ck(self.L.memci_mci_bidir(self._ctx, src, tgt, mv_fwd, mv_bwd, C.c_float(float(np.float32(dist))), out_slot))

    def mci_bidir2(self, src, tgt, mv_fwd, mv_bwd, dist_fwd, dist_bwd, weights, out_slot, min_block_size=4, pad_size=16):
        """IEWMC both ways + error-adaptive combination + BDHI + unpad (Bidirectional_2.py:330-345)."""
        w = np.ascontiguousarray(weights, np.float32)
        assert w.size == (2 * min_block_size) ** 2
        self._ck(self.L.memci_mci_bidir2(self._ctx, src, tgt, mv_fwd, mv_bwd, C.c_float(float(np.float32(dist_fwd))),
                                         C.c_float(float(np.float32(dist_bwd))), int(min_block_size), int(pad_size),
                                         w.ctypes.data_as(C.c_void_p), out_slot))

    def bidir2_stats(self):
        a, b = C.c_int(), C.c_int()
        self._ck(self.L.memci_bidir2_stats(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def bidir2_out_of_range_total(self, reset=True):
        """Blocks shifted outside the padded frame over every bidir2 call since the last reset (synchronises)."""
        n = C.c_int()
        self._ck(self.L.memci_bidir2_out_of_range_total(self._ctx, C.byref(n), 1 if reset else 0))
        return n.value

    # -- frame quality (Middlebury harness)
    def quality(self, slot_a, slot_b):
        """(mean squared error, mean SSIM) of two resident frames; see memci_quality in include/memci_b200.h."""
        mse, ssim = C.c_double(), C.c_double()
        self._ck(self.L.memci_quality(self._ctx, slot_a, slot_b, C.byref(mse), C.byref(ssim)))
        return mse.value, ssim.value

    # -- instrumentation
    def launch_count(self):
        return int(self.L.memci_launch_count(self._ctx))

    def reset_launch_count(self):
        self.L.memci_reset_launch_count(self._ctx)

    def fs_stats(self):
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_int()
        self._ck(self.L.memci_fs_stats(self._ctx, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def fs_prefilter_stats(self):
        """(survivors of the 8-bit prefilter, list overflowed?) of the last full search."""
        a, b = C.c_int(), C.c_int()
        self._ck(self.L.memci_fs_prefilter_stats(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def mci_stats(self):
        n = C.c_int()
        self._ck(self.L.memci_mci_stats(self._ctx, C.byref(n)))
        return n.value

    def enable_kernel_timing(self, on=True):
        self._ck(self.L.memci_enable_kernel_timing(self._ctx, 1 if on else 0))

    def last_kernel_ms(self):
        a, b = C.c_float(), C.c_float()
        self._ck(self.L.memci_last_kernel_ms(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def profile_start(self, capacity):
        self._ck(self.L.memci_profile_start(self._ctx, int(capacity)))

    def profile_collect(self):
        """(fs_ms_sum, fs_launches, mci_ms_sum, mci_launches) since profile_start()."""
        a, b = C.c_double(), C.c_double()
        na, nb = C.c_int(), C.c_int()
        self._ck(self.L.memci_profile_collect(self._ctx, C.byref(a), C.byref(na), C.byref(b), C.byref(nb)))
        return a.value, na.value, b.value, nb.value

    PROF_KINDS = ("fs", "mci", "hbma", "hbm_small", "refine", "holefill", "prep", "bidir2")

    def profile_collect_kinds(self):
        """{kind: (ms_sum, launches, algorithmic work)} since profile_start(); kinds as MEMCI_PROF_* in the header."""
        n = len(self.PROF_KINDS)
        ms, cnt, work = (C.c_double * n)(), (C.c_int * n)(), (C.c_double * n)()
        self._ck(self.L.memci_profile_collect_kinds(self._ctx, ms, cnt, work))
        return {k: (ms[i], cnt[i], work[i]) for i, k in enumerate(self.PROF_KINDS)}

    def microbench(self, mode, iters=4096):
        """(ms, px_ops) of one launch of the issue-rate microbenchmark (roofline denominator)."""
        ms, ops = C.c_float(), C.c_double()
        self._ck(self.L.memci_microbench(self._ctx, int(mode), int(iters), C.byref(ms), C.byref(ops)))
        return ms.value, ops.value


_CTX_CACHE = {}


def get_context(H, W, device=0):
    """Shared context for the function-level API (one per (device, H, W))."""
    key = (int(device), int(H), int(W))
    ctx = _CTX_CACHE.get(key)
    if ctx is None:
        ctx = _CTX_CACHE[key] = DeviceContext(H, W, device)
    return ctx


def clear_contexts():
    for c in list(_CTX_CACHE.values()):
        c.close()
    _CTX_CACHE.clear()


class FrameStager:
    """Ring of device frame buffers with H2D / D2H streams of their own (include/memci_b200.h, frame staging):
    a source frame crosses PCIe once and is handed to the contexts device-to-device."""

    def __init__(self, H, W, n_slots, device=0):
        self.L = N.lib()
        self.H, self.W, self.n = int(H), int(W), int(n_slots)
        st = C.c_void_p()
        rc = self.L.memci_stager_create(int(device), self.H, self.W, self.n, C.byref(st))
        if rc != 0:
            raise N.MemciError(f"memci_stager_create failed ({rc})")
        self._st = st

    def upload(self, ring_slot, frame):
        a = np.ascontiguousarray(frame, dtype=np.uint8)
        if a.shape != (self.H, self.W, 3):
            raise ValueError(f"frame shape {a.shape} != {(self.H, self.W, 3)}")
        rc = self.L.memci_stager_upload(self._st, int(ring_slot), _ptr(a))
        if rc != 0:
            raise N.MemciError(f"memci_stager_upload failed ({rc})")
        return a

    def sync(self):
        rc = self.L.memci_stager_sync(self._st)
        if rc != 0:
            raise N.MemciError(f"memci_stager_sync failed ({rc})")

    def close(self):
        if getattr(self, "_st", None):
            self.L.memci_stager_destroy(self._st)
            self._st = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
