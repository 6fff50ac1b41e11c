"""One frame pair split into horizontal bands over several GPUs (8K mode).

Every rank owns a band of block rows of BOTH frames.  What a band needs from its neighbours:

  * frame rows within  R + 10 (+ one block)  of the band: the full-search window reaches R rows
    beyond the band's blocks, the splat of an output row gathers source pixels up to R rows away
    (|round(mv * dist)| <= R), and the 21x21 hole fill looks 10 rows further;
  * motion-vector block rows covering the same source rows (computed by the neighbour);
  * the wrap link: a splat target above row 0 wraps to the bottom of the frame (numba negative
    indexing), so bands touching the last R + 10 rows also need frame rows [0, R) and the
    MV block rows above them from the top band.

So there are two halo exchanges per pair -- frame rows before motion estimation, MV rows after it
-- each a handful of point-to-point copies of < 1 MB (NCCL send/recv over NVLink between ranks;
device-to-device copies when several bands are emulated on one GPU for testing).  Nothing else is
communicated, and the assembled output is bit-identical to the single-GPU result.
"""
import numpy as np

from .sharding import band_split

HOLE_K = 10     # half width of the reference's median hole-fill window (Unidirectional.py:62)


def _isect(a, b):
    lo, hi = max(a[0], b[0]), min(a[1], b[1])
    return (lo, hi) if lo < hi else None


class BandPlan:
    """Row bookkeeping of the band split.  Pure Python: testable without a GPU."""

    def __init__(self, H, W, block_size, target_region, n_bands):
        self.H, self.W, self.bs, self.R, self.n = int(H), int(W), int(block_size), int(target_region), int(n_bands)
        self.nbr = -(-self.H // self.bs)
        self.block_rows = band_split(self.nbr, self.n)                       # [b0, b1) per band
        self.rows = [(b0 * self.bs, min(self.H, b1 * self.bs)) for b0, b1 in self.block_rows]
        reach = self.R + HOLE_K
        self.frame_need, self.mv_need = [], []
        for (y0, y1) in self.rows:
            need = [(max(0, y0 - reach), min(self.H, y1 + reach + self.bs))]
            mv = [(max(0, y0 - reach) // self.bs, -(-min(self.H, y1 + reach) // self.bs))]
            if y1 + reach > self.H - 1 and self.R > 0:                       # wrap link from the top of the frame
                need.append((0, min(self.H, self.R)))
                mv.append((0, -(-min(self.H, self.R) // self.bs)))
            self.frame_need.append(need)
            self.mv_need.append(mv)

    def transfers(self, kind):
        """[(src_band, dst_band, row0, row1)]: rows (frame rows or MV block rows) src owns and dst needs."""
        own = self.rows if kind == "frame" else self.block_rows
        need = self.frame_need if kind == "frame" else self.mv_need
        out = []
        for dst in range(self.n):
            got = []
            for span in need[dst]:
                for src in range(self.n):
                    if src == dst:
                        continue
                    x = _isect(span, own[src])
                    if x and x not in got:
                        got.append(x)
                        out.append((src, dst, x[0], x[1]))
        return out

    def bytes_moved(self):
        f = sum((r1 - r0) * self.W * 3 * 2 for _, _, r0, r1 in self.transfers("frame"))
        nbc = -(-self.W // self.bs)
        m = sum((r1 - r0) * nbc * 12 for _, _, r0, r1 in self.transfers("mv"))
        return f, m


def device_tensor(ptr, nbytes, device_index):
    """uint8 torch view of library-owned device memory (no copy) via __cuda_array_interface__."""
    import torch

    class _Iface:
        __cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_Iface(), device=torch.device("cuda", device_index))


class BandWorker:
    """One band on one DeviceContext.  Frame slots 0/1 = the pair, MV slots 0/1 = fwd/bwd, out slot 7."""

    def __init__(self, ctx, plan, band, device_index=0):
        import ctypes as C
        self.ctx, self.plan, self.band, self.dev = ctx, plan, band, device_index
        H, W, bs = plan.H, plan.W, plan.bs
        self.frame_pitch = W * 3
        nbc = -(-W // bs)
        L = ctx.L
        for s in (0, 1):
            ctx._ck(L.memci_mv_alloc(ctx._ctx, s, bs))
        self.frames = [device_tensor(ctx.frame_ptr(s), H * W * 3, device_index) for s in (0, 1)]
        self.mv_vec, self.mv_sad = [], []
        for s in (0, 1):
            v, a = C.c_void_p(), C.c_void_p()
            ctx._ck(L.memci_mv_device_ptrs(ctx._ctx, s, C.byref(v), C.byref(a)))
            self.mv_vec.append(device_tensor(v.value, plan.nbr * nbc * 8, device_index))
            self.mv_sad.append(device_tensor(a.value, plan.nbr * nbc * 4, device_index))
            self.mv_vec[-1].zero_()          # rows never received stay at zero motion: they cannot reach this band
            self.mv_sad[-1].zero_()
        self.vec_pitch, self.sad_pitch = nbc * 8, nbc * 4

    # -- the four local phases -----------------------------------------------------------------
    def upload(self, frame_a, frame_b):
        y0, y1 = self.plan.rows[self.band]
        L, c = self.ctx.L, self.ctx
        for s, f in ((0, frame_a), (1, frame_b)):
            f = np.ascontiguousarray(f, dtype=np.uint8)
            c._ck(L.memci_upload_frame_rows(c._ctx, s, f.ctypes.data, y0, y1))

    def frames_ready(self):
        for s in (0, 1):
            self.ctx._ck(self.ctx.L.memci_frame_mark(self.ctx._ctx, s))

    def estimate(self):
        b0, b1 = self.plan.block_rows[self.band]
        L, c, p = self.ctx.L, self.ctx, self.plan
        c._ck(L.memci_me_fs_rows(c._ctx, 0, 1, p.bs, p.R, 0, b0, b1))
        c._ck(L.memci_me_fs_rows(c._ctx, 1, 0, p.bs, p.R, 1, b0, b1))

    def interpolate(self, dist, bidir=True):
        import ctypes as C
        y0, y1 = self.plan.rows[self.band]
        c = self.ctx
        c._ck(c.L.memci_mci_rows(c._ctx, 0, 1, 0, 1, C.c_float(float(np.float32(dist))), 7, 1 if bidir else 0, y0, y1))

    def download(self, out):
        y0, y1 = self.plan.rows[self.band]
        c = self.ctx
        c._ck(c.L.memci_download_frame_rows(c._ctx, 7, out.ctypes.data, y0, y1))

    # -- views used by the exchanges ---------------------------------------------------------------
    def views(self, kind, r0, r1):
        if kind == "frame":
            return [t[r0 * self.frame_pitch:r1 * self.frame_pitch] for t in self.frames]
        return ([t[r0 * self.vec_pitch:r1 * self.vec_pitch] for t in self.mv_vec] +
                [t[r0 * self.sad_pitch:r1 * self.sad_pitch] for t in self.mv_sad])


def exchange_local(workers, kind):
    """Emulation of the halo exchange between bands that live on ONE GPU (tests): the same transfer
    list, executed as device-to-device copies."""
    import torch
    for w in workers:
        w.ctx.sync()
    for src, dst, r0, r1 in workers[0].plan.transfers(kind):
        for a, b in zip(workers[src].views(kind, r0, r1), workers[dst].views(kind, r0, r1)):
            b.copy_(a)
    torch.cuda.synchronize()


def exchange_distributed(worker, kind):
    """Halo exchange between ranks: one NCCL send/recv per (neighbour, buffer), batched.  Must run on
    the stream the worker's context enqueues on (create the context on torch's current stream)."""
    import torch.distributed as dist
    ops = []
    me = worker.band
    for src, dst, r0, r1 in worker.plan.transfers(kind):
        if src == me:
            ops += [dist.P2POp(dist.isend, v, dst) for v in worker.views(kind, r0, r1)]
        elif dst == me:
            ops += [dist.P2POp(dist.irecv, v, src) for v in worker.views(kind, r0, r1)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def interpolate_pair_emulated(frame_a, frame_b, block_size, target_region, dist, n_bands, bidir=True, device=0):
    """All bands of one pair on one GPU, one context per band (test / reference for the N-GPU run)."""
    from .device import DeviceContext
    H, W, _ = frame_a.shape
    plan = BandPlan(H, W, block_size, target_region, n_bands)
    ctxs = [DeviceContext(H, W, device) for _ in range(n_bands)]
    try:
        workers = [BandWorker(c, plan, i, device) for i, c in enumerate(ctxs)]
        for w in workers:
            w.upload(frame_a, frame_b)
        exchange_local(workers, "frame")
        for w in workers:
            w.frames_ready()
            w.estimate()
        exchange_local(workers, "mv")
        out = np.empty_like(frame_a)
        for w in workers:
            w.interpolate(dist, bidir)
            w.download(out)
        for c in ctxs:
            c.sync()
        return out, plan
    finally:
        for c in ctxs:
            c.close()
