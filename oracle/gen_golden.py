"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python oracle/gen_golden.py

The reference ships no tests or golden vectors, so these fixtures are
outputs of the reference's own functions / classes on seeded inputs, produced with
numba 0.65.0 / NumPy 2.3.5 / SciPy 1.18.1.  Large outputs are stored as SHA-256 digests.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_harness  # noqa: E402
from interpolatory_b200.synth import make_sequence  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest() + f":{a.dtype}:{'x'.join(map(str, a.shape))}"


def blocks_of(dense, bs):
    return np.ascontiguousarray(dense[::bs, ::bs])


def gen_tss(ns):
    """Three-step search (ME/tss.py): function level + the class API with me_mode=tss."""
    t = {}
    for (H, W) in [(72, 96), (50, 70), (37, 53), (120, 40)]:
        fr = make_sequence(H, W, 2, seed=H)
        k = f"{H}x{W}"
        for bs, steps in ((8, 3), (8, 1), (4, 2), (16, 4), (8, 5)):
            t[f"{k}/tss/{bs}_{steps}"] = blocks_of(ns.tss.get_motion_vectors(bs, steps, fr[0], fr[1]), bs)

    class SynthStream(ns.BaseVideoStream):
        def __init__(self, frames, fps):
            self.frames, self.nframes, self.fps, self.duration = frames, len(frames), fps, len(frames) / fps

        def get_frame(self, idx):
            return self.frames[idx]

    frames = make_sequence(72, 96, 3, seed=2072)
    for si, st in enumerate([dict(me_mode='tss', mci_mode='unidir'),
                             dict(me_mode='tss', mci_mode='bidir', filter_mode='mean', block_size='8', target_region='9')]):
        it = ns.MEMCI(60, **dict(st))
        it.video_stream = SynthStream(list(frames), 24)
        it.rate_ratio = 60 / 24
        t[f"e2e/set{si}/settings"] = np.array(repr(sorted(st.items())))
        t[f"e2e/set{si}/out"] = np.stack([np.asarray(it.get_interpolated_frame(i)) for i in range(5)])
    np.savez_compressed(os.path.join(OUT, "tss_small.npz"), **t)


def gen_bidir2(ns):
    """bidir2 (MCI/Bidirectional_2.py): function level (IEWMC / EAC / BDHI on fs and hbma vectors, with the
    vectors stored so that the CPU tests need no ME) + the class API for the three me_modes."""
    from src.Interpolators.MEMCI.MCI import Bidirectional_2 as b2
    t = {}
    t["weight_kern"] = b2.get_weight_kern(8)[:, :, 0].copy()
    for (H, W) in [(72, 96), (50, 70), (37, 53)]:
        fr = make_sequence(H, W, 2, seed=H + 1000)
        a, b = fr[0], fr[1]
        k = f"{H}x{W}"
        fields = {"fs": (ns.fs.get_motion_vectors(8, 7, a, b), ns.fs.get_motion_vectors(8, 7, b, a), True),
                  "hbma": (ns.hbma.get_motion_vectors(8, 7, 1, 2, 4, a, b, 'sad', False),
                           ns.hbma.get_motion_vectors(8, 7, 1, 2, 4, b, a, 'sad', False), False)}
        for me, (fwd, bwd, up) in fields.items():
            t[f"{k}/{me}/fwd"] = blocks_of(fwd, 8) if up else fwd
            t[f"{k}/{me}/bwd"] = blocks_of(bwd, 8) if up else bwd
            for di, dist in enumerate((0.5, 0.39999998, 0.8)):
                Ff, Ef = b2.IEWMC(fwd, a, b, dist, 16, 4, up)
                Fb, Eb = b2.IEWMC(bwd, b, a, 1 - dist, 16, 4, up)
                Fc = b2.Error_adaptive_combination(Ff, Ef, Fb, Eb)
                out = b2.unpad_and_downconvert(b2.BDHI(Fc, 4, 16), 16)
                if di == 0 and H == 37:          # one full set of intermediates (small)
                    t[f"{k}/{me}/Ff"], t[f"{k}/{me}/Ef"], t[f"{k}/{me}/Fc"] = Ff, Ef, Fc
                t[f"{k}/{me}/out{di}"] = out
    # fill_holes on random blocks (incl. an all-hole block -> NaN)
    rng = np.random.default_rng(42)
    blocks = []
    for i in range(24):
        blk = rng.uniform(0, 255, (8, 8, 3)).astype(np.float32)
        blk[rng.random((8, 8)) < (0.05, 0.3, 0.7)[i % 3]] = -1
        blocks.append(blk)
    blocks.append(np.full((8, 8, 3), -1, np.float32))
    t["fill_holes/in"] = np.stack(blocks)
    t["fill_holes/out"] = np.stack([b2.fill_holes(b.copy()) for b in blocks])

    class SynthStream(ns.BaseVideoStream):
        def __init__(self, frames, fps):
            self.frames, self.nframes, self.fps, self.duration = frames, len(frames), fps, len(frames) / fps

        def get_frame(self, idx):
            return self.frames[idx]

    frames = make_sequence(72, 96, 3, seed=2272)
    for si, st in enumerate([dict(mci_mode='bidir2'),
                             dict(mci_mode='bidir2', me_mode='fs', block_size='8', target_region='7'),
                             dict(mci_mode='bidir2', me_mode='tss'),
                             dict(mci_mode='bidir2', me_mode='fs', block_size='16', target_region='5', filter_mode='median')]):
        it = ns.MEMCI(60, **dict(st))
        it.video_stream = SynthStream(list(frames), 24)
        it.rate_ratio = 60 / 24
        t[f"e2e/set{si}/settings"] = np.array(repr(sorted(st.items())))
        t[f"e2e/set{si}/out"] = np.stack([np.asarray(it.get_interpolated_frame(i)) for i in range(5)])
    np.savez_compressed(os.path.join(OUT, "bidir2_small.npz"), **t)


def main():
    ns = ref_harness.load()
    os.makedirs(OUT, exist_ok=True)
    if "--only-tss" in sys.argv:
        return gen_tss(ns)
    if "--only-bidir2" in sys.argv:
        return gen_bidir2(ns)
    gen_tss(ns)
    gen_bidir2(ns)
    rng = np.random.default_rng(7)

    # ---- ratio / dist table (src/util.py:49-60) ---------------------------------
    rows = []
    for rr in (2.0, 2.5, 5.0, 60 / 23.976, 60 / 29.97, 144 / 25, 7 / 3):
        for i in range(40):
            fi, ratio = ns.get_first_frame_idx_and_ratio(i, rr)
            rows.append((rr, i, fi, ratio))
    np.savez_compressed(os.path.join(OUT, "ratio.npz"), table=np.array(rows, np.float64))

    # ---- ME / smoothing / MCI function-level, small shapes ----------------------
    d = {}
    shapes = [(72, 96), (50, 70), (37, 53)]
    fs_cfgs = [(8, 7), (8, 3), (4, 5), (16, 9), (8, 16)]
    hb_cfgs = [(8, 7, 1, 2, 4), (8, 3, 1, 1, 4), (8, 5, 2, 3, 2), (4, 2, 1, 1, 4), (16, 4, 1, 1, 4)]
    dists = [0.5, 0.39999998, 0.8, 0.19999999, 0.6]
    for (H, W) in shapes:
        fr = make_sequence(H, W, 2, seed=H)
        a, b = fr[0], fr[1]
        k = f"{H}x{W}"
        d[f"{k}/a"] = a
        d[f"{k}/b"] = b
        for bs, R in fs_cfgs:
            mv = ns.fs.get_motion_vectors(bs, R, a, b)
            d[f"{k}/fs/{bs}_{R}"] = blocks_of(mv, bs)
        d[f"{k}/fs_dense/8_7"] = ns.fs.get_motion_vectors(8, 7, a, b)
        for cfg in hb_cfgs:
            mv = ns.hbma.get_motion_vectors(*cfg, a, b, 'sad', False)
            d[f"{k}/hbma/{'_'.join(map(str, cfg))}"] = mv
        d[f"{k}/hbma_dense/8_7_1_2_4"] = ns.hbma.get_motion_vectors(8, 7, 1, 2, 4, a, b)
        lvl = a
        for s in range(1, 3):
            from scipy.ndimage import convolve
            wts = np.array([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]])[:, :, None]
            lvl = (convolve(lvl / 255.0, wts, mode='constant')[::2, ::2] * 255.0).astype(np.uint8)
            d[f"{k}/pyr/{s}"] = lvl
        fwd = ns.fs.get_motion_vectors(8, 7, a, b)
        bwd = ns.fs.get_motion_vectors(8, 7, b, a)
        fwh = ns.hbma.get_motion_vectors(8, 7, 1, 2, 4, a, b)
        bwh = ns.hbma.get_motion_vectors(8, 7, 1, 2, 4, b, a)
        for mode in ("mean", "median", "weighted"):
            for fsz in (4, 8, 3):
                d[f"{k}/smooth/{mode}_{fsz}/fs"] = np.asarray(ns.smooth(ns.filters[mode], fwd, fsz), np.float32)[:, :, :2].astype(np.float32)
                d[f"{k}/smooth/{mode}_{fsz}/hbma"] = np.asarray(ns.smooth(ns.filters[mode], fwh, fsz), np.float32)[:, :, :2].astype(np.float32)
        for dist in dists:
            dd = np.float32(dist)
            d[f"{k}/unidir/fs/{dist}"] = ns.uni.helper(a, b, fwd, dd)
            d[f"{k}/bidir/fs/{dist}"] = ns.bi.helper(a, b, fwd, bwd, dd)
            d[f"{k}/unidir/hbma/{dist}"] = ns.uni.helper(a, b, fwh, dd)
            d[f"{k}/bidir/hbma/{dist}"] = ns.bi.helper(a, b, fwh, bwh, dd)
        # adversarial random block-constant fields (wraps, OOB, fractional MVs, many overlaps)
        for trial in range(3):
            g = (4, 8, 2)[trial]
            flds = []
            for _ in range(2):
                blk = rng.integers(-12, 13, size=(-(-H // g), -(-W // g), 2)).astype(np.float32)
                if trial == 1:
                    blk += rng.integers(0, 4, size=blk.shape).astype(np.float32) / 4
                sad = rng.integers(0, 40, size=(-(-H // g), -(-W // g), 1)).astype(np.float32)
                flds.append(np.concatenate([blk, sad], axis=2))
            d[f"{k}/rand/{trial}/g"] = np.int32(g)
            d[f"{k}/rand/{trial}/fwd"] = flds[0]
            d[f"{k}/rand/{trial}/bwd"] = flds[1]
            f_ = np.kron(flds[0], np.ones((g, g, 1), np.float32))[:H, :W].copy()
            b_ = np.kron(flds[1], np.ones((g, g, 1), np.float32))[:H, :W].copy()
            for dist in (0.5, 0.39999998, 0.8):
                dd = np.float32(dist)
                d[f"{k}/rand/{trial}/unidir/{dist}"] = ns.uni.helper(a, b, f_, dd)
                d[f"{k}/rand/{trial}/bidir/{dist}"] = ns.bi.helper(a, b, f_, b_, dd)
    np.savez_compressed(os.path.join(OUT, "functions_small.npz"), **d)

    # ---- class API end to end (MEMCI factory, MV cache, dist cycle) -------------
    class SynthStream(ns.BaseVideoStream):
        def __init__(self, frames, fps):
            self.frames = frames
            self.nframes = len(frames)
            self.fps = fps
            self.duration = len(frames) / fps
            self.size = (frames[0].shape[1], frames[0].shape[0])

        def get_frame(self, idx):
            return self.frames[idx]

    e = {}
    settings = [
        dict(me_mode='fs', block_size='8', target_region='7', mci_mode='bidir'),
        dict(me_mode='fs', block_size='8', target_region='7', mci_mode='unidir'),
        dict(mci_mode='unidir'),
        dict(mci_mode='bidir', filter_mode='mean'),
        dict(me_mode='fs', mci_mode='bidir', filter_mode='median', filter_size='4', target_region='5'),
        dict(me_mode='fs', mci_mode='unidir', filter_mode='weighted', filter_size='8'),
        dict(me_mode='hbma', mci_mode='bidir', filter_mode='weighted', block_size='8', target_region='4'),
    ]
    for (H, W, fps_in, fps_out, nsrc) in ((72, 96, 24, 60, 4), (56, 88, 30, 60, 3), (264, 352, 24, 120, 2)):
        frames = make_sequence(H, W, nsrc, seed=1000 + H)
        key = f"{H}x{W}_{fps_in}_{fps_out}"
        if H < 100:
            e[f"{key}/frames"] = frames
        e[f"{key}/frames_sha"] = np.array(sha(frames))
        e[f"{key}/seed"] = np.int32(1000 + H)
        nout = int((nsrc - 1) * fps_out / fps_in)
        for si, st in enumerate(settings):
            if H >= 100 and si not in (0, 3, 6):
                continue
            it = ns.MEMCI(fps_out, **dict(st))
            it.video_stream = SynthStream(list(frames), fps_in)
            it.rate_ratio = fps_out / fps_in
            outs = np.stack([np.asarray(it.get_interpolated_frame(i)) for i in range(nout)])
            e[f"{key}/set{si}/settings"] = np.array(repr(sorted(st.items())))
            if H < 100:
                e[f"{key}/set{si}/out"] = outs
            e[f"{key}/set{si}/out_sha"] = np.array([sha(o) for o in outs])
    np.savez_compressed(os.path.join(OUT, "e2e.npz"), **e)

    # ---- mid-size cfg-1 pair (360x640, fs 8/7) : digests + block MVs -------------
    m = {}
    fr = make_sequence(360, 640, 2, seed=0)
    a, b = fr[0], fr[1]
    m["frames_sha"] = np.array(sha(fr))
    fwd = ns.fs.get_motion_vectors(8, 7, a, b)
    bwd = ns.fs.get_motion_vectors(8, 7, b, a)
    m["fs_fwd"] = blocks_of(fwd, 8)
    m["fs_bwd"] = blocks_of(bwd, 8)
    fwh = ns.hbma.get_motion_vectors(8, 7, 1, 2, 4, a, b)
    bwh = ns.hbma.get_motion_vectors(8, 7, 1, 2, 4, b, a)
    m["hbma_fwd"] = blocks_of(fwh, 4)
    m["hbma_bwd"] = blocks_of(bwh, 4)
    for dist in (0.5, 0.39999998):
        dd = np.float32(dist)
        m[f"bidir_fs_sha/{dist}"] = np.array(sha(ns.bi.helper(a, b, fwd, bwd, dd)))
        m[f"unidir_fs_sha/{dist}"] = np.array(sha(ns.uni.helper(a, b, fwd, dd)))
        m[f"bidir_hbma_sha/{dist}"] = np.array(sha(ns.bi.helper(a, b, fwh, bwh, dd)))
    np.savez_compressed(os.path.join(OUT, "cfg1_360p.npz"), **m)

    # ---- two Middlebury crops (real images; sizes not multiples of 8) -----------
    try:
        from PIL import Image
        base = "/root/reference/Simulator/python/benchmarks/middlebury"
        r = {}
        for name, (y0, x0, h, w) in (("RubberWhale", (100, 200, 100, 132)), ("Dimetrodon", (150, 120, 61, 90))):
            a = np.asarray(Image.open(f"{base}/{name}/frame10.png").convert("RGB"))[y0:y0 + h, x0:x0 + w].copy()
            b = np.asarray(Image.open(f"{base}/{name}/frame11.png").convert("RGB"))[y0:y0 + h, x0:x0 + w].copy()
            r[f"{name}/a"] = a
            r[f"{name}/b"] = b
            fwd = ns.fs.get_motion_vectors(8, 7, a, b)
            bwd = ns.fs.get_motion_vectors(8, 7, b, a)
            r[f"{name}/fs_fwd"] = blocks_of(fwd, 8)
            r[f"{name}/fs_bwd"] = blocks_of(bwd, 8)
            r[f"{name}/hbma_fwd"] = ns.hbma.get_motion_vectors(8, 7, 1, 1, 4, a, b, 'sad', False)
            r[f"{name}/bidir"] = ns.bi.helper(a, b, fwd, bwd, np.float32(0.5))
            r[f"{name}/unidir"] = ns.uni.helper(a, b, fwd, np.float32(0.5))
        np.savez_compressed(os.path.join(OUT, "middlebury_crops.npz"), **r)
    except Exception as ex:  # pragma: no cover
        print("middlebury crops skipped:", ex)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
