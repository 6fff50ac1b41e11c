"""Bit-for-bit check of the C oracle against the UNMODIFIED reference, run in the build
container (needs /root/reference + numba).  Not part of the pytest suite: the committed
fixtures under tests/golden/ (made by gen_golden.py) carry the result to the GPU box."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import oracle as O          # noqa: E402
import ref_harness          # noqa: E402
from interpolatory_b200.synth import make_sequence  # noqa: E402


def check(name, a, b):
    ok = a.shape == b.shape and np.array_equal(a, b)
    nbad = -1 if a.shape != b.shape else int(np.count_nonzero(a != b))
    print(f"{'OK  ' if ok else 'FAIL'} {name}  mismatches={nbad}")
    return ok


def main():
    ns = ref_harness.load()
    allok = True
    rng = np.random.default_rng(1)
    shapes = [(72, 96), (50, 70), (37, 53)]
    for (H, W) in shapes:
        fr = make_sequence(H, W, 3, seed=H)
        a, b = fr[0], fr[1]
        for bs, R in ((8, 7), (8, 3), (4, 5), (16, 9)):
            t = time.time(); ref = ns.fs.get_motion_vectors(bs, R, a, b); tr = time.time() - t
            t = time.time(); mine = O.fs(bs, R, a, b); tm = time.time() - t
            allok &= check(f"fs {H}x{W} bs{bs} R{R} (ref {tr:.2f}s, oracle {tm:.3f}s)", ref, mine)
        # pyramid
        from scipy.ndimage import convolve
        wts = np.array([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]])[:, :, None]
        refp = (convolve(a / 255.0, wts, mode='constant')[::2, ::2] * 255.0).astype(np.uint8)
        allok &= check(f"pyr {H}x{W}", refp, O.pyr_down(a))
        for (bs, win, sub, steps, mbs) in ((8, 7, 1, 2, 4), (8, 3, 1, 1, 4), (8, 5, 2, 3, 2), (4, 2, 1, 1, 4)):
            for up in (True, False):
                ref = ns.hbma.get_motion_vectors(bs, win, sub, steps, mbs, a, b, 'sad', up)
                mine = O.hbma(bs, win, sub, steps, mbs, a, b, 'sad', up)
                allok &= check(f"hbma {H}x{W} bs{bs} win{win} sub{sub} steps{steps} min{mbs} up{up}", ref, mine)
        ref = ns.hbma.get_motion_vectors(8, 4, 1, 2, 4, a, b, 'ssd', True)
        allok &= check(f"hbma ssd {H}x{W}", ref, O.hbma(8, 4, 1, 2, 4, a, b, 'ssd', True))
        # smoothing
        mvf = ns.fs.get_motion_vectors(8, 7, a, b)
        mvh = ns.hbma.get_motion_vectors(8, 7, 1, 2, 4, a, b)
        for mode in ("mean", "median", "weighted"):
            for fsz in (4, 8, 3):
                for nm, mv in (("fs", mvf), ("hbma", mvh)):
                    ref = ns.smooth(ns.filters[mode], mv, fsz)
                    allok &= check(f"smooth {mode} fs{fsz} on {nm} {H}x{W}", np.asarray(ref, np.float32), O.smooth(mode, mv, fsz))
        # MCI
        bw = ns.fs.get_motion_vectors(8, 7, b, a)
        bwh = ns.hbma.get_motion_vectors(8, 7, 1, 2, 4, b, a)
        for dist in (0.5, 0.39999998, 0.8, 0.19999999, 0.6):
            d = np.float32(dist)
            for nm, f_, b_ in (("fs", mvf, bw), ("hbma", mvh, bwh),
                               ("fs-mean", ns.smooth(ns.filters["mean"], mvf, 4), ns.smooth(ns.filters["mean"], bw, 4))):
                allok &= check(f"unidir {nm} {H}x{W} d={dist}", ns.uni.helper(a, b, f_, d), O.unidir_helper(a, b, f_, d))
                allok &= check(f"bidir  {nm} {H}x{W} d={dist}", ns.bi.helper(a, b, f_, b_, d), O.bidir_helper(a, b, f_, b_, d))
        # adversarial random MV fields for MCI (wraps, OOB, fractional, many overlaps)
        for trial in range(4):
            f_ = np.zeros((H, W, 3), np.float32); b_ = np.zeros((H, W, 3), np.float32)
            g = 4
            for fld in (f_, b_):
                blk = rng.integers(-12, 13, size=(-(-H // g), -(-W // g), 2)).astype(np.float32)
                if trial % 2:
                    blk += rng.integers(0, 4, size=blk.shape).astype(np.float32) / 4
                sad = rng.integers(0, 40, size=(-(-H // g), -(-W // g), 1)).astype(np.float32)
                full = np.concatenate([blk, sad], axis=2)
                fld[:] = np.kron(full, np.ones((g, g, 1), np.float32))[:H, :W]
            for dist in (0.5, 0.39999998, 0.8):
                d = np.float32(dist)
                allok &= check(f"unidir rand{trial} {H}x{W} d={dist}", ns.uni.helper(a, b, f_, d), O.unidir_helper(a, b, f_, d))
                allok &= check(f"bidir  rand{trial} {H}x{W} d={dist}", ns.bi.helper(a, b, f_, b_, d), O.bidir_helper(a, b, f_, b_, d))
    print("ALL OK" if allok else "SOME FAILED")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
