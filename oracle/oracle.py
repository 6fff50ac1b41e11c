"""ctypes front-end of the CPU oracle (oracle/memci_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of memci_oracle.c.  The function
names and argument order mirror the reference so parity tests read like calls
into the reference:

  fs(block_size, target_region, im1, im2)            ME/full_search.py:61
  tss(block_size, steps, im1, im2)                   ME/tss.py:63
  hbma(block_size, win, sub_win, steps, min_bs, im1, im2, cost_str, upscale)   ME/hbma.py:161
  smooth(mode, mv_field, filter_size)                smoothing/threeXthree_mv_smoothing.py:14
  unidir_helper(src, tgt, mv, dist)                  MCI/Unidirectional.py:36
  bidir_helper(src, tgt, fwd, bwd, dist)             MCI/Bidirectional.py:27
  get_first_frame_idx_and_ratio(idx, rate_ratio)     src/util.py:49
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libmemci_oracle.so")
_lib = None

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "memci_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_fs_blocks.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
        L.orc_tss_blocks.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
        L.orc_fs_work.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_longlong)] * 2
        L.orc_expand_dense.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
        L.orc_pyr_down.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
        L.orc_hbma_shape.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int)] * 3
        L.orc_hbma_blocks.argtypes = [_u8p, _u8p] + [C.c_int] * 8 + [_f32p]
        L.orc_smooth_dense.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
        L.orc_mci_unidir.argtypes = [_u8p, _u8p, _f32p, C.c_float, C.c_int, C.c_int, _u8p]
        L.orc_mci_bidir.argtypes = [_u8p, _u8p, _f32p, _f32p, C.c_float, C.c_int, C.c_int, _u8p]
        L.orc_iewmc.argtypes = [_u8p, _u8p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _f32p, _f32p, _f32p]
        L.orc_eac.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_size_t, _f32p]
        L.orc_bdhi_unpad.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.orc_fill_holes.argtypes = [_f32p, C.c_int, _f32p]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed with code {rc}")


def _frame(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] == 3, a.shape
    return a


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def fs_blocks(block_size, target_region, im1, im2):
    im1, im2 = _frame(im1), _frame(im2)
    H, W, _ = im1.shape
    nbr, nbc = -(-H // block_size), -(-W // block_size)
    out = np.empty((nbr, nbc, 3), np.float32)
    _chk(lib().orc_fs_blocks(im1, im2, H, W, block_size, target_region, out), "fs")
    return out


def expand_dense(blocks, bs, H, W):
    blocks = np.ascontiguousarray(blocks, np.float32)
    dense = np.empty((H, W, 3), np.float32)
    _chk(lib().orc_expand_dense(blocks, blocks.shape[0], blocks.shape[1], bs, H, W, dense), "expand")
    return dense


def fs(block_size, target_region, im1, im2):
    H, W, _ = im1.shape
    return expand_dense(fs_blocks(block_size, target_region, im1, im2), block_size, H, W)


def tss_blocks(block_size, steps, im1, im2):
    im1, im2 = _frame(im1), _frame(im2)
    H, W, _ = im1.shape
    out = np.empty((-(-H // block_size), -(-W // block_size), 3), np.float32)
    _chk(lib().orc_tss_blocks(im1, im2, H, W, block_size, steps, out), "tss")
    return out


def tss(block_size, steps, im1, im2):
    H, W, _ = im1.shape
    return expand_dense(tss_blocks(block_size, steps, im1, im2), block_size, H, W)


def fs_work(H, W, block_size, target_region):
    n, ops = C.c_longlong(), C.c_longlong()
    lib().orc_fs_work(H, W, block_size, target_region, C.byref(n), C.byref(ops))
    return n.value, ops.value


def pyr_down(im):
    im = _frame(im)
    H, W, _ = im.shape
    out = np.empty(((H + 1) // 2, (W + 1) // 2, 3), np.uint8)
    _chk(lib().orc_pyr_down(im, H, W, out), "pyr_down")
    return out


def hbma(block_size, win_size, sub_win_size, steps, min_block_size, im1, im2, cost_str="sad", upscale=True):
    im1, im2 = _frame(im1), _frame(im2)
    H, W, _ = im1.shape
    r, c, b = C.c_int(), C.c_int(), C.c_int()
    _chk(lib().orc_hbma_shape(H, W, block_size, min_block_size, C.byref(r), C.byref(c), C.byref(b)), "hbma_shape")
    out = np.empty((r.value, c.value, 3), np.float32)
    key = {"sad": 0, "ssd": 1}[cost_str]
    _chk(lib().orc_hbma_blocks(im1, im2, H, W, block_size, win_size, sub_win_size, steps,
                               min_block_size, key, out), "hbma")
    if upscale:
        return expand_dense(out, b.value, H, W)
    return out


def hbma_undefined_in_reference(H, W, block_size, steps, min_block_size):
    """True where the reference's HBMA reads out of bounds (numba does not bounds-check): increase_vec_density_jit
    takes `mvs[row + 1]` / `mvs[row, col + 1]` as the neighbour of the first row / column (hbma.py:108-115), which does
    not exist when the coarser field has a single block row or column.  The oracle (and the CUDA path) clamp there."""
    sizes = [(H, W)]
    for _ in range(steps):
        h, w = sizes[-1]
        sizes.append(((h + 1) // 2, (w + 1) // 2))
    rows, cols = -(-sizes[-1][0] // block_size), -(-sizes[-1][1] // block_size)
    for (h, w) in sizes[-2::-1]:                       # densify through the pyramid at constant block size
        if rows == 1 or cols == 1:
            return True
        rows, cols = -(-h // block_size), -(-w // block_size)
    bs = block_size
    while bs > min_block_size:                         # then halve the block size at full resolution
        if rows == 1 or cols == 1:
            return True
        bs >>= 1
        rows, cols = -(-H // bs), -(-W // bs)
    return False


_MODES = {None: 0, "none": 0, "mean": 1, "median": 2, "weighted": 3}


def smooth(mode, mv_field, filter_size):
    if _MODES[mode] == 0:
        return mv_field
    mv = np.ascontiguousarray(mv_field, np.float32)
    out = np.empty_like(mv)
    _chk(lib().orc_smooth_dense(mv, mv.shape[0], mv.shape[1], _MODES[mode], filter_size, out), "smooth")
    return out


def unidir_helper(src, tgt, mv, dist):
    src, tgt = _frame(src), _frame(tgt)
    H, W, _ = src.shape
    out = np.empty_like(src)
    _chk(lib().orc_mci_unidir(src, tgt, np.ascontiguousarray(mv, np.float32), np.float32(dist), H, W, out), "unidir")
    return out


def bidir_helper(src, tgt, fwd, bwd, dist):
    src, tgt = _frame(src), _frame(tgt)
    H, W, _ = src.shape
    out = np.empty_like(src)
    _chk(lib().orc_mci_bidir(src, tgt, np.ascontiguousarray(fwd, np.float32),
                             np.ascontiguousarray(bwd, np.float32), np.float32(dist), H, W, out), "bidir")
    return out


def get_first_frame_idx_and_ratio(idx, rate_ratio):
    """src/util.py:49-60 under its numba signature (int32, float32) -> (int32, float32):
    the division and subtraction run in float64 on the float32-rounded rate_ratio,
    the ratio is rounded to float32 on return (probed against the reference)."""
    rr = float(np.float32(rate_ratio))
    q = float(int(idx)) / rr
    fi = math.floor(q)
    return fi, float(np.float32(1.0 - (q - fi)))


def hbma_auto_steps(H, W):
    """MCI/Unidirectional.py:116-121 / Bidirectional.py:144-149 (Q8)."""
    m = min(H, W)
    s = 1
    while m > 255:
        m /= 2
        s += 1
    return s


# ---- bidir2 (MCI/Bidirectional_2.py) -------------------------------------------------------
def weight_kern(kernlen=8):
    """get_weight_kern (Bidirectional_2.py:33-47), one channel, float32."""
    x = np.linspace(1, 1.1, int(kernlen / 2))
    x = np.append(x, np.flip(x))
    xx, yy = np.meshgrid(x, x)
    return np.ascontiguousarray((6 * np.square(5 * (np.log2(xx) + np.log2(yy)))).astype(np.float32))


def iewmc(mv_field, F1, F2, dist, pad_size=16, min_block_size=4, upscale_mv=True):
    """IEWMC (Bidirectional_2.py:130-144) -> (F, E) float32 [H+2p][W+2p][3]."""
    F1 = _frame(F1); F2 = _frame(F2)
    H, W = F1.shape[:2]
    mv = np.ascontiguousarray(mv_field, np.float32)
    w = weight_kern(2 * min_block_size)
    F = np.empty((H + 2 * pad_size, W + 2 * pad_size, 3), np.float32)
    E = np.empty_like(F)
    rc = lib().orc_iewmc(F1, F2, mv, int(bool(upscale_mv)), H, W, int(min_block_size), int(pad_size),
                         C.c_float(np.float32(dist)), w, F, E)
    _chk(rc, "orc_iewmc")
    return F, E


def eac(Ff, Ef, Fb, Eb):
    """Error_adaptive_combination (Bidirectional_2.py:147-166)."""
    a = [np.ascontiguousarray(x, np.float32) for x in (Ff, Ef, Fb, Eb)]
    Fc = np.empty_like(a[0])
    lib().orc_eac(a[0], a[1], a[2], a[3], a[0].shape[0] * a[0].shape[1], Fc)
    return Fc


def bdhi_unpad(Fc, min_block_size=4, pad_size=16):
    """BDHI + unpad_and_downconvert (Bidirectional_2.py:265-293) -> uint8 [H][W][3]."""
    Fc = np.ascontiguousarray(Fc, np.float32)
    H, W = Fc.shape[0] - 2 * pad_size, Fc.shape[1] - 2 * pad_size
    out = np.empty((H, W, 3), np.uint8)
    lib().orc_bdhi_unpad(Fc, H, W, int(min_block_size), int(pad_size), out)
    return out


def bidir2_frame(src, tgt, fwd, bwd, dist, upscale_mv, pad_size=16, min_block_size=4):
    """BiDir2Interpolator.get_interpolated_frame after ME (Bidirectional_2.py:330-345); dist is the python float."""
    Ff, Ef = iewmc(fwd, src, tgt, dist, pad_size, min_block_size, upscale_mv)
    Fb, Eb = iewmc(bwd, tgt, src, 1 - dist, pad_size, min_block_size, upscale_mv)
    return bdhi_unpad(eac(Ff, Ef, Fb, Eb), min_block_size, pad_size)


def fill_holes(block):
    """fill_holes (Bidirectional_2.py:168-262) on one n x n x 3 float32 block."""
    b = np.ascontiguousarray(block, np.float32)
    out = np.empty_like(b)
    lib().orc_fill_holes(b, b.shape[0], out)
    return out
