"""Host-side helpers that define `dist` and the settings string (reference: src/util.py)."""
import math
import sys

import numpy as np


def eprint(s):
    print(s, file=sys.stderr)


progress_file_path = None

# src/Globals.py: the flags the harness and the interpolators consult
debug_flags = {
    'debug_IO': False,
    'debug_interpolator': False,
    'debug_progress': True,
    'debug_timer': False,
    'debug_benchmark_progress': True,
    'debug_MEMCI_args': True,
}


def sToMMSS(s):
    """seconds -> 'M:SS' (src/util.py:13-16)."""
    return f"{int(s / 60)}:{str(int(s % 60)).rjust(2, '0')}"


def getETA(elapsed_seconds, processed_frames, all_frames):
    """Remaining time at the rate so far (src/util.py:22-26)."""
    remaining = all_frames - processed_frames
    return sToMMSS(0 if processed_frames == 0 else int(float(elapsed_seconds) / processed_frames * remaining))


def signal_progress(s):
    """Overwrite the progress file the GUI polls, if one is set (src/util.py:29-34)."""
    if progress_file_path is not None:
        with open(progress_file_path, 'w') as f:
            f.write(s)
            f.flush()


def get_first_frame_idx_and_ratio(idx, rate_ratio):
    """(index of the last source frame at or before output frame `idx`, 1 - fractional position).

    Bit-identical to src/util.py:49-60 under its numba signature (int32, float32) ->
    (int32, float32): rate_ratio is rounded to float32 on entry, the division and the
    subtraction run in float64, the ratio is rounded to float32 on return (and comes back to
    Python as a float).  The caller's `dist = 1. - ratio` therefore reproduces the reference's
    0x3ecccccc / 0x3f4ccccd / ... cycle.
    """
    rr = float(np.float32(rate_ratio))
    q = float(int(idx)) / rr
    first = math.floor(q)
    return first, float(np.float32(1.0 - (q - first)))


def deconstruct_settings(s):
    """'key=val.key=val' -> dict of strings (src/util.py:65-80); malformed pairs are reported and skipped."""
    out = {}
    for pair in s.split('.'):
        kv = pair.split('=')
        if len(kv) != 2:
            eprint(f'Ignoring {pair}, malformed input')
            continue
        k, v = kv
        if k and v:
            out[k] = v
        else:
            eprint(f'Ignoring {k} {v}, malformed input')
    return out


def deconstruct_interpolation_mode_and_settings(s):
    """'<mode>[:<settings>]' -> (mode, settings dict)  (src/util.py:83-92)."""
    parts = s.split(':')
    if len(parts) == 1:
        return parts[0], {}
    if len(parts) == 2:
        return parts[0], deconstruct_settings(parts[1])
    eprint(f'Invalid interpolation-mode and settings config: {s}')
    raise SystemExit(1)


def hbma_auto_steps(shape):
    """Pyramid depth rule applied by the interpolator classes before calling HBMA
    (MCI/Unidirectional.py:116-121, MCI/Bidirectional.py:144-149): 1080p -> 4, 360p -> 2."""
    min_side = min(shape[0], shape[1])
    steps = 1
    while min_side > 255:
        min_side /= 2
        steps += 1
    return steps


def hbma_reference_undefined(H, W, block_size, steps, min_block_size):
    """True where the reference's HBMA reads outside its motion field: `increase_vec_density_jit` takes
    `mvs[row + 1]` / `mvs[row, col + 1]` as the neighbour of the first block row / column (ME/hbma.py:108-115), which does
    not exist when a coarser field has a single block row or column, and numba does not bounds-check.  This
    implementation clamps the neighbour index there; its result is well defined, the reference's is not."""
    sizes = [(H, W)]
    for _ in range(steps):
        h, w = sizes[-1]
        sizes.append(((h + 1) // 2, (w + 1) // 2))
    rows, cols = -(-sizes[-1][0] // block_size), -(-sizes[-1][1] // block_size)
    for (h, w) in sizes[-2::-1]:                       # densification through the pyramid at constant block size
        if rows == 1 or cols == 1:
            return True
        rows, cols = -(-h // block_size), -(-w // block_size)
    bs = block_size
    while bs > min_block_size:                         # then the block size is halved at full resolution
        if rows == 1 or cols == 1:
            return True
        bs >>= 1
        rows, cols = -(-H // bs), -(-W // bs)
    return False


def warn_if_hbma_reference_undefined(H, W, block_size, steps, min_block_size):
    """One warning per geometry (Python's default filter): the caller is comparing against a reference whose own output
    depends on whatever lies behind its array here."""
    if hbma_reference_undefined(H, W, block_size, steps, min_block_size):
        import warnings
        warnings.warn(f"HBMA on a {H}x{W} frame with block_size {block_size}, {steps} pyramid steps, min_block_size {min_block_size}: "
                      "a motion field of the pyramid has a single block row or column, where the reference reads out of bounds "
                      "(ME/hbma.py:108-115); the neighbour index is clamped here", RuntimeWarning, stacklevel=3)
