"""Seeded synthetic frame sequences for parity tests and bench.py.

A translating blurred random texture with per-frame sensor noise, plus four
features that each exercise one corner of the reference's semantics:
  (i)   an inverted square moving against the background  -> occlusion, holes, overlaps
  (ii)  a flat constant-colour rectangle                  -> SAD ties (distance / scan-order tie-break)
  (iii) a noise-free patch that gains a pure grey offset  -> luma SAD lands exactly on an integer
        between frames                                       (the f64 truncation fix-up, N2)
  (iv)  a strip near the right edge moving faster than any search radius -> unmatched blocks,
        out-of-frame and wrapped motion-compensated targets (W1)
"""
import numpy as np


def _box2(a):
    """2x2 box blur (same size, edge replicated), float32."""
    p = np.pad(a, ((0, 1), (0, 1), (0, 0)), mode="edge")
    return 0.25 * (p[:-1, :-1] + p[1:, :-1] + p[:-1, 1:] + p[1:, 1:])


def make_sequence(H, W, n_frames, seed=0):
    """Return uint8 [n_frames, H, W, 3]."""
    rng = np.random.default_rng(seed)
    s = max(1, int(round(H / 360)))
    margin = 8 * 8 * s + 64
    ch, cw = H + margin, W + margin
    low = rng.integers(0, 256, size=(ch // 8 + 8, cw // 8 + 8, 3)).astype(np.float32)
    canvas = np.kron(low, np.ones((8, 8, 1), np.float32))
    canvas = _box2(_box2(canvas))[:ch, :cw]
    frames = np.empty((n_frames, H, W, 3), np.uint8)
    sq = 24 * s
    flat_h, flat_w = 16 * s + 3, 40 * s + 5
    flat_r, flat_c = H // 2, W // 4
    pr, pc, ph, pw = H // 5, W // 2, 24 * s, 32 * s
    strip_w = 24 * s
    base_patch = None
    period = margin // (3 * s)          # frames until the crop would leave the canvas: the pan then bounces back
    for t in range(n_frames):
        tt = t % (2 * period)
        tt = tt if tt <= period else 2 * period - tt
        oy, ox = 2 * tt * s, 3 * tt * s
        f = canvas[oy:oy + H, ox:ox + W].copy()
        f += rng.normal(0.0, 2.0, size=f.shape).astype(np.float32)
        # (iv) fast strip on the right edge: jumps 40*s px per frame vertically
        sy = (oy + 40 * s * t) % (ch - H)
        f[:, W - strip_w:] = canvas[sy:sy + H, ox + W - strip_w:ox + W]
        # (ii) flat rectangle, static
        f[flat_r:flat_r + flat_h, flat_c:flat_c + flat_w] = np.array([90.0, 140.0, 60.0], np.float32)
        # (iii) noise-free patch, static, +3 grey levels per frame
        if base_patch is None:
            ph, pw = min(ph, H - pr), min(pw, W - pc)
            base_patch = np.clip(np.rint(canvas[pr:pr + ph, pc:pc + pw]), 0, 200)
        f[pr:pr + ph, pc:pc + pw] = base_patch + 3.0 * t
        # (i) inverted square moving (5,-4)*s px per frame
        r0 = (H // 3 + 5 * s * t) % max(1, H - sq)
        c0 = (2 * W // 3 - 4 * s * t) % max(1, W - sq)
        f[r0:r0 + sq, c0:c0 + sq] = 255.0 - f[r0:r0 + sq, c0:c0 + sq]
        frames[t] = np.clip(np.rint(f), 0, 255).astype(np.uint8)
    return frames


def make_pair(H, W, seed=0):
    fr = make_sequence(H, W, 2, seed)
    return fr[0], fr[1]


def make_pair_tiled(H, W, seed=0, tiles=4):
    """Full-size pair for the 4K / 8K parity and bench workloads: a (H/tiles x W/tiles) synthetic pair tiled
    tiles x tiles (so the motion stays inside a +-32 search at 8K, where make_sequence's pan would be 36 px per
    frame), with independent seeded +-3 grey-level noise per frame that breaks the periodicity."""
    base = make_sequence(H // tiles, W // tiles, 2, seed)
    rng = np.random.default_rng(seed + 7919)
    out = []
    for t in range(2):
        f = np.tile(base[t], (tiles, tiles, 1)).astype(np.int16)
        f += rng.integers(-3, 4, size=(H, W, 1), dtype=np.int16)
        out.append(np.clip(f, 0, 255).astype(np.uint8))
    return out[0], out[1]
