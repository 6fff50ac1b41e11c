"""TEST INFRASTRUCTURE ONLY.  Seeded adversarial MCI inputs shared by oracle/validate_fuzz_vs_reference.py (which runs them
through the unmodified reference and writes tests/golden/fuzz_reference_sha.npz) and tests/test_oracle_golden.py (which
checks the C oracle against those hashes): random frame sizes, MV cells of different sizes per direction, SAD cells that
nest in the MV cells or not, vectors up to +-100 px (wrapped and out-of-frame targets), fractional vectors, SAD ties and
infinite SADs, odd dists."""
import hashlib

import numpy as np


def fuzz_field(rng, H, W, g, D, frac, sg=None, holes=0.03):
    """Dense [H, W, 3] field, (dy, dx) constant on g x g cells and SAD constant on sg x sg cells."""
    nr, nc = -(-H // g), -(-W // g)
    v = rng.integers(-D, D + 1, (nr, nc, 2)).astype(np.float32)
    if rng.integers(0, 3) == 0:
        v = np.repeat(np.repeat(v[::3, ::3], 3, 0), 3, 1)[:nr, :nc]
    if frac:
        v += rng.choice([0.0, 0.25, 0.5, 1 / 3], (nr, nc, 2)).astype(np.float32)
    sg = sg or g
    sr, sc = -(-H // sg), -(-W // sg)
    sad = rng.integers(0, 6, (sr, sc, 1)).astype(np.float32) * float(rng.choice([100.0, 1.0, 37.5]))
    sad[rng.random((sr, sc, 1)) < holes] = np.inf
    mv = np.kron(v, np.ones((g, g, 1), np.float32))[:H, :W]
    sd = np.kron(sad, np.ones((sg, sg, 1), np.float32))[:H, :W]
    return np.ascontiguousarray(np.concatenate([mv, sd], axis=2))


def mci_cases(seed, rounds):
    """Yield (index, src, tgt, fwd field, bwd field, float32 dist)."""
    rng = np.random.default_rng(seed)
    for it in range(rounds):
        H, W = int(rng.integers(8, 150)), int(rng.integers(8, 200))
        a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        b = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        gf, gb = int(rng.choice([1, 2, 4, 8, 16, 5])), int(rng.choice([1, 2, 4, 8, 16, 6]))
        D = int(rng.choice([0, 1, 3, 8, 20, 45, 100]))
        sgf = gf * int(rng.choice([1, 1, 2])) if rng.integers(0, 2) else int(rng.choice([3, 4, 8]))
        f_ = fuzz_field(rng, H, W, gf, D, rng.integers(0, 2), sgf)
        b_ = fuzz_field(rng, H, W, gb, D, rng.integers(0, 2))
        dist = np.float32(rng.choice([0.5, 0.19999999, 0.8, 0.4, 0.6, 1 / 3, 0.99]))
        yield it, a, b, f_, b_, dist


def bidir2_cases(seed, rounds):
    """Yield (index, src, tgt, fwd, bwd, python-float dist, upscale_MV): random fields small enough to stay inside the
    16-px pad most of the time (dense fields on 4 / 8 / 16-px cells, or block-level fields as HBMA hands them over)."""
    rng = np.random.default_rng(seed)
    for it in range(rounds):
        H, W = int(rng.integers(8, 120)), int(rng.integers(8, 160))
        a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        b = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        up = bool(rng.integers(0, 2)); D2 = int(rng.choice([0, 2, 6, 14])); dd = float(rng.choice([0.5, 0.4, 0.8, 0.25]))
        if up:
            g2 = int(rng.choice([4, 8, 16]))
            f2 = fuzz_field(rng, H, W, g2, D2, rng.integers(0, 2), holes=0)
            b2 = fuzz_field(rng, H, W, g2, D2, rng.integers(0, 2), holes=0)
        else:
            nr, nc = -(-H // 4), -(-W // 4)
            f2 = fuzz_field(rng, nr, nc, 1, D2, rng.integers(0, 2), holes=0)
            b2 = fuzz_field(rng, nr, nc, 1, D2, rng.integers(0, 2), holes=0)
        yield it, a, b, f2, b2, dd, up


def me_cases(seed, rounds):
    """Yield (index, frame pair, dict of parameters) for full search, TSS, HBMA and the three smoothing filters."""
    rng = np.random.default_rng(seed)
    from interpolatory_b200.synth import make_sequence
    for it in range(rounds):
        H, W = int(rng.integers(8, 100)), int(rng.integers(8, 130))
        kind = int(rng.integers(0, 4))
        if kind == 0 and H >= 48 and W >= 48:
            fr = make_sequence(H, W, 2, seed=int(rng.integers(1 << 30)))
        elif kind == 1:
            fr = np.repeat(rng.integers(0, 2, (2, H, W, 1), dtype=np.uint8) * 50 + 20, 3, axis=3)      # ties, exact-integer SADs
        elif kind == 2:
            a = np.repeat(rng.integers(0, 200, (H, W, 1), dtype=np.uint8), 3, axis=2)                    # shifted + grey offset
            b = (np.roll(a, (int(rng.integers(-4, 5)), int(rng.integers(-4, 5))), (0, 1)).astype(np.int32) + int(rng.integers(0, 5))).astype(np.uint8)
            fr = np.stack([a, b])
        else:
            fr = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
        prm = dict(bs=int(rng.choice([2, 3, 4, 5, 6, 8, 8, 12, 16])), R=int(rng.integers(0, 12)), steps=int(rng.integers(1, 5)),
                   hb=int(rng.choice([4, 8, 8, 16])), win=int(rng.integers(1, 8)), sub=int(rng.integers(1, 4)), hsteps=int(rng.integers(0, 4)),
                   mbs=int(rng.choice([2, 4, 4, 8])), cost=str(rng.choice(['sad', 'ssd'])), up=bool(rng.integers(0, 2)),
                   g=int(rng.choice([1, 2, 4, 8])), fsz=int(rng.choice([1, 2, 4, 8, 3])), frac=int(rng.integers(0, 2)))
        mv = fuzz_field(rng, H, W, prm['g'], 20, prm['frac'])
        yield it, fr, prm, mv


def sha(x):
    return hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
