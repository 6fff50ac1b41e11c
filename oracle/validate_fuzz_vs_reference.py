"""The seeded random MV-field geometry of tests/test_gpu_parity.py::test_mci_fuzz_random_geometry (wrapped and out-of-frame
targets, fractional vectors, MV / SAD cells of different sizes, SAD ties and infinite SADs) and random full-search / TSS / HBMA /
smoothing parameters, run through the UNMODIFIED reference in the build container and compared bit for bit with the C oracle.
Not part of the pytest suite (needs /root/reference + numba); run:  python oracle/validate_fuzz_vs_reference.py [rounds]
`--write-golden` also runs oracle/fuzz_cases.py's seeded MCI cases through the reference and stores the SHA-256 of its frames
in tests/golden/fuzz_reference_sha.npz, which tests/test_oracle_golden.py checks the oracle (and the GPU suite the kernels) against."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import oracle as O          # noqa: E402
import ref_harness          # noqa: E402
from interpolatory_b200.synth import make_sequence  # noqa: E402


def fuzz_field(rng, H, W, g, D, frac, sg=None, holes=0.03):
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


def write_golden(ns, seed=2025, rounds=96):
    import fuzz_cases as F
    out = {"seed": np.int64(seed), "rounds": np.int64(rounds)}
    n = 0
    for it, a, b, f_, b_, dist in F.mci_cases(seed, rounds):
        try:
            O.unidir_helper(a, b, f_, dist); O.bidir_helper(a, b, f_, b_, dist)
        except RuntimeError:
            out[f"{it}/undefined"] = np.int64(1)       # a displaced index below -H: out-of-bounds write in the reference
            continue
        out[f"{it}/unidir"] = F.sha(np.asarray(ns.uni.helper(a, b, f_, dist)))
        out[f"{it}/bidir"] = F.sha(np.asarray(ns.bi.helper(a, b, f_, b_, dist)))
        n += 1
    # bidir2: IEWMC both ways + error-adaptive combination + BDHI + unpad (MCI/Bidirectional_2.py:330-345)
    from src.Interpolators.MEMCI.MCI import Bidirectional_2 as b2
    out["b2_seed"] = np.int64(seed + 1); out["b2_rounds"] = np.int64(64)
    n2 = n2_fail = 0
    for it, a, b, f2, bw2, dd, up in F.bidir2_cases(seed + 1, 64):
        try:
            mine = O.bidir2_frame(a, b, f2, bw2, dd, up)
        except RuntimeError:
            out[f"b2/{it}/out_of_pad"] = np.int64(1)   # a shifted block leaves the padded frame: the reference raises too
            continue
        Ff, Ef = b2.IEWMC(f2, a, b, dd, 16, 4, up)
        Fb, Eb = b2.IEWMC(bw2, b, a, 1 - dd, 16, 4, up)
        ref = b2.unpad_and_downconvert(b2.BDHI(b2.Error_adaptive_combination(Ff, Ef, Fb, Eb), 4, 16), 16)
        if not np.array_equal(np.asarray(ref), mine):
            n2_fail += 1
            print("FAIL bidir2", it, a.shape, dd, up)
        out[f"b2/{it}/out"] = F.sha(np.asarray(ref))
        n2 += 1
    print(f"bidir2: {n2} cases vs oracle, {n2_fail} mismatches")
    # motion estimation and smoothing with random parameters
    out["me_seed"] = np.int64(seed + 2); out["me_rounds"] = np.int64(60)
    n3 = n3_fail = 0
    for it, fr, q, mv in F.me_cases(seed + 2, 60):
        a, b = fr[0], fr[1]
        calls = {
            "fs": (lambda: ns.fs.get_motion_vectors(q['bs'], q['R'], a, b), lambda: O.fs(q['bs'], q['R'], a, b)),
            "tss": (lambda: ns.tss.get_motion_vectors(q['bs'], q['steps'], a, b), lambda: O.tss(q['bs'], q['steps'], a, b)),
            "hbma": (lambda: ns.hbma.get_motion_vectors(q['hb'], q['win'], q['sub'], q['hsteps'], q['mbs'], a, b, q['cost'], q['up']),
                     lambda: O.hbma(q['hb'], q['win'], q['sub'], q['hsteps'], q['mbs'], a, b, q['cost'], q['up'])),
        }
        for mode in ("mean", "median", "weighted"):
            calls[f"smooth_{mode}"] = (lambda mode=mode: ns.smooth(ns.filters[mode], mv.copy(), q['fsz']), lambda mode=mode: O.smooth(mode, mv, q['fsz']))
        for name, (fr_ref, fr_mine) in calls.items():
            if name == "hbma" and O.hbma_undefined_in_reference(a.shape[0], a.shape[1], q['hb'], q['hsteps'], q['mbs']):
                out[f"me/{it}/hbma/undefined"] = np.int64(1)      # single block row / column: out-of-bounds neighbour read
                continue
            try:
                ref = np.asarray(fr_ref())
            except Exception as e:                     # parameters the reference itself rejects / crashes on: recorded, not compared
                out[f"me/{it}/{name}/raises"] = type(e).__name__
                continue
            mine = fr_mine()
            if ref.shape != mine.shape or not np.array_equal(ref, mine):
                n3_fail += 1
                print("FAIL", name, it, a.shape, q)
            out[f"me/{it}/{name}"] = F.sha(ref.astype(np.float32))
            n3 += 1
    print(f"ME / smoothing: {n3} cases vs oracle, {n3_fail} mismatches")
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "fuzz_reference_sha.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {n} defined cases of {rounds}")


def main():
    args = [x for x in sys.argv[1:] if not x.startswith("--")]
    rounds = int(args[0]) if args else 120
    ns = ref_harness.load()
    if "--write-golden" in sys.argv:
        write_golden(ns)
        return 0
    rng = np.random.default_rng(11)
    bad = n_mci = n_undef = n_me = 0
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
        try:
            mine_u, mine_b = O.unidir_helper(a, b, f_, dist), O.bidir_helper(a, b, f_, b_, dist)
        except RuntimeError:
            n_undef += 1            # a displaced index below -H: the reference writes out of bounds there, not run
            continue
        ref_u = ns.uni.helper(a, b, f_, dist)
        ref_b = ns.bi.helper(a, b, f_, b_, dist)
        n_mci += 1
        for name, r, m in (("unidir", ref_u, mine_u), ("bidir", ref_b, mine_b)):
            if not np.array_equal(np.asarray(r), m):
                bad += 1
                print("FAIL", name, (H, W, gf, gb, D, sgf, float(dist)), int((np.asarray(r) != m).any(axis=2).sum()))
        if it % 4 == 0:             # motion estimation / smoothing with random parameters (slower in the reference)
            Hs, Ws = min(H, 80), min(W, 96)
            if Hs >= 48 and Ws >= 48:
                fr = make_sequence(Hs, Ws, 2, seed=int(rng.integers(1 << 30)))
            else:
                fr = rng.integers(0, 256, (2, Hs, Ws, 3), dtype=np.uint8)
            bs = int(rng.choice([2, 4, 5, 8, 16])); R = int(rng.integers(0, 9)); steps = int(rng.integers(1, 4))
            checks = [("fs", ns.fs.get_motion_vectors(bs, R, fr[0], fr[1]), O.fs(bs, R, fr[0], fr[1])),
                      ("tss", ns.tss.get_motion_vectors(bs, steps, fr[0], fr[1]), O.tss(bs, steps, fr[0], fr[1]))]
            mv = fuzz_field(rng, Hs, Ws, int(rng.choice([1, 2, 4, 8])), 20, rng.integers(0, 2))
            fsz = int(rng.choice([1, 2, 4, 8, 3]))
            for mode in ("mean", "median", "weighted"):
                checks.append((f"smooth {mode}", ns.smooth(ns.filters[mode], mv.copy(), fsz), O.smooth(mode, mv, fsz)))
            for name, r, m in checks:
                n_me += 1
                if not np.array_equal(np.asarray(r), m):
                    bad += 1
                    print("FAIL", name, (Hs, Ws, bs, R, steps, fsz))
    print(f"{n_mci} MCI geometries x 2 helpers, {n_me} ME / smoothing cases, {n_undef} undefined in the reference (skipped): "
          f"{'ALL OK' if bad == 0 else str(bad) + ' FAILED'}")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
