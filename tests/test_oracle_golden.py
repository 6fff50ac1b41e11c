"""CPU suite, part 1: the oracle (oracle/memci_oracle.c) against the committed golden fixtures,
i.e. against outputs of the UNMODIFIED reference (tests/golden/*.npz, made by oracle/gen_golden.py).
This is what pins the oracle; the GPU tests then compare the CUDA path with the oracle."""
import hashlib

import numpy as np
import pytest

from interpolatory_b200.synth import make_sequence

SHAPES = [(72, 96), (50, 70), (37, 53)]
FS_CFGS = [(8, 7), (8, 3), (4, 5), (16, 9), (8, 16)]
HB_CFGS = [(8, 7, 1, 2, 4), (8, 3, 1, 1, 4), (8, 5, 2, 3, 2), (4, 2, 1, 1, 4), (16, 4, 1, 1, 4)]
DISTS = [0.5, 0.39999998, 0.8, 0.19999999, 0.6]


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest() + f":{a.dtype}:{'x'.join(map(str, a.shape))}"


def test_ratio_table(oracle, golden):
    """get_first_frame_idx_and_ratio: oracle copy and the product's host copy (src/util.py:49-60)."""
    from interpolatory_b200.util import get_first_frame_idx_and_ratio as host_fn
    tab = golden("ratio.npz")["table"]
    for rr, idx, first, ratio in tab:
        assert oracle.get_first_frame_idx_and_ratio(int(idx), rr) == (int(first), float(ratio))
        assert host_fn(int(idx), rr) == (int(first), float(ratio))


@pytest.mark.parametrize("shape", SHAPES)
def test_fs(oracle, golden, shape):
    g = golden("functions_small.npz")
    k = f"{shape[0]}x{shape[1]}"
    a, b = g[f"{k}/a"], g[f"{k}/b"]
    fr = make_sequence(shape[0], shape[1], 2, seed=shape[0])
    assert np.array_equal(fr[0], a) and np.array_equal(fr[1], b), "synthetic generator drifted from the fixtures"
    for bs, R in FS_CFGS:
        assert np.array_equal(oracle.fs_blocks(bs, R, a, b), g[f"{k}/fs/{bs}_{R}"]), (shape, bs, R)
    assert np.array_equal(oracle.fs(8, 7, a, b), g[f"{k}/fs_dense/8_7"])


@pytest.mark.parametrize("shape", SHAPES)
def test_hbma_and_pyramid(oracle, golden, shape):
    g = golden("functions_small.npz")
    k = f"{shape[0]}x{shape[1]}"
    a, b = g[f"{k}/a"], g[f"{k}/b"]
    lvl = a
    for s in (1, 2):
        lvl = oracle.pyr_down(lvl)
        assert np.array_equal(lvl, g[f"{k}/pyr/{s}"])
    for cfg in HB_CFGS:
        ref = g[f"{k}/hbma/{'_'.join(map(str, cfg))}"]
        assert np.array_equal(oracle.hbma(*cfg, a, b, 'sad', False), ref), (shape, cfg)
    assert np.array_equal(oracle.hbma(8, 7, 1, 2, 4, a, b), g[f"{k}/hbma_dense/8_7_1_2_4"])


@pytest.mark.parametrize("shape", SHAPES)
def test_smooth(oracle, golden, shape):
    g = golden("functions_small.npz")
    k = f"{shape[0]}x{shape[1]}"
    fields = {"fs": g[f"{k}/fs_dense/8_7"], "hbma": g[f"{k}/hbma_dense/8_7_1_2_4"]}
    for mode in ("mean", "median", "weighted"):
        for fsz in (4, 8, 3):
            for nm, mv in fields.items():
                out = oracle.smooth(mode, mv, fsz)
                assert np.array_equal(out[:, :, :2], g[f"{k}/smooth/{mode}_{fsz}/{nm}"]), (shape, mode, fsz, nm)
                assert np.array_equal(out[:, :, 2], mv[:, :, 2])


@pytest.mark.parametrize("shape", SHAPES)
def test_mci(oracle, golden, shape):
    g = golden("functions_small.npz")
    k = f"{shape[0]}x{shape[1]}"
    H, W = shape
    a, b = g[f"{k}/a"], g[f"{k}/b"]
    fwd, bwd = g[f"{k}/fs_dense/8_7"], oracle.fs(8, 7, b, a)
    fwh, bwh = g[f"{k}/hbma_dense/8_7_1_2_4"], oracle.hbma(8, 7, 1, 2, 4, b, a)
    for dist in DISTS:
        d = np.float32(dist)
        assert np.array_equal(oracle.unidir_helper(a, b, fwd, d), g[f"{k}/unidir/fs/{dist}"])
        assert np.array_equal(oracle.bidir_helper(a, b, fwd, bwd, d), g[f"{k}/bidir/fs/{dist}"])
        assert np.array_equal(oracle.unidir_helper(a, b, fwh, d), g[f"{k}/unidir/hbma/{dist}"])
        assert np.array_equal(oracle.bidir_helper(a, b, fwh, bwh, d), g[f"{k}/bidir/hbma/{dist}"])
    for trial in range(3):
        gsz = int(g[f"{k}/rand/{trial}/g"])
        f_ = np.kron(g[f"{k}/rand/{trial}/fwd"], np.ones((gsz, gsz, 1), np.float32))[:H, :W].copy()
        b_ = np.kron(g[f"{k}/rand/{trial}/bwd"], np.ones((gsz, gsz, 1), np.float32))[:H, :W].copy()
        for dist in (0.5, 0.39999998, 0.8):
            d = np.float32(dist)
            assert np.array_equal(oracle.unidir_helper(a, b, f_, d), g[f"{k}/rand/{trial}/unidir/{dist}"])
            assert np.array_equal(oracle.bidir_helper(a, b, f_, b_, d), g[f"{k}/rand/{trial}/bidir/{dist}"])


def test_middlebury_crops(oracle, golden):
    g = golden("middlebury_crops.npz")
    for name in ("RubberWhale", "Dimetrodon"):
        a, b = g[f"{name}/a"], g[f"{name}/b"]
        fwd, bwd = oracle.fs(8, 7, a, b), oracle.fs(8, 7, b, a)
        assert np.array_equal(fwd[::8, ::8], g[f"{name}/fs_fwd"])
        assert np.array_equal(bwd[::8, ::8], g[f"{name}/fs_bwd"])
        assert np.array_equal(oracle.hbma(8, 7, 1, 1, 4, a, b, 'sad', False), g[f"{name}/hbma_fwd"])
        assert np.array_equal(oracle.bidir_helper(a, b, fwd, bwd, np.float32(0.5)), g[f"{name}/bidir"])
        assert np.array_equal(oracle.unidir_helper(a, b, fwd, np.float32(0.5)), g[f"{name}/unidir"])


def test_class_api_small_clip(oracle, golden):
    """The reference classes' whole-clip output (dist cycle, pass-through, MV cache) rebuilt from
    oracle calls, for the two small clips of e2e.npz."""
    import ast
    from interpolatory_b200.util import get_first_frame_idx_and_ratio, hbma_auto_steps
    g = golden("e2e.npz")
    for key in ("72x96_24_60", "56x88_30_60"):
        frames = g[f"{key}/frames"]
        H, W = frames.shape[1:3]
        fps_in, fps_out = int(key.split("_")[1]), int(key.split("_")[2])
        rr = fps_out / fps_in
        for si in range(7):
            st = dict(ast.literal_eval(str(g[f"{key}/set{si}/settings"])))
            want = g[f"{key}/set{si}/out"]
            me = st.get('me_mode', 'hbma')
            bs, R = int(st.get('block_size', 8)), int(st.get('target_region', 7))
            mode, fsz = st.get('filter_mode', 'none'), int(st.get('filter_size', 4))
            bidir = st.get('mci_mode', 'unidir') == 'bidir'
            cache = {}
            for i in range(len(want)):
                first, ratio = get_first_frame_idx_and_ratio(i, rr)
                dist = 1. - ratio
                if dist == 0.:
                    assert np.array_equal(frames[first], want[i])
                    continue
                a, b = frames[first], frames[first + 1]
                if first not in cache:
                    def est(x, y):
                        mv = oracle.fs(bs, R, x, y) if me == 'fs' else oracle.hbma(bs, R, 1, hbma_auto_steps((H, W)), 4, x, y)
                        return oracle.smooth(mode, mv, fsz)
                    cache = {first: (est(a, b), est(b, a) if bidir else None)}
                f_, b_ = cache[first]
                out = oracle.bidir_helper(a, b, f_, b_, np.float32(dist)) if bidir else oracle.unidir_helper(a, b, f_, np.float32(dist))
                assert np.array_equal(out, want[i]), (key, st, i)


def test_cfg1_360p_digests(oracle, golden):
    g = golden("cfg1_360p.npz")
    fr = make_sequence(360, 640, 2, seed=0)
    assert sha(fr) == str(g["frames_sha"])
    a, b = fr[0], fr[1]
    fwd, bwd = oracle.fs(8, 7, a, b), oracle.fs(8, 7, b, a)
    assert np.array_equal(fwd[::8, ::8], g["fs_fwd"]) and np.array_equal(bwd[::8, ::8], g["fs_bwd"])
    assert sha(oracle.bidir_helper(a, b, fwd, bwd, np.float32(0.5))) == str(g["bidir_fs_sha/0.5"])
    assert sha(oracle.unidir_helper(a, b, fwd, np.float32(0.39999998))) == str(g["unidir_fs_sha/0.39999998"])
    fwh = oracle.hbma(8, 7, 1, 2, 4, a, b)
    assert np.array_equal(fwh[::4, ::4], g["hbma_fwd"])


def test_fs_work_counts(oracle):
    """Algorithmic candidate counts quoted in SURVEY.md 8d (incl. the H-for-W bound quirk)."""
    assert oracle.fs_work(360, 640, 8, 7) == (622001, 39808064)
    assert oracle.fs_work(1080, 1920, 8, 16) == (27311000, 1747904000)
    n, ops = oracle.fs_work(2160, 3840, 16, 32)
    assert ops == n * 256 and abs(ops - 2.701e10) / 2.701e10 < 1e-3


def test_edge_cases(oracle):
    """Frames smaller than a block, single-row/column block grids, flat frames (all-tie search)."""
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (5, 7, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (5, 7, 3), dtype=np.uint8)
    mv = oracle.fs(8, 3, a, b)
    assert mv.shape == (5, 7, 3) and np.all(mv[:, :, :2] == 0)
    flat = np.full((24, 40, 3), 77, np.uint8)
    mv = oracle.fs(8, 4, flat, flat)
    assert np.all(mv == 0)          # SAD 0 everywhere: the zero vector wins the distance tie-break
    out = oracle.bidir_helper(flat, flat, mv, mv, np.float32(0.5))
    assert np.array_equal(out, flat)


def test_tss(oracle, golden):
    g = golden("tss_small.npz")
    for (H, W) in [(72, 96), (50, 70), (37, 53), (120, 40)]:
        fr = make_sequence(H, W, 2, seed=H)
        for bs, steps in ((8, 3), (8, 1), (4, 2), (16, 4), (8, 5)):
            assert np.array_equal(oracle.tss_blocks(bs, steps, fr[0], fr[1]), g[f"{H}x{W}/tss/{bs}_{steps}"]), (H, W, bs, steps)


# ---------------------------------------------------------------- bidir2 (MCI/Bidirectional_2.py)
def _b2_fields(g, k, me):
    """MV fields as the reference indexes them: dense (fs: stored on 8x8 blocks) or block level (hbma)."""
    fwd, bwd = g[f"{k}/{me}/fwd"], g[f"{k}/{me}/bwd"]
    if me == "fs":
        H, W = map(int, k.split("x"))
        fwd = np.repeat(np.repeat(fwd, 8, axis=0), 8, axis=1)[:H, :W]
        bwd = np.repeat(np.repeat(bwd, 8, axis=0), 8, axis=1)[:H, :W]
    return np.ascontiguousarray(fwd), np.ascontiguousarray(bwd), me == "fs"


def test_bidir2(oracle, golden):
    """IEWMC, Error_adaptive_combination, fill_holes / BDHI and the final uint8 frame against the reference."""
    g = golden("bidir2_small.npz")
    assert np.array_equal(oracle.weight_kern(8), g["weight_kern"])
    for (H, W) in SHAPES:
        fr = make_sequence(H, W, 2, seed=H + 1000)
        k = f"{H}x{W}"
        for me in ("fs", "hbma"):
            fwd, bwd, up = _b2_fields(g, k, me)
            for di, dist in enumerate((0.5, 0.39999998, 0.8)):
                out = oracle.bidir2_frame(fr[0], fr[1], fwd, bwd, dist, up)
                assert np.array_equal(out, g[f"{k}/{me}/out{di}"]), (k, me, dist)
            if H == 37:
                Ff, Ef = oracle.iewmc(fwd, fr[0], fr[1], 0.5, 16, 4, up)
                Fb, Eb = oracle.iewmc(bwd, fr[1], fr[0], 1 - 0.5, 16, 4, up)
                assert np.array_equal(Ff, g[f"{k}/{me}/Ff"]) and np.array_equal(Ef, g[f"{k}/{me}/Ef"])
                assert np.array_equal(oracle.eac(Ff, Ef, Fb, Eb), g[f"{k}/{me}/Fc"])
    blocks, want = g["fill_holes/in"], g["fill_holes/out"]
    for b, w in zip(blocks, want):
        assert np.array_equal(oracle.fill_holes(b), w, equal_nan=True)


def test_oracle_reports_undefined_wraparound(oracle):
    """A displaced index below -H is an out-of-bounds write in the reference (numba does not bounds-check the
    wraparound); the oracle refuses the case instead of guessing."""
    H, W = 16, 24
    a = np.zeros((H, W, 3), np.uint8)
    mv = np.zeros((H, W, 3), np.float32)
    mv[..., 0] = -40.0
    with pytest.raises(RuntimeError):
        oracle.unidir_helper(a, a, mv, np.float32(0.5))
    with pytest.raises(RuntimeError):
        oracle.bidir_helper(a, a, mv, mv, np.float32(0.5))
    mv[..., 0] = -31.0                                           # 0 + rint(-15.5) = -16 = -H: still a valid wrap
    oracle.unidir_helper(a, a, mv, np.float32(0.5))


def test_quality_oracle_known_answers():
    """The PSNR / SSIM restatement (oracle/quality.py; skimage itself is absent, parity unpinned) against the SSIM
    definition evaluated window by window and against closed-form answers."""
    import quality as Q
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (19, 23, 3), dtype=np.uint8)
    b = np.clip(a.astype(np.int32) + rng.integers(-30, 31, a.shape), 0, 255).astype(np.uint8)
    assert Q.structural_similarity(a, a) == 1.0
    assert Q.peak_signal_noise_ratio(a, a) == np.inf
    assert abs(Q.structural_similarity(a, b) - Q.ssim_by_definition(a, b)) < 1e-12
    c = np.full((16, 16, 3), 100, np.uint8); d = np.full((16, 16, 3), 110, np.uint8)
    assert abs(Q.peak_signal_noise_ratio(c, d) - 10 * np.log10(255.0 ** 2 / 100.0)) < 1e-12
    # two constant frames: variances vanish, S = (2 ux uy + C1) / (ux^2 + uy^2 + C1)
    C1 = (0.01 * 255) ** 2
    assert abs(Q.structural_similarity(c, d) - (2 * 100 * 110 + C1) / (100 ** 2 + 110 ** 2 + C1)) < 1e-12
    with pytest.raises(ValueError):
        Q.structural_similarity(a[:6], b[:6])


def test_prefilter_bound_holds_for_the_oracle_winner(oracle):
    """The pruning rule of the two-phase full search (csrc/fs8.cu), checked on the CPU against the oracle: with luma rounded
    to 8 bits, A = sum |L8a - L8b| stays within bs^2 of S1000 / 1000 for every candidate, and the reference's winner always
    satisfies A(winner) <= min A + 2 bs^2 + 1 -- on texture, noise, two grey levels (ties, flagged sums) and flat frames."""
    rng = np.random.default_rng(21)
    H, W = 72, 104

    def l1000(f):
        f = f.astype(np.int64)
        return 299 * f[..., 0] + 587 * f[..., 1] + 114 * f[..., 2]

    two = np.repeat(rng.integers(0, 2, (2, H, W, 1), dtype=np.uint8) * 50 + 20, 3, axis=3)
    flat = np.empty((2, H, W, 3), np.uint8); flat[0] = 77; flat[1] = 80
    cases = [make_sequence(H, W, 2, seed=5), rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8), two, flat]
    for fr in cases:
        for bs, R in ((8, 7), (16, 5)):
            want = oracle.fs_blocks(bs, R, fr[0], fr[1])
            La, Lb = l1000(fr[0]), l1000(fr[1])
            pad = lambda x: np.pad(x, ((0, bs), (0, bs)))                  # the reference's zero padding
            La, Lb = pad(La), pad(Lb)
            A8, B8 = (La + 500) // 1000, (Lb + 500) // 1000
            n = bs * bs
            for br in range(want.shape[0]):
                for bc in range(want.shape[1]):
                    if not np.isfinite(want[br, bc, 2]):
                        continue
                    r, c = br * bs, bc * bs
                    tmr = r + 1 if r + bs >= H else H - bs
                    tmc = c + 1 if c + bs >= H else W - bs                 # full_search.py:37 compares with H
                    rows = range(max(r - R, 0), min(tmr, r + R + 1))
                    cols = range(max(c - R, 0), min(tmc, c + R + 1))
                    a_min, a_win = None, None
                    for tr in rows:
                        for tc in cols:
                            S = int(np.abs(La[r:r + bs, c:c + bs] - Lb[tr:tr + bs, tc:tc + bs]).sum())
                            A = int(np.abs(A8[r:r + bs, c:c + bs] - B8[tr:tr + bs, tc:tc + bs]).sum())
                            assert abs(S / 1000.0 - A) <= n
                            a_min = A if a_min is None else min(a_min, A)
                            if tr - r == want[br, bc, 0] and tc - c == want[br, bc, 1]:
                                a_win = A
                    assert a_win is not None and a_win <= a_min + 2 * n + 1, (bs, R, br, bc)


def test_oracle_vs_reference_hashes_on_fuzz_geometry(oracle, golden):
    """82 seeded adversarial MCI inputs (oracle/fuzz_cases.py: wraps, out-of-frame targets, fractional vectors, mismatched
    MV / SAD cells, ties, infinite SADs) that were run through the unmodified reference (oracle/validate_fuzz_vs_reference.py
    --write-golden): the oracle reproduces every frame, and refuses exactly the cases the reference leaves undefined."""
    import fuzz_cases as F
    g = golden("fuzz_reference_sha.npz")
    n = 0
    for it, a, b, f_, b_, dist in F.mci_cases(int(g["seed"]), int(g["rounds"])):
        if f"{it}/undefined" in g:
            with pytest.raises(RuntimeError):
                oracle.unidir_helper(a, b, f_, dist); oracle.bidir_helper(a, b, f_, b_, dist)
            continue
        assert F.sha(oracle.unidir_helper(a, b, f_, dist)) == str(g[f"{it}/unidir"]), it
        assert F.sha(oracle.bidir_helper(a, b, f_, b_, dist)) == str(g[f"{it}/bidir"]), it
        n += 1
    assert n >= 80


def test_oracle_bidir2_vs_reference_hashes_on_fuzz_fields(oracle, golden):
    """64 seeded random bidir2 inputs run through the unmodified reference (IEWMC x 2, error-adaptive combination, BDHI,
    unpad): the oracle reproduces every frame and fails exactly where the reference raises (a block shifted out of the pad)."""
    import fuzz_cases as F
    g = golden("fuzz_reference_sha.npz")
    n = 0
    for it, a, b, f2, bw2, dd, up in F.bidir2_cases(int(g["b2_seed"]), int(g["b2_rounds"])):
        if f"b2/{it}/out_of_pad" in g:
            with pytest.raises(RuntimeError):
                oracle.bidir2_frame(a, b, f2, bw2, dd, up)
            continue
        assert F.sha(oracle.bidir2_frame(a, b, f2, bw2, dd, up)) == str(g[f"b2/{it}/out"]), it
        n += 1
    assert n >= 45


def test_oracle_me_vs_reference_hashes_on_random_parameters(oracle, golden):
    """60 seeded frame pairs with random full-search / TSS / HBMA (sad, ssd, upscale or not) / smoothing parameters, run
    through the unmodified reference: the oracle reproduces every field.  HBMA cases with a single block row or column at
    some level are left out: the reference reads a neighbour out of bounds there (oracle.hbma_undefined_in_reference)."""
    import fuzz_cases as F
    g = golden("fuzz_reference_sha.npz")
    n = 0
    for it, fr, q, mv in F.me_cases(int(g["me_seed"]), int(g["me_rounds"])):
        a, b = fr[0], fr[1]
        calls = {"fs": lambda: oracle.fs(q['bs'], q['R'], a, b),
                 "tss": lambda: oracle.tss(q['bs'], q['steps'], a, b),
                 "hbma": lambda: oracle.hbma(q['hb'], q['win'], q['sub'], q['hsteps'], q['mbs'], a, b, q['cost'], q['up'])}
        for mode in ("mean", "median", "weighted"):
            calls[f"smooth_{mode}"] = lambda mode=mode: oracle.smooth(mode, mv, q['fsz'])
        for name, f in calls.items():
            key = f"me/{it}/{name}"
            if key + "/undefined" in g:
                assert name == "hbma" and oracle.hbma_undefined_in_reference(a.shape[0], a.shape[1], q['hb'], q['hsteps'], q['mbs'])
                continue
            if key + "/raises" in g:
                continue
            assert F.sha(np.asarray(f(), np.float32)) == str(g[key]), (it, name, q)
            n += 1
    assert n >= 300


def test_oracle_middlebury_pairs_vs_reference(oracle, golden, middlebury):
    """All 12 Middlebury pairs of the reference's own benchmark set (640x480, 584x388, 420x380: partial blocks on
    real image content) through the oracle chain that mirrors `MEMCI(2, **settings).get_benchmark_frame(f10, f11)`
    (src/Benchmark.py:44, base.py:136-157: dist = 0.5) against the digests of the unmodified reference's outputs
    for fs / hbma x unidir / bidir / bidir2 and fs +-16 + median smoothing (tests/golden/middlebury_ref_sha.npz)."""
    g = golden("middlebury_ref_sha.npz")
    half = np.float32(0.5)
    assert len(middlebury) == 12
    for name, (a, b, _) in middlebury.items():
        H, W, _c = a.shape
        assert sha(np.stack([a, b])) == str(g[f"{name}/frames_sha"]), "PNG decode differs from the fixture generator's"
        fwd_b, bwd_b = oracle.fs_blocks(8, 7, a, b), oracle.fs_blocks(8, 7, b, a)
        assert np.array_equal(fwd_b, g[f"{name}/fs_fwd"]) and np.array_equal(bwd_b, g[f"{name}/fs_bwd"]), name
        fwd, bwd = oracle.expand_dense(fwd_b, 8, H, W), oracle.expand_dense(bwd_b, 8, H, W)
        steps = oracle.hbma_auto_steps(H, W)
        fwh, bwh = oracle.hbma(8, 7, 1, steps, 4, a, b), oracle.hbma(8, 7, 1, steps, 4, b, a)
        f16 = oracle.smooth('median', oracle.fs(8, 16, a, b), 4)
        b16 = oracle.smooth('median', oracle.fs(8, 16, b, a), 4)
        got = {
            "fs_unidir": oracle.unidir_helper(a, b, fwd, half),
            "fs_bidir": oracle.bidir_helper(a, b, fwd, bwd, half),
            "fs_bidir2": oracle.bidir2_frame(a, b, fwd, bwd, 0.5, True),
            "hbma_unidir": oracle.unidir_helper(a, b, fwh, half),
            "hbma_bidir": oracle.bidir_helper(a, b, fwh, bwh, half),
            "hbma_bidir2": oracle.bidir2_frame(a, b, fwh, bwh, 0.5, True),
            "fs16_bidir_median": oracle.bidir_helper(a, b, f16, b16, half),
        }
        for key, out in got.items():
            assert sha(out) == str(g[f"{name}/{key}"]), (name, key)
