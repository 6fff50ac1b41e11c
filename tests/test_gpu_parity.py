"""Parity proper: the CUDA path, called through the C ABI (ctypes) and through the mirrored
reference interface, against the CPU oracle on seeded inputs and against the committed golden
fixtures (outputs of the unmodified reference).  Bar: bit-exact (np.array_equal) everywhere."""
import hashlib
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from interpolatory_b200.synth import make_sequence  # noqa: E402


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest() + f":{a.dtype}:{'x'.join(map(str, a.shape))}"


def blocks_of(dense, bs):
    return np.ascontiguousarray(dense[::bs, ::bs])


@pytest.fixture(scope="module")
def gpu(native_lib):
    import interpolatory_b200 as pkg
    return pkg


SHAPES = [(72, 96), (50, 70), (37, 53)]
FS_CFGS = [(8, 7), (8, 3), (4, 5), (16, 9), (8, 16)]
HB_CFGS = [(8, 7, 1, 2, 4), (8, 3, 1, 1, 4), (8, 5, 2, 3, 2), (4, 2, 1, 1, 4), (16, 4, 1, 1, 4)]
DISTS = [0.5, 0.39999998, 0.8, 0.19999999, 0.6]


# ---------------------------------------------------------------- golden fixtures (reference outputs)
@pytest.mark.parametrize("shape", SHAPES)
def test_fs_vs_golden(gpu, golden, shape):
    from interpolatory_b200.ME.full_search import get_motion_vectors as fs
    g = golden("functions_small.npz")
    k = f"{shape[0]}x{shape[1]}"
    a, b = g[f"{k}/a"], g[f"{k}/b"]
    for bs, R in FS_CFGS:
        mv = fs(bs, R, a, b)
        assert mv.dtype == np.float32 and mv.shape == a.shape
        assert np.array_equal(blocks_of(mv, bs), g[f"{k}/fs/{bs}_{R}"]), (shape, bs, R)
    assert np.array_equal(fs(8, 7, a, b), g[f"{k}/fs_dense/8_7"])


@pytest.mark.parametrize("shape", SHAPES)
def test_hbma_vs_golden(gpu, golden, shape):
    from interpolatory_b200.ME.hbma import get_motion_vectors as hbma
    g = golden("functions_small.npz")
    k = f"{shape[0]}x{shape[1]}"
    a, b = g[f"{k}/a"], g[f"{k}/b"]
    for cfg in HB_CFGS:
        mv = hbma(*cfg, a, b, 'sad', False)
        ref = g[f"{k}/hbma/{'_'.join(map(str, cfg))}"]
        assert mv.shape == ref.shape and np.array_equal(mv, ref), (shape, cfg)
    assert np.array_equal(hbma(8, 7, 1, 2, 4, a, b), g[f"{k}/hbma_dense/8_7_1_2_4"])


@pytest.mark.parametrize("shape", SHAPES)
def test_pyramid_vs_golden(gpu, golden, shape):
    from interpolatory_b200.device import get_context
    g = golden("functions_small.npz")
    k = f"{shape[0]}x{shape[1]}"
    ctx = get_context(*shape)
    ctx.upload_frame(0, g[f"{k}/a"])
    for lvl in (1, 2):
        assert np.array_equal(ctx.pyramid_level(0, lvl), g[f"{k}/pyr/{lvl}"]), (shape, lvl)


@pytest.mark.parametrize("shape", SHAPES)
def test_smooth_vs_golden(gpu, golden, oracle, shape):
    from interpolatory_b200.smoothing import smooth, mean_filter, median_filter, weighted_mean_filter
    g = golden("functions_small.npz")
    k = f"{shape[0]}x{shape[1]}"
    a, b = g[f"{k}/a"], g[f"{k}/b"]
    fields = {"fs": g[f"{k}/fs_dense/8_7"], "hbma": g[f"{k}/hbma_dense/8_7_1_2_4"]}
    for name, f in (("mean", mean_filter), ("median", median_filter), ("weighted", weighted_mean_filter)):
        for fsz in (4, 8, 3):
            for nm, mv in fields.items():
                out = smooth(f, mv, fsz)
                assert np.array_equal(out[:, :, :2], g[f"{k}/smooth/{name}_{fsz}/{nm}"]), (shape, name, fsz, nm)
                assert np.array_equal(out[:, :, 2], mv[:, :, 2])
    assert smooth(None, fields["fs"], 4) is fields["fs"]


@pytest.mark.parametrize("shape", SHAPES)
def test_mci_vs_golden(gpu, golden, oracle, shape):
    from interpolatory_b200.MCI.Unidirectional import helper as uni
    from interpolatory_b200.MCI.Bidirectional import helper as bi
    g = golden("functions_small.npz")
    k = f"{shape[0]}x{shape[1]}"
    H, W = shape
    a, b = g[f"{k}/a"], g[f"{k}/b"]
    fwd = g[f"{k}/fs_dense/8_7"]
    bwd = oracle.fs(8, 7, b, a)
    fwh = g[f"{k}/hbma_dense/8_7_1_2_4"]
    bwh = oracle.hbma(8, 7, 1, 2, 4, b, a)
    for dist in DISTS:
        d = np.float32(dist)
        assert np.array_equal(uni(a, b, fwd, d), g[f"{k}/unidir/fs/{dist}"]), (shape, dist)
        assert np.array_equal(bi(a, b, fwd, bwd, d), g[f"{k}/bidir/fs/{dist}"]), (shape, dist)
        assert np.array_equal(uni(a, b, fwh, d), g[f"{k}/unidir/hbma/{dist}"]), (shape, dist)
        assert np.array_equal(bi(a, b, fwh, bwh, d), g[f"{k}/bidir/hbma/{dist}"]), (shape, dist)
    # adversarial block-constant random fields: wraps, out-of-frame targets, fractional vectors
    for trial in range(3):
        gsz = int(g[f"{k}/rand/{trial}/g"])
        f_ = np.kron(g[f"{k}/rand/{trial}/fwd"], np.ones((gsz, gsz, 1), np.float32))[:H, :W].copy()
        b_ = np.kron(g[f"{k}/rand/{trial}/bwd"], np.ones((gsz, gsz, 1), np.float32))[:H, :W].copy()
        for dist in (0.5, 0.39999998, 0.8):
            d = np.float32(dist)
            assert np.array_equal(uni(a, b, f_, d), g[f"{k}/rand/{trial}/unidir/{dist}"]), (shape, trial, dist)
            assert np.array_equal(bi(a, b, f_, b_, d), g[f"{k}/rand/{trial}/bidir/{dist}"]), (shape, trial, dist)


def test_middlebury_crops_vs_golden(gpu, golden):
    from interpolatory_b200.ME.full_search import get_motion_vectors as fs
    from interpolatory_b200.ME.hbma import get_motion_vectors as hbma
    from interpolatory_b200.MCI.Unidirectional import helper as uni
    from interpolatory_b200.MCI.Bidirectional import helper as bi
    g = golden("middlebury_crops.npz")
    for name in ("RubberWhale", "Dimetrodon"):
        a, b = g[f"{name}/a"], g[f"{name}/b"]
        fwd, bwd = fs(8, 7, a, b), fs(8, 7, b, a)
        assert np.array_equal(blocks_of(fwd, 8), g[f"{name}/fs_fwd"])
        assert np.array_equal(blocks_of(bwd, 8), g[f"{name}/fs_bwd"])
        assert np.array_equal(hbma(8, 7, 1, 1, 4, a, b, 'sad', False), g[f"{name}/hbma_fwd"])
        assert np.array_equal(bi(a, b, fwd, bwd, np.float32(0.5)), g[f"{name}/bidir"])
        assert np.array_equal(uni(a, b, fwd, np.float32(0.5)), g[f"{name}/unidir"])


def test_class_api_vs_golden(gpu, golden):
    """MEMCI(...) factory + get_interpolated_frame over a whole clip: dist cycle, pass-through
    frames, MV cache, auto pyramid depth, smoothing -- against the reference classes' output."""
    import ast
    from interpolatory_b200 import MEMCI, ArrayVideoStream
    g = golden("e2e.npz")
    keys = sorted({k.split("/")[0] for k in g.files})
    for key in keys:
        H, W = map(int, key.split("_")[0].split("x"))
        fps_in, fps_out = int(key.split("_")[1]), int(key.split("_")[2])
        seed = int(g[f"{key}/seed"])
        nsrc = {72: 4, 56: 3, 264: 2}[H]
        frames = make_sequence(H, W, nsrc, seed=seed)
        assert sha(frames) == str(g[f"{key}/frames_sha"]), "synthetic generator drifted from the fixture"
        for si in range(7):
            if f"{key}/set{si}/settings" not in g.files:
                continue
            st = dict(ast.literal_eval(str(g[f"{key}/set{si}/settings"])))
            it = MEMCI(fps_out, **st)
            it.set_video_stream(ArrayVideoStream(list(frames), fps_in))
            want = [str(s) for s in g[f"{key}/set{si}/out_sha"]]
            for i, w in enumerate(want):
                out = np.asarray(it.get_interpolated_frame(i))
                assert sha(out) == w, (key, st, i)
            it.close()


def test_cfg1_360p_vs_golden(gpu, golden):
    from interpolatory_b200.ME.full_search import get_motion_vectors as fs
    from interpolatory_b200.ME.hbma import get_motion_vectors as hbma
    from interpolatory_b200.MCI.Unidirectional import helper as uni
    from interpolatory_b200.MCI.Bidirectional import helper as bi
    g = golden("cfg1_360p.npz")
    fr = make_sequence(360, 640, 2, seed=0)
    assert sha(fr) == str(g["frames_sha"])
    a, b = fr[0], fr[1]
    fwd, bwd = fs(8, 7, a, b), fs(8, 7, b, a)
    assert np.array_equal(blocks_of(fwd, 8), g["fs_fwd"])
    assert np.array_equal(blocks_of(bwd, 8), g["fs_bwd"])
    fwh, bwh = hbma(8, 7, 1, 2, 4, a, b), hbma(8, 7, 1, 2, 4, b, a)
    assert np.array_equal(blocks_of(fwh, 4), g["hbma_fwd"])
    assert np.array_equal(blocks_of(bwh, 4), g["hbma_bwd"])
    for dist in (0.5, 0.39999998):
        d = np.float32(dist)
        assert sha(bi(a, b, fwd, bwd, d)) == str(g[f"bidir_fs_sha/{dist}"])
        assert sha(uni(a, b, fwd, d)) == str(g[f"unidir_fs_sha/{dist}"])
        assert sha(bi(a, b, fwh, bwh, d)) == str(g[f"bidir_hbma_sha/{dist}"])


# ---------------------------------------------------------------- oracle on fresh seeded inputs
@pytest.mark.parametrize("H,W,bs,R", [(120, 200, 8, 16), (97, 131, 8, 7), (64, 256, 16, 12), (136, 80, 8, 5),
                                      (90, 160, 24, 6), (48, 64, 6, 4), (40, 300, 8, 20)])
def test_fs_vs_oracle(gpu, oracle, H, W, bs, R):
    from interpolatory_b200.device import get_context
    fr = make_sequence(H, W, 2, seed=H * 7 + W)
    ctx = get_context(H, W)
    ctx.upload_frame(0, fr[0]); ctx.upload_frame(1, fr[1])
    for (s, t) in ((0, 1), (1, 0)):
        ctx.me_fs(s, t, bs, R, 0)
        _, vec, _, sad = ctx.download_mv_blocks(0)
        ref = oracle.fs_blocks(bs, R, fr[s], fr[t])
        assert np.array_equal(vec, ref[:, :, :2]), (H, W, bs, R)
        assert np.array_equal(sad, ref[:, :, 2])
        n_cand, px_ops, n_redo = ctx.fs_stats()
        assert (n_cand, px_ops) == oracle.fs_work(H, W, bs, R)


def test_fs_truncation_fixup_exercised(gpu, oracle):
    """Frames that differ by a pure grey offset make sum|dL| an exact integer for every
    candidate: the float64 truncation decides K vs K-1 and the redo pass must run."""
    from interpolatory_b200.device import get_context
    rng = np.random.default_rng(5)
    H, W = 64, 96
    a = rng.integers(0, 200, size=(H, W, 3)).astype(np.uint8)
    a[:, :] = a[:, :, :1]                      # grey image: dL1000 = 1000 * dGrey for every pixel pair
    b = (np.roll(a, (1, -2), (0, 1)).astype(np.int32) + 3).astype(np.uint8)
    ctx = get_context(H, W)
    ctx.upload_frame(0, a); ctx.upload_frame(1, b)
    ctx.me_fs(0, 1, 8, 7, 0)
    _, vec, _, sad = ctx.download_mv_blocks(0)
    n_redo = ctx.fs_stats()[2]
    ref = oracle.fs_blocks(8, 7, a, b)
    assert n_redo > 0
    assert np.array_equal(vec, ref[:, :, :2]) and np.array_equal(sad, ref[:, :, 2])


@pytest.mark.parametrize("H,W,cfg", [(130, 210, (8, 7, 1, 2, 4)), (300, 420, (8, 7, 1, 2, 4)), (75, 99, (8, 4, 2, 1, 2)),
                                     (128, 128, (16, 5, 1, 2, 4))])
def test_hbma_vs_oracle(gpu, oracle, H, W, cfg):
    from interpolatory_b200.ME.hbma import get_motion_vectors as hbma
    fr = make_sequence(H, W, 2, seed=H + W)
    assert np.array_equal(hbma(*cfg, fr[0], fr[1], 'sad', False), oracle.hbma(*cfg, fr[0], fr[1], 'sad', False))
    assert np.array_equal(hbma(*cfg, fr[1], fr[0], 'ssd', False), oracle.hbma(*cfg, fr[1], fr[0], 'ssd', False))


@pytest.mark.parametrize("H,W", [(120, 200), (97, 131)])
@pytest.mark.parametrize("mode", ["none", "mean", "median", "weighted"])
def test_pipeline_vs_oracle(gpu, oracle, H, W, mode):
    """Whole pair through the device-resident path (no dense fields on the host) vs the oracle chain."""
    from interpolatory_b200.device import get_context
    from interpolatory_b200._native import PairParams
    fr = make_sequence(H, W, 2, seed=3 * H + W)
    a, b = fr[0], fr[1]
    ctx = get_context(H, W)
    ctx.upload_frame(0, a); ctx.upload_frame(1, b)
    for me in ("fs", "hbma"):
        p = PairParams(0 if me == "fs" else 1, 8, 7, 1, 2, 4, {"none": 0, "mean": 1, "median": 2, "weighted": 3}[mode], 4, 1)
        ctx.estimate_pair(p, 0, 1, 0, 1)
        if me == "fs":
            f_, b_ = oracle.fs(8, 7, a, b), oracle.fs(8, 7, b, a)
        else:
            f_, b_ = oracle.hbma(8, 7, 1, 2, 4, a, b), oracle.hbma(8, 7, 1, 2, 4, b, a)
        f_, b_ = oracle.smooth(mode, f_, 4), oracle.smooth(mode, b_, 4)
        assert np.array_equal(ctx.download_mv_dense(0), f_)
        assert np.array_equal(ctx.download_mv_dense(1), b_)
        for dist in (0.5, 0.19999999):
            ctx.mci_bidir(0, 1, 0, 1, np.float32(dist), 7)
            assert np.array_equal(ctx.download_frame(7), oracle.bidir_helper(a, b, f_, b_, np.float32(dist)))
            ctx.mci_unidir(0, 1, 0, np.float32(dist), 7)
            assert np.array_equal(ctx.download_frame(7), oracle.unidir_helper(a, b, f_, np.float32(dist)))


# ---------------------------------------------------------------- spatial band split (8K mode)
@pytest.mark.parametrize("H,W,bs,R,n_bands", [(360, 640, 8, 7, 2), (360, 640, 8, 16, 4), (288, 512, 16, 32, 3), (200, 320, 8, 16, 8)])
def test_band_split_matches_single_gpu(gpu, oracle, H, W, bs, R, n_bands):
    """One pair split into horizontal bands with frame-row and MV-row halo exchange (emulated on one
    GPU with one context per band and device-to-device copies) == the single-context result == oracle."""
    from interpolatory_b200.tiling import interpolate_pair_emulated
    from interpolatory_b200.device import get_context
    fr = make_sequence(H, W, 2, seed=H + n_bands)
    a, b = fr[0], fr[1]
    # make the wrap link matter: strong upward motion at the top of the frame
    b = b.copy(); b[: 3 * bs] = np.roll(a, -min(R, 6), axis=0)[: 3 * bs]
    ctx = get_context(H, W)
    ctx.upload_frame(0, a); ctx.upload_frame(1, b)
    ctx.me_fs(0, 1, bs, R, 0); ctx.me_fs(1, 0, bs, R, 1)
    for dist in (0.5, 0.8):
        ctx.mci_bidir(0, 1, 0, 1, np.float32(dist), 7)
        want = ctx.download_frame(7)
        got, plan = interpolate_pair_emulated(a, b, bs, R, np.float32(dist), n_bands)
        assert np.array_equal(got, want), (H, W, bs, R, n_bands, dist)
    f_, b_ = oracle.fs(bs, R, a, b), oracle.fs(bs, R, b, a)
    assert np.array_equal(want, oracle.bidir_helper(a, b, f_, b_, np.float32(0.8)))
    fbytes, mbytes = plan.bytes_moved()
    assert fbytes > 0 and mbytes > 0


# ---------------------------------------------------------------- BASELINE.json's full sizes
def test_cfg2_1080p_full_size_vs_oracle(gpu, oracle):
    """cfg2 at full size (1080p, fs 8x8 +-16, bidir): motion fields and the interpolated frame against
    the oracle (OpenMP on the host cores: a few seconds)."""
    from interpolatory_b200.device import DeviceContext
    H, W, bs, R = 1080, 1920, 8, 16
    fr = make_sequence(H, W, 2, seed=42)
    a, b = fr[0], fr[1]
    ctx = DeviceContext(H, W)
    try:
        ctx.upload_frame(0, a); ctx.upload_frame(1, b)
        ctx.me_fs(0, 1, bs, R, 0); ctx.me_fs(1, 0, bs, R, 1)
        assert ctx.fs_stats()[:2] == (27311000, 1747904000)
        fwd, bwd = oracle.fs_blocks(bs, R, a, b), oracle.fs_blocks(bs, R, b, a)
        for slot, ref in ((0, fwd), (1, bwd)):
            _, vec, _, sad = ctx.download_mv_blocks(slot)
            assert np.array_equal(vec, ref[:, :, :2]) and np.array_equal(sad, ref[:, :, 2])
        ctx.mci_bidir(0, 1, 0, 1, np.float32(0.5), 7)
        got = ctx.download_frame(7)
        want = oracle.bidir_helper(a, b, oracle.expand_dense(fwd, bs, H, W), oracle.expand_dense(bwd, bs, H, W), np.float32(0.5))
        assert np.array_equal(got, want)
        ctx.mci_unidir(0, 1, 0, np.float32(0.5), 7)
        assert np.array_equal(ctx.download_frame(7), oracle.unidir_helper(a, b, oracle.expand_dense(fwd, bs, H, W), np.float32(0.5)))
    finally:
        ctx.close()


def test_tiled_kernel_equals_generic_kernel_1080p(gpu):
    """Two independent code paths (TMA-tiled register kernel vs warp-per-block exact kernel) agree on
    every block of a 1080p pair."""
    import os
    from interpolatory_b200.device import DeviceContext
    H, W = 1080, 1920
    fr = make_sequence(H, W, 2, seed=77)
    res = []
    for force in (False, True):
        if force:
            os.environ["MEMCI_FS_FORCE_GENERIC"] = "1"
        try:
            ctx = DeviceContext(H, W)
            ctx.upload_frame(0, fr[0]); ctx.upload_frame(1, fr[1])
            ctx.me_fs(0, 1, 8, 16, 0)
            res.append(ctx.download_mv_blocks(0))
            ctx.close()
        finally:
            os.environ.pop("MEMCI_FS_FORCE_GENERIC", None)
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][3], res[1][3])


def _full_size_case(gpu, monkeypatch, name, prefilter, refs):
    """One seeded full-size pair (interpolatory_b200.synth.make_pair_tiled, the generator oracle/gen_fullsize_sha.py
    used) through the CUDA path with the two-phase search pinned on or off; `refs` maps field / frame names to
    either a digest (committed oracle output) or an array (oracle run live)."""
    from interpolatory_b200.device import DeviceContext
    from interpolatory_b200.synth import make_pair_tiled
    H, W, bs, R, seed = {"cfg4": (2160, 3840, 16, 32, 4), "cfg5": (4320, 7680, 16, 32, 5)}[name]
    monkeypatch.setenv("MEMCI_FS_PREFILTER", "1" if prefilter else "0")
    a, b = make_pair_tiled(H, W, seed)
    assert sha(np.stack([a, b])) == refs["frames_sha"], "synthetic generator drifted from the fixture"
    ctx = DeviceContext(H, W)
    try:
        ctx.upload_frame(0, a); ctx.upload_frame(1, b)
        ctx.me_fs_pair(0, 1, bs, R, 0, 1)
        n_surv, overflow = ctx.fs_prefilter_stats()
        if prefilter:                    # the 16 x 16 / +-32 instantiation really ran and kept its list
            assert n_surv > 0 and not overflow, (n_surv, overflow)
        else:
            assert n_surv == 0 and not overflow
        assert tuple(ctx.fs_stats()[:2]) == tuple(2 * int(x) for x in refs["work"])
        for slot, key in ((0, "fwd"), (1, "bwd")):
            _, vec, _, sad = ctx.download_mv_blocks(slot)
            got = np.concatenate([vec, sad[..., None]], axis=2)
            ref = refs[key]
            assert (sha(got) == ref) if isinstance(ref, str) else np.array_equal(got, ref), (name, prefilter, key)
        ctx.mci_bidir(0, 1, 0, 1, np.float32(0.5), 7)
        got = ctx.download_frame(7)
        assert (sha(got) == refs["bidir"]) if isinstance(refs["bidir"], str) else np.array_equal(got, refs["bidir"])
        ctx.mci_unidir(0, 1, 0, np.float32(0.5), 7)
        assert sha(ctx.download_frame(7)) == refs["unidir_sha"]
    finally:
        ctx.close()


def test_cfg4_4k_full_size_vs_oracle(gpu, oracle, golden, monkeypatch):
    """cfg4 at full size (2160 x 3840, fs 16 x 16 +-32, bidir): both motion fields and the interpolated frame
    against the oracle run LIVE on the host cores (2 x 2.7e10 SAD px-ops: ~25 s with OpenMP), with the two-phase
    search on and off; the live oracle outputs must also equal the committed digests (same oracle, other machine)."""
    from interpolatory_b200.synth import make_pair_tiled
    g = golden("fullsize_sha.npz")
    H, W, bs, R = 2160, 3840, 16, 32
    a, b = make_pair_tiled(H, W, 4)
    fwd, bwd = oracle.fs_blocks(bs, R, a, b), oracle.fs_blocks(bs, R, b, a)
    assert sha(fwd) == str(g["cfg4/fwd_sha"]) and sha(bwd) == str(g["cfg4/bwd_sha"])
    frame = oracle.bidir_helper(a, b, oracle.expand_dense(fwd, bs, H, W), oracle.expand_dense(bwd, bs, H, W), np.float32(0.5))
    assert sha(frame) == str(g["cfg4/bidir_sha"])
    refs = {"frames_sha": str(g["cfg4/frames_sha"]), "work": g["cfg4/work"], "fwd": fwd, "bwd": bwd, "bidir": frame,
            "unidir_sha": str(g["cfg4/unidir_sha"])}
    for prefilter in (True, False):
        _full_size_case(gpu, monkeypatch, "cfg4", prefilter, refs)


@pytest.mark.parametrize("prefilter", [True, False])
def test_cfg5_8k_full_size_vs_oracle_digests(gpu, golden, monkeypatch, prefilter):
    """cfg5 at full size (4320 x 7680 single pair, fs 16 x 16 +-32: 129 600 blocks, 1 % under the 2^17 block
    field of the survivor list) against the digests of the oracle's fields and frames (oracle/gen_fullsize_sha.py:
    2 x 1.09e11 SAD px-ops, minutes of host time, so generated once and committed), two-phase search on and off."""
    g = golden("fullsize_sha.npz")
    refs = {"frames_sha": str(g["cfg5/frames_sha"]), "work": g["cfg5/work"], "fwd": str(g["cfg5/fwd_sha"]),
            "bwd": str(g["cfg5/bwd_sha"]), "bidir": str(g["cfg5/bidir_sha"]), "unidir_sha": str(g["cfg5/unidir_sha"])}
    _full_size_case(gpu, monkeypatch, "cfg5", prefilter, refs)


def test_middlebury_pairs_vs_reference_digests(gpu, golden, middlebury):
    """The reference's own 12 Middlebury pairs through the drop-in class API, `MEMCI(2, **settings)
    .get_benchmark_frame(frame10, frame11)` (the call src/Benchmark.py:44 makes), against the digests of the
    unmodified reference's output frames: fs / hbma x unidir / bidir / bidir2, and fs +-16 with median smoothing;
    plus the block-level full-search fields of both directions."""
    import ast
    from interpolatory_b200 import MEMCI
    from interpolatory_b200.ME.full_search import get_motion_vectors as fs
    g = golden("middlebury_ref_sha.npz")
    settings = {k: dict(v) for k, v in ast.literal_eval(str(g["settings"])).items()}
    assert len(middlebury) == 12 and len(settings) == 7
    for name, (a, b, _) in middlebury.items():
        assert sha(np.stack([a, b])) == str(g[f"{name}/frames_sha"])
        assert np.array_equal(blocks_of(fs(8, 7, a, b), 8), g[f"{name}/fs_fwd"]), name
        assert np.array_equal(blocks_of(fs(8, 7, b, a), 8), g[f"{name}/fs_bwd"]), name
        for key, st in settings.items():
            it = MEMCI(2, **st)
            out = np.asarray(it.get_benchmark_frame(a, b))
            it.close()
            assert sha(out) == str(g[f"{name}/{key}"]), (name, key)


@pytest.mark.parametrize("H,W,bs,R", [(2160, 3840, 16, 32), (4320, 7680, 16, 32)])
def test_full_size_properties_4k_8k(gpu, oracle, H, W, bs, R):
    """Size-independent properties at cfg4 / cfg5 sizes (the oracle would take minutes here):
    identical frames -> zero field, zero SAD, output == input; a pure translation inside the search
    range -> that vector with SAD 0 on every interior block, backward field = minus forward field;
    candidate count == the reference's bounds."""
    from interpolatory_b200.device import DeviceContext
    base = make_sequence(H // 4, W // 4, 1, seed=H)[0]
    rng = np.random.default_rng(H)
    a = np.ascontiguousarray(np.tile(base, (4, 4, 1)))
    a = (a.astype(np.int16) + rng.integers(-20, 21, size=(H, W, 1))).clip(0, 255).astype(np.uint8)   # break the 4x4 periodicity
    dy, dx = 11, -23
    b = np.roll(a, (dy, dx), axis=(0, 1))
    ctx = DeviceContext(H, W)
    try:
        ctx.upload_frame(0, a); ctx.upload_frame(1, a)
        ctx.me_fs(0, 1, bs, R, 0); ctx.me_fs(1, 0, bs, R, 1)
        assert ctx.fs_stats()[:2] == oracle.fs_work(H, W, bs, R)
        _, vec, _, sad = ctx.download_mv_blocks(0)
        assert not vec.any() and not sad.any()
        ctx.mci_bidir(0, 1, 0, 1, np.float32(0.5), 7)
        assert np.array_equal(ctx.download_frame(7), a) and ctx.mci_stats() == 0
        ctx.upload_frame(1, b)
        ctx.me_fs(0, 1, bs, R, 0); ctx.me_fs(1, 0, bs, R, 1)
        _, vf, _, sf = ctx.download_mv_blocks(0)
        _, vb, _, sb = ctx.download_mv_blocks(1)
        m = R // bs + 3                                  # stay clear of the frame edges and of the column-bound quirk
        nbc_ok = (H - bs) // bs - 1                      # columns where the reference searches dx > 0 (full_search.py:37)
        inner_f = vf[m:-m, m:min(nbc_ok, vf.shape[1] - m)]
        assert np.all(inner_f[..., 0] == dy) and np.all(inner_f[..., 1] == dx) and not sf[m:-m, m:min(nbc_ok, vf.shape[1] - m)].any()
        inner_b = vb[m:-m, m:min(nbc_ok, vb.shape[1] - m)]
        assert np.all(inner_b[..., 0] == -dy) and np.all(inner_b[..., 1] == -dx)
    finally:
        ctx.close()


# ---------------------------------------------------------------- three-step search
def test_tss_vs_golden_and_oracle(gpu, golden, oracle):
    import ast
    from interpolatory_b200 import MEMCI, ArrayVideoStream
    from interpolatory_b200.ME.tss import get_motion_vectors as tss
    g = golden("tss_small.npz")
    for (H, W) in [(72, 96), (50, 70), (37, 53), (120, 40)]:
        fr = make_sequence(H, W, 2, seed=H)
        for bs, steps in ((8, 3), (8, 1), (4, 2), (16, 4), (8, 5)):
            assert np.array_equal(blocks_of(tss(bs, steps, fr[0], fr[1]), bs), g[f"{H}x{W}/tss/{bs}_{steps}"]), (H, W, bs, steps)
    fr = make_sequence(300, 500, 2, seed=8)
    assert np.array_equal(tss(8, 3, fr[0], fr[1]), oracle.tss(8, 3, fr[0], fr[1]))
    frames = make_sequence(72, 96, 3, seed=2072)
    for si in range(2):
        st = dict(ast.literal_eval(str(g[f"e2e/set{si}/settings"])))
        it = MEMCI(60, **st)
        it.set_video_stream(ArrayVideoStream(list(frames), 24))
        for i, want in enumerate(g[f"e2e/set{si}/out"]):
            assert np.array_equal(np.asarray(it.get_interpolated_frame(i)), want), (st, i)
        it.close()


# ---------------------------------------------------------------- bidir2 (MCI/Bidirectional_2.py)
def test_bidir2_vs_golden_and_oracle(gpu, golden, oracle):
    """IEWMC + EAC + BDHI + unpad on the GPU: function level against the reference's frames (fs and hbma
    vectors, three dists), whole clips through the class API (hbma / fs / tss), and against the oracle on
    larger fresh inputs (odd sizes: the hole-fill blocks then reach into the padding)."""
    import ast
    from interpolatory_b200 import MEMCI, ArrayVideoStream
    from interpolatory_b200.MCI.Bidirectional_2 import interpolate, get_weight_kern
    g = golden("bidir2_small.npz")
    assert np.array_equal(get_weight_kern(8), g["weight_kern"])
    for (H, W) in SHAPES:
        fr = make_sequence(H, W, 2, seed=H + 1000)
        k = f"{H}x{W}"
        for me in ("fs", "hbma"):
            fwd, bwd = g[f"{k}/{me}/fwd"], g[f"{k}/{me}/bwd"]
            up = me == "fs"
            if up:
                fwd = np.repeat(np.repeat(fwd, 8, axis=0), 8, axis=1)[:H, :W]
                bwd = np.repeat(np.repeat(bwd, 8, axis=0), 8, axis=1)[:H, :W]
            for di, dist in enumerate((0.5, 0.39999998, 0.8)):
                out = interpolate(fr[0], fr[1], fwd, bwd, dist, 16, 4, up)
                assert np.array_equal(out, g[f"{k}/{me}/out{di}"]), (k, me, dist)
    frames = make_sequence(72, 96, 3, seed=2272)
    for si in range(4):
        st = dict(ast.literal_eval(str(g[f"e2e/set{si}/settings"])))
        it = MEMCI(60, **st)
        it.set_video_stream(ArrayVideoStream(list(frames), 24))
        for i, want in enumerate(g[f"e2e/set{si}/out"]):
            assert np.array_equal(np.asarray(it.get_interpolated_frame(i)), want), (st, i)
        it.close()
    # fresh, larger inputs against the oracle (class API for the vectors, oracle for the MCI half)
    for (H, W, me) in ((270, 483, 'fs'), (360, 640, 'hbma'), (203, 301, 'tss')):
        fr = make_sequence(H, W, 2, seed=H)
        it = MEMCI(60, mci_mode='bidir2', me_mode=me)
        it.set_video_stream(ArrayVideoStream(list(fr), 30))
        out = np.asarray(it.get_interpolated_frame(1))
        fwd, bwd = it.fwr_MV_field_dense, it.bwr_MV_field_dense
        assert np.array_equal(out, oracle.bidir2_frame(fr[0], fr[1], fwd, bwd, 0.5, True)), (H, W, me)
        it.close()


def test_fs_pair_launch_equals_two_launches(gpu):
    """memci_me_fs_pair (both directions in one persistent launch, one redo pass) against two memci_me_fs calls,
    on sizes that exercise clipped tiles, the generic (non multiple of 8) kernel and 16-px blocks."""
    from interpolatory_b200.device import DeviceContext
    for (H, W, bs, R) in ((270, 483, 8, 16), (131, 97, 8, 7), (120, 200, 6, 5), (256, 320, 16, 20)):
        fr = make_sequence(H, W, 2, seed=H + W)
        ctx = DeviceContext(H, W)
        ctx.upload_frame(0, fr[0]); ctx.upload_frame(1, fr[1])
        ctx.me_fs(0, 1, bs, R, 2); ctx.me_fs(1, 0, bs, R, 3)
        ctx.me_fs_pair(0, 1, bs, R, 0, 1)
        for a, b in ((0, 2), (1, 3)):
            got, want = ctx.download_mv_blocks(a), ctx.download_mv_blocks(b)
            assert np.array_equal(got[1], want[1]) and np.array_equal(got[3], want[3]), (H, W, bs, R, a)
        ctx.close()


def test_pipeline_bidir2_equals_class_api(gpu):
    """PairPipeline (the e2e path of bench.py) with mci_mode=bidir2 against BiDir2Interpolator frame by frame
    (24 -> 120 fps: the same four dists inside every pair, which is what the pipeline API assumes)."""
    from interpolatory_b200 import MEMCI, ArrayVideoStream
    from interpolatory_b200.pipeline import PairPipeline, interpolation_dists
    H, W = 136, 200
    frames = make_sequence(H, W, 4, seed=77)
    st = dict(me_mode='hbma', mci_mode='bidir2')
    dists = interpolation_dists(120 / 24)
    assert len(dists) == 4
    pipe = PairPipeline(H, W, n_contexts=2, **st)
    out = np.empty((3 * len(dists), H, W, 3), np.uint8)
    pipe.run(frames, dists, out)
    pipe.close()
    it = MEMCI(120, **st)
    it.set_video_stream(ArrayVideoStream(list(frames), 24))
    k = 0
    for i in range(1, 15):
        first, ratio = it_first_ratio(i, 120 / 24)
        if (1. - ratio) == 0.:
            continue                                       # pass-through source frame
        got = np.asarray(it.get_interpolated_frame(i))
        assert np.array_equal(got, out[k]), (i, k)
        k += 1
    assert k == 12
    it.close()


def it_first_ratio(idx, rr):
    from interpolatory_b200.util import get_first_frame_idx_and_ratio
    return get_first_frame_idx_and_ratio(idx, rr)


def test_bidir2_out_of_range_blocks_raise(gpu):
    """A vector that shifts an expanded block outside the padded frame: the reference's slices would clip and its
    in-place add fail; we raise as well (after zero-filling on the device)."""
    from interpolatory_b200.MCI.Bidirectional_2 import interpolate
    H, W = 64, 96
    fr = make_sequence(H, W, 2, seed=3)
    mv = np.zeros((H, W, 3), np.float32)
    ok = interpolate(fr[0], fr[1], mv, mv, 0.5, 16, 4, True)
    assert ok.shape == (H, W, 3)
    mv[0:8, 0:8, 0] = -40.0                    # 14 + (-40) < 0: leaves the 16-px padding
    with pytest.raises(Exception, match='outside the padded frame'):
        interpolate(fr[0], fr[1], mv, mv, 0.5, 16, 4, True)


def test_cfg3_1080p_full_size_vs_oracle(gpu, oracle):
    """cfg3 at full size (1080p, HBMA with the class API's auto depth 4, median smoothing, the four 24->120 dists):
    the parent-level densify kernel, the splat's 4x4-cell lists, dense hole runs (dist 0.8) and bidir2 on the same
    vectors, against the oracle."""
    from interpolatory_b200.device import DeviceContext
    from interpolatory_b200._native import PairParams
    from interpolatory_b200.MCI.Bidirectional_2 import get_weight_kern
    from interpolatory_b200.util import hbma_auto_steps
    H, W = 1080, 1920
    fr = make_sequence(H, W, 2, seed=7)
    a, b = fr[0], fr[1]
    steps = hbma_auto_steps((H, W))
    assert steps == 4
    ctx = DeviceContext(H, W)
    try:
        ctx.upload_frame(0, a); ctx.upload_frame(1, b)
        # block-level vectors (what bidir2 consumes: upscale=False)
        ctx.estimate_pair(PairParams(1, 8, 7, 1, steps, 4, 0, 4, 1), 0, 1, 0, 1)
        fb = oracle.hbma(8, 7, 1, steps, 4, a, b, 'sad', False)
        bb = oracle.hbma(8, 7, 1, steps, 4, b, a, 'sad', False)
        for slot, ref in ((0, fb), (1, bb)):
            _, vec, _, sad = ctx.download_mv_blocks(slot)
            assert np.array_equal(vec, ref[:, :, :2]) and np.array_equal(sad, ref[:, :, 2]), slot
        ctx.mci_bidir2(0, 1, 0, 1, np.float32(0.4), np.float32(1 - 0.4), get_weight_kern(8), 7)
        assert np.array_equal(ctx.download_frame(7), oracle.bidir2_frame(a, b, fb, bb, 0.4, False))
        assert ctx.bidir2_stats()[1] == 0
        # median-smoothed dense fields (what unidir / bidir consume)
        ctx.estimate_pair(PairParams(1, 8, 7, 1, steps, 4, 2, 4, 1), 0, 1, 0, 1)
        fd = oracle.smooth('median', oracle.hbma(8, 7, 1, steps, 4, a, b), 4)
        bd = oracle.smooth('median', oracle.hbma(8, 7, 1, steps, 4, b, a), 4)
        for dist in (0.19999999, 0.8):
            ctx.mci_bidir(0, 1, 0, 1, np.float32(dist), 7)
            assert np.array_equal(ctx.download_frame(7), oracle.bidir_helper(a, b, fd, bd, np.float32(dist))), dist
        ctx.mci_unidir(0, 1, 0, np.float32(0.6), 7)
        assert np.array_equal(ctx.download_frame(7), oracle.unidir_helper(a, b, fd, np.float32(0.6)))
    finally:
        ctx.close()


def test_run_clip_24_to_60_equals_class_api(gpu):
    """PairPipeline.run_clip on a rate ratio whose dists differ from pair to pair (24 -> 60, BASELINE cfg1's ratio):
    every output frame -- pass-through and interpolated -- equals the class API's get_interpolated_frame(i)."""
    from interpolatory_b200 import MEMCI, ArrayVideoStream
    from interpolatory_b200.pipeline import PairPipeline
    H, W = 120, 168
    frames = make_sequence(H, W, 5, seed=321)
    st = dict(me_mode='fs', block_size='8', target_region='7', mci_mode='bidir')
    pipe = PairPipeline(H, W, n_contexts=3, **st)
    plan = pipe.plan_clip(len(frames), 24, 60)
    out = np.zeros((len(plan), H, W, 3), np.uint8)
    pipe.run_clip(frames, 24, 60, out)
    pipe.close()
    it = MEMCI(60, **st)
    it.set_video_stream(ArrayVideoStream(list(frames), 24))
    assert len(plan) >= 10 and sum(1 for _, d in plan if d == 0.0) >= 2
    for i in range(len(plan)):
        assert np.array_equal(np.asarray(it.get_interpolated_frame(i)), out[i]), (i, plan[i])
    it.close()


@pytest.mark.parametrize("st,fps_in,fps_out", [
    (dict(me_mode='fs', block_size='8', target_region='7', mci_mode='bidir'), 24, 60),
    (dict(me_mode='hbma', mci_mode='unidir', filter_mode='median'), 30, 60),
    (dict(me_mode='hbma', mci_mode='bidir2'), 24, 120),
    (dict(me_mode='tss', mci_mode='bidir', filter_mode='weighted'), 25, 60),
])
def test_interpolate_video_is_the_fast_path_and_equals_frame_by_frame(gpu, st, fps_in, fps_out):
    """BaseInterpolator.interpolate_video (the call the reference's CLI makes, base.py:78-115) runs the staged,
    batched pipeline in windows of max_cache_size / stream_window source frames: its frames must equal
    get_interpolated_frame(i) one by one, whatever the window size, on one device and on a two-way device list."""
    from interpolatory_b200 import MEMCI, ArrayVideoStream
    H, W, n_src = 120, 168, 8
    frames = list(make_sequence(H, W, n_src, seed=31))
    n_out = int(math.floor((n_src - 1) * fps_out / fps_in)) + 1          # up to and including the last source frame
    ref_it = MEMCI(fps_out, **st)
    ref_it.set_video_stream(ArrayVideoStream(frames, fps_in))
    want = [np.asarray(ref_it.get_interpolated_frame(i)).copy() for i in range(n_out)]
    ref_it.close()
    for extra in (dict(max_cache_size=2), dict(max_cache_size=3), dict(stream_window='5'), dict(stream_window='64'),
                  dict(stream_window='4', devices='0,0')):
        it = MEMCI(fps_out, max_out_frames=n_out, **extra, **st)
        it.set_video_stream(ArrayVideoStream(frames, fps_in))
        got = []
        it.interpolate_video(sink=got.append)
        launches = sum(p_.launch_count() for p_ in it._pipe)
        it.close()
        assert len(got) == n_out, extra
        for i, (g, w) in enumerate(zip(got, want)):
            assert np.array_equal(g, w), (st, extra, i)
        assert launches > 0 and it._dev is None            # the pipeline ran; the per-frame context was never created


def test_multigpu_pipeline_equals_single_pipeline(gpu):
    """One clip over N pipelines by pair index (multigpu.MultiGpuPipeline, one host thread each): byte-identical to the
    single-GPU run_clip.  Runs on two physical GPUs when the box has them, and always on a [0, 0, 0] device list (three
    pipelines sharing one GPU: the same partitioning, threading and staging code)."""
    from interpolatory_b200._native import lib
    from interpolatory_b200.multigpu import MultiGpuPipeline
    from interpolatory_b200.pipeline import PairPipeline
    H, W = 136, 200
    frames = make_sequence(H, W, 10, seed=77)
    st = dict(me_mode='fs', block_size='8', target_region='16', mci_mode='bidir')
    single = PairPipeline(H, W, n_contexts=2, **st)
    plan = single.plan_clip(len(frames), 24, 60)
    want = np.zeros((len(plan), H, W, 3), np.uint8)
    single.run_clip(frames, 24, 60, want)
    single.close()
    device_lists = [[0, 0, 0]]
    if lib().memci_device_count() >= 2:
        device_lists.append([0, 1])
    for devs in device_lists:
        mg = MultiGpuPipeline(H, W, devices=devs, n_contexts=2, **st)
        got = np.zeros_like(want)
        assert mg.run_clip(frames, 24, 60, got) == plan
        assert sum(mg.last_shards, []) == sorted({k for k, d in plan if d != 0.0})
        assert all(len(s_) > 0 for s_ in mg.last_shards)
        mg.close()
        assert np.array_equal(got, want), devs


@pytest.mark.parametrize("H,W", [(8, 8), (16, 24), (9, 40), (33, 17)])
def test_tiny_frames_vs_oracle(gpu, oracle, H, W):
    """Frames smaller than a tile / a search window / a hole-fill window: every kernel's edge clamps."""
    from interpolatory_b200.ME.full_search import get_motion_vectors as fs
    from interpolatory_b200.MCI.Bidirectional import helper as bi
    from interpolatory_b200.MCI.Unidirectional import helper as uni
    from interpolatory_b200.MCI.Bidirectional_2 import interpolate as b2
    rng = np.random.default_rng(H * 100 + W)
    a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    b = np.roll(a, (1, 2), axis=(0, 1))
    b[:2] = rng.integers(0, 256, (2, W, 3), dtype=np.uint8)
    for bs, R in ((8, 3), (4, 2)):
        f, g = fs(bs, R, a, b), fs(bs, R, b, a)
        assert np.array_equal(f, oracle.fs(bs, R, a, b)) and np.array_equal(g, oracle.fs(bs, R, b, a)), (bs, R)
        for dist in (0.5, 0.8):
            assert np.array_equal(bi(a, b, f, g, np.float32(dist)), oracle.bidir_helper(a, b, f, g, np.float32(dist)))
            assert np.array_equal(uni(a, b, f, np.float32(dist)), oracle.unidir_helper(a, b, f, np.float32(dist)))
        assert np.array_equal(b2(a, b, f, g, 0.5, 16, 4, True), oracle.bidir2_frame(a, b, f, g, 0.5, True))


def test_splat_fuzz_vs_oracle(gpu, oracle, monkeypatch):
    """Random block-constant MV fields through both MCI helpers against the oracle: cell sizes 1/2/4/6/8, vectors up to
    +-60 px (wrapped and out-of-frame targets, neighbourhoods too large to stage, key lists that overflow into the dense
    fallback, frames no larger than twice the largest displacement), fractional vectors, SAD ties and infinite SADs,
    several dists; each case through the kernel's automatic path choice and with the generic / dense paths forced."""
    from interpolatory_b200.MCI.Bidirectional import helper as bi
    from interpolatory_b200.MCI.Unidirectional import helper as uni
    rng = np.random.default_rng(2024)
    cases = [(64, 96, 8, 2), (64, 96, 8, 20), (90, 130, 4, 8), (72, 72, 6, 12), (48, 160, 1, 3), (120, 96, 8, 60), (56, 88, 4, 40),
             (40, 40, 4, 30), (33, 70, 1, 16), (64, 64, 8, 31), (64, 64, 8, 32), (70, 50, 2, 24)]
    for (H, W, g, D) in cases:
        a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        b = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        nr, nc = -(-H // g), -(-W // g)
        fields = []
        for _ in range(2):
            v = rng.integers(-D, D + 1, (nr, nc, 2)).astype(np.float32)
            if g == 4:
                v += rng.choice([0.0, 0.25, 0.5], (nr, nc, 2)).astype(np.float32)      # fractional (mean-smoothed) vectors
            sad = rng.integers(0, 6, (nr, nc, 1)).astype(np.float32) * 100.0              # many ties
            sad[rng.random((nr, nc, 1)) < 0.03] = np.inf
            f = np.concatenate([v, sad], axis=2)
            fields.append(np.ascontiguousarray(np.kron(f, np.ones((g, g, 1), np.float32))[:H, :W]))
        f_, b_ = fields
        for dist in (0.5, 0.19999999, 0.8):
            d = np.float32(dist)
            want_u, want_b = oracle.unidir_helper(a, b, f_, d), oracle.bidir_helper(a, b, f_, b_, d)
            for path in ("0", "1", "2"):            # MEMCI_MCI_PATH: 0 automatic, 1 no staged fast path, 2 dense enumeration
                monkeypatch.setenv("MEMCI_MCI_PATH", path)
                assert np.array_equal(uni(a, b, f_, d), want_u), (H, W, g, D, dist, 'unidir', path)
                assert np.array_equal(bi(a, b, f_, b_, d), want_b), (H, W, g, D, dist, 'bidir', path)


def _fuzz_field(rng, H, W, g, D, frac, sg=None, holes=0.03):
    """Dense [H, W, 3] field, (dy, dx) constant on g x g cells and SAD constant on sg x sg cells."""
    nr, nc = -(-H // g), -(-W // g)
    v = rng.integers(-D, D + 1, (nr, nc, 2)).astype(np.float32)
    if rng.integers(0, 3) == 0:                                   # few distinct vectors, like a real field
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


def test_mci_fuzz_random_geometry(gpu, oracle, monkeypatch):
    """Seeded random geometry for all three MCI modes: frame sizes 8..150 x 8..200, independent MV cell sizes per
    direction (1, 2, 4, 5, 6, 8, 16), SAD cells that nest in the MV cells or not, vectors up to +-100 px, random dist,
    a random forced splat path.  Where the reference itself is undefined (a negative index below -H: numba writes out
    of bounds) the oracle reports it and the case is skipped; bidir2 cases that leave the 16-px pad must raise on both
    sides."""
    from interpolatory_b200.MCI.Bidirectional import helper as bi
    from interpolatory_b200.MCI.Unidirectional import helper as uni
    from interpolatory_b200.MCI.Bidirectional_2 import interpolate as b2
    rng = np.random.default_rng(11)
    n_mci = n_b2 = n_b2_raise = 0
    for _ in range(120):
        H, W = int(rng.integers(8, 150)), int(rng.integers(8, 200))
        a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        b = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        gf, gb = int(rng.choice([1, 2, 4, 8, 16, 5])), int(rng.choice([1, 2, 4, 8, 16, 6]))
        D = int(rng.choice([0, 1, 3, 8, 20, 45, 100]))
        sgf = gf * int(rng.choice([1, 1, 2])) if rng.integers(0, 2) else int(rng.choice([3, 4, 8]))
        f_ = _fuzz_field(rng, H, W, gf, D, rng.integers(0, 2), sgf)
        b_ = _fuzz_field(rng, H, W, gb, D, rng.integers(0, 2))
        dist = np.float32(rng.choice([0.5, 0.19999999, 0.8, 0.4, 0.6, 1 / 3, 0.99]))
        path = str(rng.integers(0, 3))
        monkeypatch.setenv("MEMCI_MCI_PATH", path)
        cfg = (H, W, gf, gb, D, sgf, float(dist), path)
        try:
            want_u, want_b = oracle.unidir_helper(a, b, f_, dist), oracle.bidir_helper(a, b, f_, b_, dist)
        except RuntimeError:
            want_u = None                                          # undefined in the reference
        if want_u is not None:
            assert np.array_equal(uni(a, b, f_, dist), want_u), ('unidir',) + cfg
            assert np.array_equal(bi(a, b, f_, b_, dist), want_b), ('bidir',) + cfg
            n_mci += 1
        up = bool(rng.integers(0, 2)); D2 = int(rng.choice([0, 2, 6, 14])); dd = float(rng.choice([0.5, 0.4, 0.8, 0.25]))
        if up:
            g2 = int(rng.choice([4, 8, 16]))
            f2, b2_ = _fuzz_field(rng, H, W, g2, D2, rng.integers(0, 2), holes=0), _fuzz_field(rng, H, W, g2, D2, rng.integers(0, 2), holes=0)
        else:
            nr, nc = -(-H // 4), -(-W // 4)
            f2, b2_ = _fuzz_field(rng, nr, nc, 1, D2, rng.integers(0, 2), holes=0), _fuzz_field(rng, nr, nc, 1, D2, rng.integers(0, 2), holes=0)
        try:
            want2 = oracle.bidir2_frame(a, b, f2, b2_, dd, up)
        except RuntimeError:
            want2 = None
        if want2 is None:
            with pytest.raises(Exception, match="shifted outside the padded frame"):
                b2(a, b, f2, b2_, dd, 16, 4, up)
            n_b2_raise += 1
        else:
            assert np.array_equal(b2(a, b, f2, b2_, dd, 16, 4, up), want2), ('bidir2', H, W, up, D2, dd)
            n_b2 += 1
    assert n_mci >= 80 and n_b2 >= 60 and n_b2_raise >= 5


def test_me_fuzz_vs_oracle(gpu, oracle):
    """Seeded random sizes and parameters for full search, three-step search, HBMA (sad / ssd, upscale or not) and the
    three smoothing filters, on five kinds of frames: noise, two grey levels (every SAD a tie and an exact integer),
    flat, shifted noise plus a grey offset, and the synthetic sequence."""
    from interpolatory_b200.ME.full_search import get_motion_vectors as fs
    from interpolatory_b200.ME.tss import get_motion_vectors as tss
    from interpolatory_b200.ME.hbma import get_motion_vectors as hbma
    from interpolatory_b200.smoothing import smooth
    rng = np.random.default_rng(7)

    def frames(kind, H, W):
        if kind == 0:
            return rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
        if kind == 1:
            return np.repeat(rng.integers(0, 2, (2, H, W, 1), dtype=np.uint8) * 50 + 20, 3, axis=3)
        if kind == 2:
            a = np.empty((2, H, W, 3), np.uint8); a[0] = 77; a[1] = 80
            return a
        if kind == 3:
            a = np.repeat(rng.integers(0, 200, (H, W, 1), dtype=np.uint8), 3, axis=2)
            b = (np.roll(a, (int(rng.integers(-5, 6)), int(rng.integers(-5, 6))), (0, 1)).astype(np.int32) + int(rng.integers(0, 5))).astype(np.uint8)
            return np.stack([a, b])
        return make_sequence(H, W, 2, seed=int(rng.integers(1 << 30)))

    for _ in range(600):
        H, W = int(rng.integers(8, 200)), int(rng.integers(8, 260))
        kind = int(rng.integers(0, 5))
        if kind == 4 and (H < 48 or W < 48):
            kind = 0
        a, b = frames(kind, H, W)
        bs = int(rng.choice([2, 3, 4, 5, 6, 8, 8, 8, 12, 16, 16, 24, 32]))
        R = int(rng.integers(0, 24))
        assert np.array_equal(fs(bs, R, a, b), oracle.fs(bs, R, a, b)), ('fs', H, W, kind, bs, R)
        steps = int(rng.integers(1, 6))
        assert np.array_equal(tss(bs, steps, a, b), oracle.tss(bs, steps, a, b)), ('tss', H, W, kind, bs, steps)
        hb, win, sub, st = int(rng.choice([4, 8, 8, 16, 32])), int(rng.integers(1, 10)), int(rng.integers(1, 4)), int(rng.integers(0, 4))
        mbs, cost, up = int(rng.choice([2, 4, 4, 8])), str(rng.choice(['sad', 'ssd'])), bool(rng.integers(0, 2))
        assert np.array_equal(hbma(hb, win, sub, st, mbs, a, b, cost, up), oracle.hbma(hb, win, sub, st, mbs, a, b, cost, up)), \
            ('hbma', H, W, kind, hb, win, sub, st, mbs, cost, up)
        g = int(rng.choice([1, 2, 4, 8, 16]))
        mv = _fuzz_field(rng, H, W, g, 20, rng.integers(0, 2))
        fsz = int(rng.choice([1, 2, 4, 8, 16, 3]))
        for mode in ("mean", "median", "weighted"):
            assert np.array_equal(np.asarray(smooth(mode, mv, fsz)), oracle.smooth(mode, mv, fsz)), ('smooth', H, W, g, fsz, mode)


# ---------------------------------------------------------------- Middlebury harness: PSNR / SSIM on the device
SSIM_RTOL = 1e-9      # float64 on both sides; the box sums are exact integers here, running f64 sums in scipy


def test_quality_metrics_vs_oracle(gpu):
    """memci_quality through metrics.py against the restated skimage algorithm (oracle/quality.py): MSE exactly,
    SSIM within SSIM_RTOL, on noise, near-identical, flat and synthetic frames, sizes off the 32 x 8 tile grid."""
    import quality as Q
    from interpolatory_b200 import metrics as M
    rng = np.random.default_rng(17)
    for (H, W) in [(7, 7), (8, 40), (37, 53), (100, 131), (360, 640), (480, 584)]:
        a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        noisy = np.clip(a.astype(np.int32) + rng.integers(-12, 13, a.shape), 0, 255).astype(np.uint8)
        flat = np.full((H, W, 3), 93, np.uint8)
        cases = [(a, noisy), (a, a), (a, rng.integers(0, 256, a.shape, dtype=np.uint8)), (flat, a), (flat, flat + 4)]
        if H >= 48 and W >= 48:
            fr = make_sequence(H, W, 2, seed=H)
            cases.append((fr[0], fr[1]))
        for x, y in cases:
            mse, ssim = M.frame_quality(x, y)
            assert mse == Q.mean_squared_error(x, y)
            want = Q.structural_similarity(x, y)
            assert abs(ssim - want) <= SSIM_RTOL * abs(want), (H, W, ssim, want)
            assert M.structural_similarity(x, y, data_range=255, multichannel=True) == ssim
            p, wp = M.peak_signal_noise_ratio(x, y, data_range=255), Q.peak_signal_noise_ratio(x, y)
            assert p == wp or abs(p - wp) <= 1e-12 * abs(wp)
    with pytest.raises(ValueError):
        M.structural_similarity(np.zeros((6, 20, 3), np.uint8), np.zeros((6, 20, 3), np.uint8), data_range=255, multichannel=True)
    with pytest.raises(NotImplementedError):
        M.structural_similarity(a, a, data_range=255, multichannel=True, gaussian_weights=True)


def test_middlebury_harness(gpu, tmp_path, capsys):
    """benchmark() and get_middle_frame() (src/Benchmark.py:18-114) over a synthetic <name>/frame10.png,
    frame10i11.png, frame11.png dataset: written PNGs equal get_benchmark_frame, results.txt / printed JSON carry
    the oracle's mean PSNR and SSIM of those frames."""
    import json
    import quality as Q
    from PIL import Image
    from interpolatory_b200 import MEMCI
    from interpolatory_b200.Benchmark import benchmark, get_middle_frame
    data, outd = tmp_path / "middlebury", tmp_path / "out"
    outd.mkdir()
    truths = {}
    for i, (name, (H, W)) in enumerate({"Alpha": (96, 128), "Beta": (120, 90), "Gamma": (75, 141)}.items()):
        fr = make_sequence(H, W, 3, seed=400 + i)
        (data / name).mkdir(parents=True)
        for fn, im in (("frame10", fr[0]), ("frame10i11", fr[1]), ("frame11", fr[2])):
            Image.fromarray(im).save(data / name / f"{fn}.png")
        truths[name] = fr
    for settings in (dict(me_mode='fs', mci_mode='bidir'), dict(me_mode='hbma', mci_mode='bidir2')):
        it = MEMCI(2, **settings)
        res = benchmark(it, str(outd), dataset_dir=str(data))
        ps, ss = [], []
        for name, fr in truths.items():
            want = np.asarray(MEMCI(2, **settings).get_benchmark_frame(fr[0], fr[2]))
            got = np.asarray(Image.open(outd / f"{name}.png"))
            assert np.array_equal(got, want), (settings, name)
            ps.append(Q.peak_signal_noise_ratio(fr[1], want)); ss.append(Q.structural_similarity(fr[1], want))
        assert abs(res['PSNR'] - np.mean(ps)) < 1e-9 and abs(res['SSIM'] - np.mean(ss)) < 1e-9
        on_disk = json.loads((outd / "results.txt").read_text())
        assert on_disk['PSNR'] == res['PSNR'] and on_disk['SSIM'] == res['SSIM'] and 'Time taken' in on_disk
        assert json.loads(capsys.readouterr().out.strip().splitlines()[-1])['SSIM'] == res['SSIM']
        fr = truths["Alpha"]
        r2 = get_middle_frame(it, str(data / "Alpha" / "frame10.png"), str(data / "Alpha" / "frame11.png"),
                              str(outd / "mid.png"), str(data / "Alpha" / "frame10i11.png"))
        assert abs(r2['SSIM'] - ss[0]) < 1e-9 and abs(r2['PSNR'] - ps[0]) < 1e-9
        # ground truth = the frame just written: PSNR is reported as the string the reference prints (:111-112)
        assert get_middle_frame(it, str(data / "Alpha" / "frame10.png"), str(data / "Alpha" / "frame11.png"),
                                str(outd / "same.png"), str(outd / "same.png")) == {'PSNR': 'Infinity', 'SSIM': 1.0}
        it.close()
    with pytest.raises(FileNotFoundError):
        benchmark(MEMCI(2), None, dataset_dir=str(tmp_path / "nothing"))


# ---------------------------------------------------------------- full search: two-phase path (8-bit prefilter + exact refinement)
@pytest.mark.parametrize("H,W,bs,R", [(120, 200, 8, 16), (97, 131, 8, 7), (270, 483, 8, 16), (136, 264, 16, 12), (360, 640, 16, 32)])
def test_fs_prefilter_path_equals_exact_kernel_and_oracle(gpu, oracle, monkeypatch, H, W, bs, R):
    """The default full search prunes with an 8-bit prefilter and evaluates the survivors exactly (csrc/fs8.cu); with
    MEMCI_FS_PREFILTER=0 the exact tiled kernel searches every candidate.  Both must give the oracle's field, the
    prefilter must actually prune, and a survivor list that overflows (forced here with a tiny segment capacity; flat
    frames do it on their own) must fall back to the exact kernel on the device."""
    from interpolatory_b200.device import DeviceContext
    fr = make_sequence(H, W, 2, seed=11 * H + W)
    ctx = DeviceContext(H, W)
    ctx.upload_frame(0, fr[0]); ctx.upload_frame(1, fr[1])
    want = [oracle.fs_blocks(bs, R, fr[0], fr[1]), oracle.fs_blocks(bs, R, fr[1], fr[0])]

    def run():
        ctx.me_fs_pair(0, 1, bs, R, 0, 1)
        out = [ctx.download_mv_blocks(0), ctx.download_mv_blocks(1)]
        return [np.concatenate([o[1], o[3][:, :, None]], axis=2) for o in out]

    monkeypatch.setenv("MEMCI_FS_PREFILTER", "1")
    got = run()
    n_surv, overflow = ctx.fs_prefilter_stats()
    n_cand = ctx.fs_stats()[0]
    assert overflow == 0 and 0 < n_surv < 0.2 * n_cand, (n_surv, n_cand)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    # no room on the survivor list: blocks go to the block-level redo pass, and past its limit the pair falls back
    for cap in ("40", "0"):
        monkeypatch.setenv("MEMCI_F8_SEG_CAP", cap)
        got = run()
        if cap == "0":
            assert ctx.fs_prefilter_stats()[1] == 1
        for g, w in zip(got, want):
            assert np.array_equal(g, w), cap
    monkeypatch.delenv("MEMCI_F8_SEG_CAP")
    monkeypatch.setenv("MEMCI_FS_PREFILTER", "0")
    got = run()
    assert ctx.fs_prefilter_stats() == (0, 0)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    # one direction, and a block-row range (the band-split entry point), through the prefilter
    monkeypatch.setenv("MEMCI_FS_PREFILTER", "1")
    ctx.me_fs(1, 0, bs, R, 2)
    o = ctx.download_mv_blocks(2)
    assert np.array_equal(np.concatenate([o[1], o[3][:, :, None]], axis=2), want[1])
    ctx.close()


def test_fs_prefilter_segment_partial_fill(gpu, oracle, monkeypatch):
    """Sub-lists of the survivor list that fill up PART of the way through a search (round-1 advisor finding: the old
    cursor rolled reservations back, which could drop entries and hand unwritten ones to the refine pass; the scan pass
    now only ever moves a cursor forward and raises the pair-level fallback flag when a reservation does not fit).  The
    capacity is set just below, at a third of, and far below what a sub-list collects; the warps of many CTAs reserve
    concurrently.  Whatever happens, the exact kernel must then redo the pair: fields equal to the oracle's, every time."""
    from interpolatory_b200.device import DeviceContext
    H, W, bs, R = 270, 483, 8, 16
    fr = make_sequence(H, W, 2, seed=5)
    want = [oracle.fs_blocks(bs, R, fr[0], fr[1]), oracle.fs_blocks(bs, R, fr[1], fr[0])]
    ctx = DeviceContext(H, W)
    ctx.upload_frame(0, fr[0]); ctx.upload_frame(1, fr[1])
    monkeypatch.setenv("MEMCI_FS_PREFILTER", "1")
    ctx.me_fs_pair(0, 1, bs, R, 0, 1)
    n_surv, overflow = ctx.fs_prefilter_stats()
    assert n_surv > 0 and not overflow
    per_list = max(n_surv // 64, 8)                       # 64 sub-lists (F8_NLIST), filled round-robin by the scan CTAs
    for cap in (per_list - 1, per_list // 3, 7, 1):
        monkeypatch.setenv("MEMCI_F8_SEG_CAP", str(cap))
        for _ in range(3):                                  # the interleaving of the emitting warps differs from run to run
            ctx.me_fs_pair(0, 1, bs, R, 0, 1)
            for slot, ref in ((0, want[0]), (1, want[1])):
                _, vec, _, sad = ctx.download_mv_blocks(slot)
                assert np.array_equal(vec, ref[:, :, :2]) and np.array_equal(sad, ref[:, :, 2]), (cap, slot)
    monkeypatch.delenv("MEMCI_F8_SEG_CAP")
    ctx.close()


def test_fs_prefilter_shapes_and_chunks(gpu, oracle, monkeypatch):
    """The two-phase path under every launch shape and list geometry it can take: 256-thread jobs, no two-row jobs,
    one / two stages, the code array cut into chunks of a few jobs (8K frames do that for real), every per-column
    survivor cap from 1 (nearly everything goes to the redo pass) to 127 (nothing does), and search ranges whose
    row offsets need the second half of the hit mask (2R + 1 > 64) and several batches of code words per column."""
    from interpolatory_b200.device import DeviceContext
    monkeypatch.setenv("MEMCI_FS_PREFILTER", "1")
    cases = (((270, 483, 8, 16), ({}, {"MEMCI_F8_NT": "256"}, {"MEMCI_F8_NOSPLIT": "1"}, {"MEMCI_F8_STAGES": "1"}, {"MEMCI_F8_STAGES": "2"},
                                  {"MEMCI_F8_CHUNK_JOBS": "5"}, {"MEMCI_F8_CHUNK_JOBS": "1", "MEMCI_F8_NT": "256"}, {"MEMCI_F8_CTAS": "1"},
                                  {"MEMCI_F8_COL_HITS": "1"}, {"MEMCI_F8_COL_HITS": "3"}, {"MEMCI_F8_COL_HITS": "127"}, {"MEMCI_F8_REDO_DIV": "1000"})),
             ((200, 328, 8, 40), ({}, {"MEMCI_F8_CHUNK_JOBS": "3"}, {"MEMCI_F8_COL_HITS": "127"})),
             ((131, 250, 16, 32), ({}, {"MEMCI_F8_COL_HITS": "127"}, {"MEMCI_F8_CHUNK_JOBS": "2"})),
             ((96, 160, 16, 5), ({}, {"MEMCI_F8_NT": "256"})))
    knobs = ("MEMCI_F8_NT", "MEMCI_F8_NOSPLIT", "MEMCI_F8_STAGES", "MEMCI_F8_CHUNK_JOBS", "MEMCI_F8_CTAS", "MEMCI_F8_COL_HITS", "MEMCI_F8_REDO_DIV")
    for (H, W, bs, R), envs in cases:
        fr = make_sequence(H, W, 2, seed=H + R)
        want = [oracle.fs_blocks(bs, R, fr[0], fr[1]), oracle.fs_blocks(bs, R, fr[1], fr[0])]
        for env in envs:
            for k in knobs:
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            ctx = DeviceContext(H, W)
            ctx.upload_frame(0, fr[0]); ctx.upload_frame(1, fr[1])
            for rep in range(2):                            # the second search runs on re-armed buffers
                ctx.me_fs_pair(0, 1, bs, R, 0, 1)
                for slot, ref in ((0, want[0]), (1, want[1])):
                    _, vec, _, sad = ctx.download_mv_blocks(slot)
                    assert np.array_equal(vec, ref[:, :, :2]) and np.array_equal(sad, ref[:, :, 2]), (H, W, bs, R, env, rep, slot)
            assert ctx.fs_prefilter_stats() != (0, 0), (H, W, bs, R, env)       # the two-phase path did run
            ctx.close()


def test_fs_prefilter_not_applicable_beyond_its_entry_format(gpu, monkeypatch):
    """A survivor entry carries a 17-bit block index: a frame with 2^17 blocks or more (8K at 8 x 8) must take the exact
    kernel even with the prefilter pinned on -- same field as with it pinned off, and the statistics say it did not run."""
    from interpolatory_b200.device import DeviceContext
    H = W = 2904                                            # 363 x 363 = 131 769 blocks of 8 x 8
    fr = make_sequence(H, W, 2, seed=17)
    got = {}
    for pf in ("1", "0"):
        monkeypatch.setenv("MEMCI_FS_PREFILTER", pf)
        ctx = DeviceContext(H, W)
        ctx.upload_frame(0, fr[0]); ctx.upload_frame(1, fr[1])
        ctx.me_fs_pair(0, 1, 8, 4, 0, 1)
        got[pf] = [ctx.download_mv_blocks(s) for s in (0, 1)]
        assert ctx.fs_prefilter_stats() == (0, 0)
        ctx.close()
    for a, b in zip(got["1"], got["0"]):
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])
    # one block row fewer is inside the format: the two-phase path runs and agrees with the exact kernel
    H2 = 2896 - 8 * 2                                       # 360 x 363 = 130 680 blocks
    fr2 = fr[:, :H2]
    for pf in ("1", "0"):
        monkeypatch.setenv("MEMCI_FS_PREFILTER", pf)
        ctx = DeviceContext(H2, W)
        ctx.upload_frame(0, np.ascontiguousarray(fr2[0])); ctx.upload_frame(1, np.ascontiguousarray(fr2[1]))
        ctx.me_fs_pair(0, 1, 8, 4, 0, 1)
        got[pf] = [ctx.download_mv_blocks(s) for s in (0, 1)]
        assert (ctx.fs_prefilter_stats() != (0, 0)) == (pf == "1")
        ctx.close()
    for a, b in zip(got["1"], got["0"]):
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])


def test_fs_prefilter_flat_and_tie_frames(gpu, oracle):
    """Content where the prefilter cannot prune: flat frames (every candidate survives, the list overflows), two grey
    levels (every SAD an exact multiple of 1000: all survivors are flagged and go to the float64 redo pass)."""
    from interpolatory_b200.ME.full_search import get_motion_vectors as fs
    rng = np.random.default_rng(3)
    H, W = 96, 160
    flat = np.empty((2, H, W, 3), np.uint8); flat[0] = 77; flat[1] = 80
    two = np.repeat(rng.integers(0, 2, (2, H, W, 1), dtype=np.uint8) * 50 + 20, 3, axis=3)
    grad = np.repeat((np.arange(W, dtype=np.uint8)[None, :, None] + np.zeros((H, 1, 1), np.uint8)), 3, axis=2)
    white, black = np.full((H, W, 3), 255, np.uint8), np.zeros((H, W, 3), np.uint8)      # 16 x 16: A = 65280, threshold past the uint16 sentinel
    for a, b in ((flat[0], flat[1]), (two[0], two[1]), (grad, np.roll(grad, 3, axis=1)), (flat[0], two[1]), (white, black), (white, two[1])):
        for bs, R in ((8, 16), (16, 9), (8, 3)):
            assert np.array_equal(fs(bs, R, a, b), oracle.fs(bs, R, a, b)), (bs, R)


def test_fs_prefilter_backs_off_on_flat_content(gpu, oracle, monkeypatch):
    """Content feedback: a search whose prefilter could not prune (flat frames: every block goes to the redo list, past
    its limit the exact kernel redoes the pair) makes the next searches skip the prefilter; textured content keeps it on."""
    from interpolatory_b200.device import DeviceContext
    monkeypatch.delenv("MEMCI_FS_PREFILTER", raising=False)
    H, W = 270, 480            # (on frames of a few dozen blocks the border blocks dominate and the cost estimate rightly prefers the exact kernel)
    flat = np.empty((2, H, W, 3), np.uint8); flat[0] = 77; flat[1] = 80
    tex = make_sequence(H, W, 2, seed=9)
    for frames, expect_on in ((tex, True), (flat, False)):
        ctx = DeviceContext(H, W)
        ctx.upload_frame(0, frames[0]); ctx.upload_frame(1, frames[1])
        want = oracle.fs_blocks(8, 16, frames[0], frames[1])
        used = []
        for _ in range(4):
            ctx.me_fs(0, 1, 8, 16, 0)
            _, vec, _, sad = ctx.download_mv_blocks(0)          # synchronises: the feedback copy has landed
            assert np.array_equal(vec, want[:, :, :2]) and np.array_equal(sad, want[:, :, 2])
            used.append(ctx.fs_prefilter_stats() != (0, 0))
        assert used[0] and used[1:] == [expect_on] * 3, used
        ctx.close()


def test_mci_vs_reference_hashes_on_fuzz_geometry(gpu, golden, monkeypatch):
    """The same 82 adversarial inputs through the CUDA helpers, against the hashes of the reference's own frames
    (no oracle in between), with each splat path forced in turn."""
    import fuzz_cases as F
    from interpolatory_b200.MCI.Bidirectional import helper as bi
    from interpolatory_b200.MCI.Unidirectional import helper as uni
    g = golden("fuzz_reference_sha.npz")
    for it, a, b, f_, b_, dist in F.mci_cases(int(g["seed"]), int(g["rounds"])):
        if f"{it}/undefined" in g:
            continue
        monkeypatch.setenv("MEMCI_MCI_PATH", str(it % 3))
        assert F.sha(uni(a, b, f_, dist)) == str(g[f"{it}/unidir"]), it
        assert F.sha(bi(a, b, f_, b_, dist)) == str(g[f"{it}/bidir"]), it


def test_bidir2_vs_reference_hashes_on_fuzz_fields(gpu, golden):
    """The same random bidir2 inputs through the CUDA path against the hashes of the reference's own frames; where the
    reference raises (a block shifted out of the pad) the GPU path raises too."""
    import fuzz_cases as F
    from interpolatory_b200.MCI.Bidirectional_2 import interpolate as b2
    g = golden("fuzz_reference_sha.npz")
    for it, a, b, f2, bw2, dd, up in F.bidir2_cases(int(g["b2_seed"]), int(g["b2_rounds"])):
        if f"b2/{it}/out_of_pad" in g:
            with pytest.raises(Exception, match="shifted outside the padded frame"):
                b2(a, b, f2, bw2, dd, 16, 4, up)
            continue
        assert F.sha(b2(a, b, f2, bw2, dd, 16, 4, up)) == str(g[f"b2/{it}/out"]), it


def test_me_vs_reference_hashes_on_random_parameters(gpu, golden):
    """Full search, TSS, HBMA and the smoothing filters with random parameters against the hashes of the reference's own
    fields (oracle/validate_fuzz_vs_reference.py --write-golden); HBMA cases where the reference reads out of bounds are
    left out."""
    import fuzz_cases as F
    from interpolatory_b200.ME.full_search import get_motion_vectors as fs
    from interpolatory_b200.ME.tss import get_motion_vectors as tss
    from interpolatory_b200.ME.hbma import get_motion_vectors as hbma
    from interpolatory_b200.smoothing import smooth
    g = golden("fuzz_reference_sha.npz")
    for it, fr, q, mv in F.me_cases(int(g["me_seed"]), int(g["me_rounds"])):
        a, b = fr[0], fr[1]
        calls = {"fs": lambda: fs(q['bs'], q['R'], a, b),
                 "tss": lambda: tss(q['bs'], q['steps'], a, b),
                 "hbma": lambda: hbma(q['hb'], q['win'], q['sub'], q['hsteps'], q['mbs'], a, b, q['cost'], q['up'])}
        for mode in ("mean", "median", "weighted"):
            calls[f"smooth_{mode}"] = lambda mode=mode: smooth(mode, mv, q['fsz'])
        for name, f in calls.items():
            key = f"me/{it}/{name}"
            if key + "/undefined" in g or key + "/raises" in g:
                continue
            assert F.sha(np.asarray(f(), np.float32)) == str(g[key]), (it, name, q)
