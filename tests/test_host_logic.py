"""CPU suite, part 2: host-side logic of the product (no GPU, no compute calls into the library)."""
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol(native_lib):
    """include/memci_b200.h is the ABI: every MEMCI_API function must be exported by the .so
    and bound by the ctypes layer."""
    from interpolatory_b200 import _native
    hdr = open(os.path.join(ROOT, "include", "memci_b200.h")).read()
    declared = set(re.findall(r"MEMCI_API\s+[\w\s\*]+?\b(memci_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(native_lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    assert native_lib.memci_version() >= 100


def test_no_cpu_fallback(native_lib):
    """Without a device the product path must fail loudly, never compute on the host."""
    from interpolatory_b200 import _native
    from interpolatory_b200.ME.full_search import get_motion_vectors as fs
    if native_lib.memci_device_count() > 0:
        pytest.skip("a GPU is visible")
    z = np.zeros((16, 16, 3), np.uint8)
    with pytest.raises(_native.MemciError):
        fs(8, 7, z, z)


def test_product_never_imports_oracle():
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "interpolatory_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M) or "memci_oracle" in txt or "oracle/" in txt:
                    bad.append(f)
    assert not bad, bad


def test_settings_parsing_and_dispatch():
    from interpolatory_b200.MCI.MEMCIBaseInterpolator import MEMCIBaseInterpolator, fs, hbma
    from interpolatory_b200.smoothing import median_filter
    it = MEMCIBaseInterpolator(60)
    assert (it.block_size, it.region, it.filter_size, it.sub_region, it.min_block_size, it.steps) == (8, 7, 4, 1, 4, 3)
    assert it.me_mode is hbma and it.filter_mode is None and it.MV_field_idx == -1 and it.pad_size == 16
    assert it.ME_args["upscale"] is False and it.ME_args["cost_str"] == "sad"
    it = MEMCIBaseInterpolator(60, me_mode='fs', block_size='16', target_region='32', filter_mode='median', filter_size='8')
    assert it.me_mode is fs and it.ME_args == {"block_size": 16, "target_region": 32}
    assert it.filter_mode is median_filter and it.filter_size == 8
    with pytest.raises(KeyError):
        MEMCIBaseInterpolator(60, me_mode='nope')
    with pytest.raises(KeyError):
        MEMCIBaseInterpolator(60, filter_mode='nope')
    with pytest.raises(ValueError):
        MEMCIBaseInterpolator(60, block_size='8.5')


def test_factory():
    from interpolatory_b200 import MEMCI, UniDirInterpolator, BiDirInterpolator, InterpolatorDictionary
    assert isinstance(MEMCI(60), UniDirInterpolator) and str(MEMCI(60)) == 'uniMEMCI'
    assert isinstance(MEMCI(60, mci_mode='bidir'), BiDirInterpolator) and str(MEMCI(60, mci_mode='bidir')) == 'bwMEMCI'
    assert InterpolatorDictionary['MEMCI'] is MEMCI
    with pytest.raises(Exception, match='Unknown mci_mode'):
        MEMCI(60, mci_mode='sideways')
    from interpolatory_b200.MCI.Bidirectional_2 import BiDir2Interpolator
    b2 = MEMCI(60, mci_mode='bidir2', filter_mode='median')
    assert isinstance(b2, BiDir2Interpolator) and str(b2) == 'Uni_2' and b2.filter_mode is None      # bidir2 forces no smoothing
    assert b2.upscale_MV is False and b2.ME_args["upscale"] is False and b2.pad_size == 16            # hbma: block-level vectors
    it = MEMCI(60, me_mode='tss', target_region='9')
    assert it.region == it.steps == 3 and it.ME_args == {"block_size": 8, "steps": 3}      # tss forces region = steps


def test_cli_settings_string():
    from interpolatory_b200.util import deconstruct_interpolation_mode_and_settings as dec
    mode, st = dec("MEMCI:me_mode=fs.mci_mode=bidir.block_size=8.target_region=16")
    assert mode == "MEMCI" and st == dict(me_mode='fs', mci_mode='bidir', block_size='8', target_region='16')
    assert dec("MEMCI") == ("MEMCI", {})
    assert dec("MEMCI:junk.block_size=4")[1] == {"block_size": "4"}


def test_auto_pyramid_depth_and_dist_cycle():
    from interpolatory_b200.util import hbma_auto_steps, get_first_frame_idx_and_ratio
    from interpolatory_b200.pipeline import interpolation_dists
    assert hbma_auto_steps((1080, 1920)) == 4 and hbma_auto_steps((360, 640)) == 2 and hbma_auto_steps((2160, 3840)) == 5
    assert hbma_auto_steps((72, 96)) == 1 and hbma_auto_steps((256, 256)) == 2
    bits = lambda d: int(np.float32(d).view(np.uint32))
    assert [bits(d) for d in interpolation_dists(2.0)] == [0x3f000000]
    assert [bits(1. - get_first_frame_idx_and_ratio(i, 2.5)[1]) for i in (1, 2, 3, 4)] == [0x3ecccccc, 0x3f4ccccd, 0x3e4ccccc, 0x3f19999a]
    assert [bits(d) for d in interpolation_dists(5.0)] == [0x3e4ccccc, 0x3ecccccc, 0x3f19999a, 0x3f4ccccd]
    assert math.isclose(1. - get_first_frame_idx_and_ratio(4, 2.0)[1], 0.)


def test_smoothing_tokens():
    from interpolatory_b200.smoothing import filter_mode_of, mean_filter, median_filter, weighted_mean_filter

    def weighted_mean_filter_ref(block):      # a reference-style callable is recognised by name
        return 0
    weighted_mean_filter_ref.__name__ = "weighted_mean_filter"
    assert filter_mode_of(None) == "none" and filter_mode_of(mean_filter) == "mean"
    assert filter_mode_of(median_filter) == "median" and filter_mode_of(weighted_mean_filter) == "weighted"
    assert filter_mode_of(weighted_mean_filter_ref) == "weighted" and filter_mode_of("median") == "median"
    with pytest.raises(ValueError):
        filter_mode_of(lambda b: b)


def test_sharding_plan():
    from interpolatory_b200.sharding import pairs_of_rank, runs_of_rank, frames_needed, band_split
    for n, w in ((16, 8), (17, 8), (5, 2), (3, 4)):
        allp = sorted(p for r in range(w) for p in pairs_of_rank(n, r, w))
        assert allp == list(range(n))
        spans = [runs_of_rank(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
    assert frames_needed([2, 3, 7]) == [2, 3, 4, 7, 8]
    bands = band_split(270, 8)          # 8K, 16-px blocks
    assert bands[0][0] == 0 and bands[-1][1] == 270 and max(b - a for a, b in bands) - min(b - a for a, b in bands) <= 1


def test_synthetic_sequence_is_deterministic():
    from interpolatory_b200.synth import make_sequence
    a = make_sequence(72, 96, 3, seed=5)
    b = make_sequence(72, 96, 3, seed=5)
    assert a.dtype == np.uint8 and a.shape == (3, 72, 96, 3) and np.array_equal(a, b)
    assert not np.array_equal(a[0], a[1])


def test_bench_reference_arm_runs_on_cpu():
    """bench.py --impl reference needs no GPU: it times the CPU restatement on the host cores."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "frames/s"


def test_band_plan_covers_every_needed_row():
    """8K band split: every frame row / MV block row a band needs is owned or arrives by a transfer."""
    from interpolatory_b200.tiling import BandPlan, HOLE_K
    for (H, W, bs, R, n) in ((4320, 7680, 16, 32, 8), (4320, 7680, 16, 32, 2), (360, 640, 8, 7, 4), (200, 320, 8, 16, 8), (1080, 1920, 8, 16, 3)):
        plan = BandPlan(H, W, bs, R, n)
        assert plan.block_rows[0][0] == 0 and plan.block_rows[-1][1] == plan.nbr
        assert plan.rows[-1][1] == H
        for kind, own, need in (("frame", plan.rows, plan.frame_need), ("mv", plan.block_rows, plan.mv_need)):
            tr = plan.transfers(kind)
            for dst in range(n):
                have = set(range(*own[dst]))
                for s, d, r0, r1 in tr:
                    if d == dst:
                        assert set(range(r0, r1)) <= set(range(*own[s]))       # the sender owns what it sends
                        have |= set(range(r0, r1))
                for span in need[dst]:
                    assert set(range(*span)) <= have, (kind, dst, span)
            # ME window and splat reach are inside the needed rows
        for i, (y0, y1) in enumerate(plan.rows):
            lo, hi = plan.frame_need[i][0]
            assert lo <= max(0, y0 - R - HOLE_K) and hi >= min(H, y1 + R + HOLE_K)
        if (H, n) == (4320, 8):
            f, m = plan.bytes_moved()
            assert f < 40e6 and m < 2e6                                         # halo traffic is latency-, not bandwidth-bound


def test_metrics_argument_checks_need_no_device():
    """metrics.py mirrors two skimage functions for one configuration and refuses the rest before touching the GPU."""
    import numpy as np
    import pytest
    from interpolatory_b200 import metrics as M
    a = np.zeros((16, 16, 3), np.uint8)
    with pytest.raises(ValueError):
        M.structural_similarity(a, a[:8], data_range=255, multichannel=True)
    with pytest.raises(NotImplementedError):
        M.structural_similarity(a.astype(np.float32), a.astype(np.float32), data_range=255, multichannel=True)
    with pytest.raises(NotImplementedError):
        M.peak_signal_noise_ratio(a, a, data_range=1.0)
    with pytest.raises(ValueError):
        M.frame_quality(a[:6], a[:6])
    assert M.psnr_from_mse(0.0) == float('inf')
    assert abs(M.psnr_from_mse(100.0) - 10 * np.log10(255.0 ** 2 / 100.0)) < 1e-12


def test_harness_helpers():
    from interpolatory_b200.util import sToMMSS, getETA
    assert sToMMSS(0) == '0:00' and sToMMSS(61) == '1:01' and sToMMSS(600) == '10:00'
    assert getETA(10, 0, 5) == '0:00' and getETA(10, 2, 6) == '0:20'


def test_frame_slot_lru_never_evicts_the_pair_partner():
    """Random access through get_interpolated_frame(idx) (the reference's VideoStream supports seeking backwards,
    src/VideoStream.py:61-96): after frames 5, 6, 7 are resident, asking for the pair (3, 4) must not put both
    frames into one device slot (round-1 advisor finding: eviction by lowest index did)."""
    from interpolatory_b200 import MEMCI

    class FakeDev:
        def __init__(self):
            self.uploads = []

        def upload_frame(self, slot, frame):
            self.uploads.append((slot, frame))

        def sync(self):
            pass

    it = MEMCI(60, me_mode='fs', mci_mode='bidir')
    it.video_stream = it._resident_stream = object()
    dev = FakeDev()
    for k in (5, 6, 7):
        it._frame_slot(dev, k, k)
    assert sorted(it._resident) == [5, 6, 7]
    for back in (3, 1, 6, 0):
        s = it._frame_slot(dev, back, back)
        t = it._frame_slot(dev, back + 1, back + 1, keep=(back,))
        assert s != t, (back, it._resident)
        assert it._resident[back] == s and it._resident[back + 1] == t
        # what the slots hold is what was last uploaded into them
        held = {}
        for slot, f in dev.uploads:
            held[slot] = f
        assert held[s] == back and held[t] == back + 1
    # a resident frame is not uploaded again and becomes the most recently used
    n = len(dev.uploads)
    it._frame_slot(dev, 1, 1)
    assert len(dev.uploads) == n and list(it._resident)[-1] == 1


def test_reader_video_stream_window():
    """ReaderVideoStream (reference: VideoStream, src/VideoStream.py:29-107): bounded forward-moving window, preload,
    fall-through reads behind the window / after a jump, duration from the metadata, nframes counted when the reader
    reports inf."""
    from interpolatory_b200.VideoStream import ReaderVideoStream

    class FakeReader:
        def __init__(self, n, nframes_meta):
            self.n, self.meta_n, self.reads = n, nframes_meta, []

        def get_meta_data(self):
            return {'fps': 24.0, 'size': (64, 48), 'duration': 1.7, 'nframes': self.meta_n}

        def count_frames(self):
            return self.n

        def get_data(self, i):
            self.reads.append(i)
            return ('frame', i % self.n)                 # loop=True readers wrap around

    r = FakeReader(40, float('inf'))
    vs = ReaderVideoStream(r, 3)
    assert vs.nframes == 40 and vs.duration == 1.7 and vs.fps == 24.0
    assert r.reads == [0, 1, 2]                           # preload
    assert vs.get_frame(1) == ('frame', 1) and r.reads == [0, 1, 2]
    assert vs.get_frame(3) == ('frame', 3) and r.reads[-1] == 3
    assert vs.get_frame(0) == ('frame', 0) and r.reads[-1] == 0      # behind the window: straight from the reader
    assert vs.get_frame(2) == ('frame', 2) and r.reads[-1] == 0      # still cached (window is 1..3)
    assert vs.get_frame(9) == ('frame', 9) and r.reads[-1] == 9      # a jump: no caching
    assert vs.get_frame(4) == ('frame', 4) and vs.get_frame(4) == ('frame', 4) and r.reads.count(4) == 1
    assert len(vs._window) == 3 and vs._first == 2
    r0 = FakeReader(5, 5)
    v0 = ReaderVideoStream(r0, 0)
    assert r0.reads == [] and v0.get_frame(2) == ('frame', 2) and v0.get_frame(2) == ('frame', 2) and r0.reads == [2, 2]
    import pytest
    with pytest.raises(Exception):
        ReaderVideoStream(r0, -1)


def test_multigpu_partition_is_contiguous_and_complete():
    from interpolatory_b200.multigpu import rank_pairs
    pairs = [0, 1, 2, 5, 6, 9, 10, 11, 12]
    for world in (1, 2, 3, 4, 8, 16):
        parts = [rank_pairs(pairs, r, world) for r in range(world)]
        assert sum(parts, []) == pairs
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_hbma_reference_undefined_regime_is_named(oracle):
    """The geometry where the reference's HBMA reads out of bounds (single block row / column at some level,
    ME/hbma.py:108-115): the product's predicate equals the oracle's on random geometry, and the class path warns."""
    import warnings
    from interpolatory_b200.util import hbma_reference_undefined, warn_if_hbma_reference_undefined
    rng = np.random.default_rng(7)
    hits = 0
    for _ in range(2000):
        H, W = int(rng.integers(4, 400)), int(rng.integers(4, 400))
        bs = int(rng.choice([4, 8, 16, 32])); steps = int(rng.integers(0, 5)); mbs = int(rng.choice([2, 4, 8]))
        got = hbma_reference_undefined(H, W, bs, steps, mbs)
        assert got == oracle.hbma_undefined_in_reference(H, W, bs, steps, mbs), (H, W, bs, steps, mbs)
        hits += got
    assert 0 < hits < 2000
    assert not hbma_reference_undefined(1080, 1920, 8, 4, 4) and not hbma_reference_undefined(360, 640, 8, 2, 4)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        warn_if_hbma_reference_undefined(1080, 1920, 8, 4, 4)
        assert not w
        warn_if_hbma_reference_undefined(24, 200, 8, 2, 4)
        assert len(w) == 1 and issubclass(w[0].category, RuntimeWarning) and "hbma.py:108-115" in str(w[0].message)
