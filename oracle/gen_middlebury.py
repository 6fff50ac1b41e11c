"""Middlebury fixtures (build container only): the reference's own 12 evaluation triples and the
outputs of the UNMODIFIED reference on them.

    python oracle/gen_middlebury.py

Writes
  tests/golden/middlebury_pairs.npz    the PNG files of benchmarks/middlebury/*/frame10.png, frame11.png and the
                                       ground truth frame10i11.png, byte for byte (decoded with PIL in the tests:
                                       the reference reads them with imageio.imread, src/Benchmark.py:39-42, which
                                       yields the same uint8 RGB array)
  tests/golden/middlebury_ref_sha.npz  SHA-256 of `MEMCI(2, **settings).get_benchmark_frame(frame10, frame11)`
                                       (the call src/Benchmark.py:44 makes) for fs / hbma x unidir / bidir / bidir2,
                                       produced by the reference classes (numba 0.65.0 / NumPy 2.3.5 / SciPy 1.18.1),
                                       plus the block-level full-search fields of both directions.

640x480 x8, 584x388 x3, 420x380 x1: 388, 380 and 420 are not multiples of 8, so the zero-pad / partial-block
paths of every kernel are exercised on real image content.
"""
import glob
import hashlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
BASE = "/root/reference/Simulator/python/benchmarks/middlebury"

SETTINGS = {
    "fs_unidir": dict(me_mode='fs', mci_mode='unidir'),
    "fs_bidir": dict(me_mode='fs', mci_mode='bidir'),
    "fs_bidir2": dict(me_mode='fs', mci_mode='bidir2'),
    "hbma_unidir": dict(me_mode='hbma', mci_mode='unidir'),
    "hbma_bidir": dict(me_mode='hbma', mci_mode='bidir'),
    "hbma_bidir2": dict(me_mode='hbma', mci_mode='bidir2'),
    "fs16_bidir_median": dict(me_mode='fs', mci_mode='bidir', block_size='8', target_region='16', filter_mode='median'),
}


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest() + f":{a.dtype}:{'x'.join(map(str, a.shape))}"


def main():
    from PIL import Image
    ns = ref_harness.load()
    pairs, shas = {}, {}
    names = sorted(os.path.basename(os.path.dirname(p)) for p in glob.glob(f"{BASE}/*/frame10i11.png"))
    assert len(names) == 12, names
    pairs["names"] = np.array(names)
    for name in names:
        imgs = {}
        for f in ("frame10", "frame11", "frame10i11"):
            raw = open(f"{BASE}/{name}/{f}.png", "rb").read()
            pairs[f"{name}/{f}"] = np.frombuffer(raw, np.uint8)
            imgs[f] = np.array(Image.open(io.BytesIO(raw)).convert("RGB"))     # writable, like imageio.imread
        a, b = imgs["frame10"], imgs["frame11"]
        shas[f"{name}/frames_sha"] = np.array(sha(np.stack([a, b])))
        fwd = ns.fs.get_motion_vectors(8, 7, a, b)
        bwd = ns.fs.get_motion_vectors(8, 7, b, a)
        shas[f"{name}/fs_fwd"] = np.ascontiguousarray(fwd[::8, ::8])
        shas[f"{name}/fs_bwd"] = np.ascontiguousarray(bwd[::8, ::8])
        for key, st in SETTINGS.items():
            it = ns.MEMCI(2, **dict(st))
            out = np.asarray(it.get_benchmark_frame(a, b))
            assert out.dtype == np.uint8 and out.shape == a.shape
            shas[f"{name}/{key}"] = np.array(sha(out))
            print(name, key, shas[f"{name}/{key}"][()][:16], flush=True)
    shas["settings"] = np.array(repr({k: sorted(v.items()) for k, v in SETTINGS.items()}))
    np.savez(os.path.join(OUT, "middlebury_pairs.npz"), **pairs)
    np.savez_compressed(os.path.join(OUT, "middlebury_ref_sha.npz"), **shas)
    for f in ("middlebury_pairs.npz", "middlebury_ref_sha.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
