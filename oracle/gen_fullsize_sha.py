"""SHA-256 fixtures of the CPU oracle at BASELINE.json's full sizes (cfg4: 4K 16x16 +-32, cfg5: 8K 16x16 +-32).

    python oracle/gen_fullsize_sha.py            # ~5 min on 8 cores

TEST INFRASTRUCTURE.  The oracle (oracle/memci_oracle.c, pinned bit-for-bit against the unmodified reference by
validate_vs_reference.py) searches both directions of one seeded full-size pair and interpolates the middle frame;
the digests of the block-level fields and of the frames go to tests/golden/fullsize_sha.npz, so that the -m gpu
suite can check the CUDA path at 8K without spending two minutes of host time per run.  The cfg4 test also runs
the oracle live (tests/test_gpu_parity.py::test_cfg4_4k_full_size_vs_oracle).
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from interpolatory_b200.synth import make_pair_tiled  # noqa: E402

CASES = {"cfg4": (2160, 3840, 16, 32, 4), "cfg5": (4320, 7680, 16, 32, 5)}     # H, W, block, range, seed


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest() + f":{a.dtype}:{'x'.join(map(str, a.shape))}"


def main():
    O.build()
    d = {}
    for name, (H, W, bs, R, seed) in CASES.items():
        t0 = time.time()
        a, b = make_pair_tiled(H, W, seed)
        d[f"{name}/frames_sha"] = np.array(sha(np.stack([a, b])))
        fwd, bwd = O.fs_blocks(bs, R, a, b), O.fs_blocks(bs, R, b, a)
        d[f"{name}/fwd_sha"], d[f"{name}/bwd_sha"] = np.array(sha(fwd)), np.array(sha(bwd))
        fd, bd = O.expand_dense(fwd, bs, H, W), O.expand_dense(bwd, bs, H, W)
        d[f"{name}/bidir_sha"] = np.array(sha(O.bidir_helper(a, b, fd, bd, np.float32(0.5))))
        d[f"{name}/unidir_sha"] = np.array(sha(O.unidir_helper(a, b, fd, np.float32(0.5))))
        d[f"{name}/work"] = np.array(O.fs_work(H, W, bs, R), np.int64)
        print(name, f"{time.time() - t0:.0f} s", d[f"{name}/bidir_sha"][()][:16], flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fullsize_sha.npz"), **d)


if __name__ == "__main__":
    main()
