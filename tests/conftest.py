import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the fuzz tests deliberately visit tiny frames where the reference HBMA is undefined (util.hbma_reference_undefined)
    config.addinivalue_line("filterwarnings", "ignore:HBMA on a:RuntimeWarning")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/memci_oracle.c): test infrastructure, never on the product path."""
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load libmemci_b200.so; compiling needs nvcc, not a GPU."""
    from interpolatory_b200 import build as B
    B.build()
    from interpolatory_b200 import _native
    return _native.lib()


@pytest.fixture(scope="session")
def middlebury():
    """The reference's 12 Middlebury evaluation triples (benchmarks/middlebury/*/frame10|frame11|frame10i11.png,
    committed byte for byte in tests/golden/middlebury_pairs.npz by oracle/gen_middlebury.py) decoded to uint8 RGB,
    as imageio.imread delivers them to src/Benchmark.py:39-42.  Returns {name: (frame10, frame11, ground truth)}."""
    import io
    from PIL import Image
    z = np.load(os.path.join(GOLDEN, "middlebury_pairs.npz"), allow_pickle=False)

    def dec(key):
        return np.array(Image.open(io.BytesIO(z[key].tobytes())).convert("RGB"))
    return {str(n): tuple(dec(f"{n}/{f}") for f in ("frame10", "frame11", "frame10i11")) for n in z["names"]}
