"""Import harness for the UNMODIFIED reference (lhl2617/Interpolatory, Simulator/python).

TEST INFRASTRUCTURE ONLY.  Used in the build container to (a) validate the C
restatement in oracle/memci_oracle.c and (b) generate the golden fixtures under
tests/golden/ (see oracle/gen_golden.py).  /root/reference does not exist on
the GPU box, so nothing in tests -m gpu / smoke() / bench.py imports this file.

Rules:
  * never import the reference in place: numba's cache=True would write
    __pycache__/*.nbi into /root/reference.  We copy Simulator/python/src to a
    scratch directory and point NUMBA_CACHE_DIR at scratch as well.
  * imageio / skimage are not installed; the arithmetic never calls them, so
    5-line stub packages are enough.
"""
import os
import shutil
import sys
import tempfile

REFERENCE_SRC = "/root/reference/Simulator/python/src"
_SCRATCH = os.environ.get("MEMCI_REF_SCRATCH", os.path.join(tempfile.gettempdir(), "memci_ref_scratch"))


def available() -> bool:
    return os.path.isdir(REFERENCE_SRC)


def load():
    """Return a namespace object with the reference's hot-path callables."""
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    root = os.path.join(_SCRATCH, "ref")
    stubs = os.path.join(_SCRATCH, "stubs")
    if not os.path.isdir(os.path.join(root, "src")):
        os.makedirs(root, exist_ok=True)
        shutil.copytree(REFERENCE_SRC, os.path.join(root, "src"))
    os.makedirs(os.path.join(stubs, "imageio"), exist_ok=True)
    os.makedirs(os.path.join(stubs, "skimage"), exist_ok=True)
    with open(os.path.join(stubs, "imageio", "__init__.py"), "w") as f:
        f.write("def _no(*a, **k):\n    raise RuntimeError('imageio stub')\n"
                "imread = imwrite = get_reader = get_writer = _no\n")
    for name in ("__init__.py", "metrics.py"):
        with open(os.path.join(stubs, "skimage", name), "w") as f:
            f.write("")
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(_SCRATCH, "nbcache"))
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    for p in (root, stubs):
        if p not in sys.path:
            sys.path.insert(0, p)

    import src.Globals as G
    for k in G.debug_flags:
        G.debug_flags[k] = False

    class NS:
        pass

    ns = NS()
    from src.Interpolators.MEMCI.ME import full_search as fs_mod, hbma as hbma_mod, tss as tss_mod
    from src.Interpolators.MEMCI.smoothing.threeXthree_mv_smoothing import smooth
    from src.Interpolators.MEMCI.smoothing.mean_filter import mean_filter
    from src.Interpolators.MEMCI.smoothing.median_filter import median_filter
    from src.Interpolators.MEMCI.smoothing.weighted_mean_filter import weighted_mean_filter
    from src.Interpolators.MEMCI.MCI import Unidirectional as uni_mod, Bidirectional as bi_mod
    from src.Interpolators.MEMCI.MCI.Interpolator import MEMCI
    from src.VideoStream import BaseVideoStream, BenchmarkVideoStream
    from src.util import get_first_frame_idx_and_ratio
    ns.fs = fs_mod
    ns.hbma = hbma_mod
    ns.tss = tss_mod
    ns.smooth = smooth
    ns.filters = {"mean": mean_filter, "median": median_filter, "weighted": weighted_mean_filter, "none": None}
    ns.uni = uni_mod
    ns.bi = bi_mod
    ns.MEMCI = MEMCI
    ns.BaseVideoStream = BaseVideoStream
    ns.BenchmarkVideoStream = BenchmarkVideoStream
    ns.get_first_frame_idx_and_ratio = get_first_frame_idx_and_ratio
    return ns
