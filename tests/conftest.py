import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


def golden_names():
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))) if not n.startswith(("post_", "aug_", "deconv_")))


def post_golden_names():
    """Fixtures of the test-time post-processing (K6), written by `python -m oracle.make_golden --post`."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "post_*.npz")))


def aug_golden_names():
    """Fixtures of the sample preparation (K7 / K8), written by `python -m oracle.make_golden --aug` from the reference's own
    DatasetLoader.__getitem__."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "aug_*.npz")))


def load_aug_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    g["pairs"] = tuple((int(a), int(b)) for a, b in g["pairs"])
    for k in ("h", "w", "J", "depth_dim"):
        g[k] = int(g[k])
    g["augs"] = [(float(a[0]), float(a[1]), bool(a[2]), [float(a[3]), float(a[4]), float(a[5])]) for a in g["aug"]]
    return g


def load_post_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    g["pairs"] = tuple((int(a), int(b)) for a, b in g["pairs"])
    for k in ("B", "J", "D", "H", "W", "root"):
        g[k] = int(g[k])
    g.setdefault("flipped", None)
    return g


def deconv_golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "deconv_*.npz")))


def load_deconv_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def load_golden(name):
    """Golden produced by oracle/make_golden.py from the unmodified reference.  Inputs of the big cases are
    regenerated from (dist, shape, seed) and checked against the stored sha256."""
    import hashlib
    from oracle import inputs
    z = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    g = {k: (v.item() if v.shape == () else v) for k, v in z.items()}
    for k in ("B", "J", "D", "H", "W", "seed"):
        g[k] = int(g[k])
    g["dist"] = str(g["dist"])
    if "heat" not in g:
        g["heat"] = inputs.make_heat(g["dist"], g["B"], g["J"], g["D"], g["H"], g["W"], g["seed"])
    sha = np.frombuffer(hashlib.sha256(g["heat"].tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, g["heat_sha"]), "regenerated input differs from the one the golden was made from"
    g["name"] = name
    return g


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree C-ABI library; built on demand (nvcc cross-compiles without a GPU)."""
    from ihpr_b200 import build
    return build.build_library()
