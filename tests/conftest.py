import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def oracle():
    from oracle.loader import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled into oracle/_ref (present in the dev container and shipped to the GPU box as a
    prebuilt file); tests that need it are skipped where it is missing."""
    from oracle import loader
    if not os.path.exists(loader.REF_SO):
        if os.path.isdir("/root/reference"):
            loader.build(("ref",))
        else:
            pytest.skip("oracle/_ref/libcvtt_ref.so not built")
    return loader.Reference()


def first_mismatch(a, b):
    bad = np.nonzero((a != b).any(axis=1))[0]
    if len(bad) == 0:
        return "equal"
    i = int(bad[0])
    return "%d/%d blocks differ; first at %d: %s vs %s" % (len(bad), len(a), i, bytes(a[i]).hex(), bytes(b[i]).hex())
