import json
import os
import sys

import numpy as np
import pytest

# tests/test_gpu_routed.py runs several ranks on one GPU, each with three streams of its own, and a
# rank's drain kernel waits for the epoch another rank's route kernel publishes: the two streams must
# not share a hardware work queue (the default is 8 queues per process).  Read at CUDA initialisation.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the compiled reference (tests/golden/make_golden.py)."""
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        meta = json.load(f)
    return {
        "meta": meta,
        "fields": dict(np.load(os.path.join(GOLDEN, "fields.npz"))),
        "advect": dict(np.load(os.path.join(GOLDEN, "advect_small.npz"))),
        "splat": dict(np.load(os.path.join(GOLDEN, "splat_small.npz"))),
        "c1_pts": np.load(os.path.join(GOLDEN, "c1_pts.npy")),
        "next": dict(np.load(os.path.join(GOLDEN, "next_small.npz"))),
        "c2_pts": np.load(os.path.join(GOLDEN, "c2_pts.npy")),
        "c3_corner": np.load(os.path.join(GOLDEN, "c3_octaves_corner.npz"))["total"],
    }


@pytest.fixture(scope="session")
def c2_frames():
    """Frames 0, 199 and 399 of README example6 from the reference CLI (4000 x 2000 u8 each)."""
    return dict(np.load(os.path.join(GOLDEN, "c2_frames.npz")))


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.lib()  # builds liboracle.so on first use
    return oracle


@pytest.fixture(scope="session")
def ctx():
    """A libclumpy_cuda context on cuda:0.  No fallback: errors propagate."""
    import clumpy_b200
    c = clumpy_b200.Context(0)
    yield c
    c.close()


def frame_agreement(a, b):
    """(fraction of pixels within 1 LSB, max abs difference) between two u8 images."""
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return float(np.mean(d <= 1)), int(d.max()) if d.size else 0
