"""Test configuration.  `-m "not gpu"` runs here on CPU (oracle vs golden fixtures, host logic, the
C-ABI library's symbols); `-m gpu` runs on a B200 and is the parity suite proper (CUDA path through
the C-ABI vs the oracle and the golden fixtures written from the reference by oracle/make_golden.py).
"""
import json
import os
import sys
from hashlib import sha256

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture()
def fixed_random_seed(request) -> int:
    """Same scheme as the reference's tests/conftest.py:21-23."""
    return abs(int.from_bytes(sha256(request.node.name.encode("utf-8")).digest()[:8]))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {k: z[k] for k in z.files}
    if "meta" in out:
        out["meta"] = json.loads(str(out["meta"]))
    return out


def golden_batch(meta, dtype=None):
    """Regenerate the inputs of a solve fixture from its seed and verify the stored digest."""
    import davo_b200

    np_dt = dtype or (np.float64 if meta["dtype"] == "float64" else np.float32)
    batch = getattr(davo_b200.synthetic, meta["generator"])(dtype=np_dt, **meta["generator_kwargs"])
    assert batch.digest() == meta["digest"], "synthetic generator drifted from the golden fixture's inputs"
    return batch
