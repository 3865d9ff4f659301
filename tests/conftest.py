"""Shared fixtures.  `-m gpu` tests need a B200 and call the product through the C ABI; everything else runs on CPU."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

# The in-process band tests run several bands on ONE device, each on its own streams, and a band's stream may spin on a
# flag another band's stream has to set: the streams must not share a hardware queue (8 by default).  One process per GPU,
# the deployment this emulates, has no such coupling.  Must be set before the process creates its CUDA context.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU oracle cases")


@pytest.fixture(scope="session")
def oracle():
    """The checker: the compiled, unmodified reference when it is present (oracle/_ref), else the C port."""
    from oracle import oracle as orc

    if orc.have("ref"):
        return orc.Oracle("ref")
    if not orc.have("port"):
        orc.build("port")
    return orc.Oracle("port")


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import oracle as orc

    if not orc.have("port"):
        orc.build("port")
    return orc.Oracle("port")


@pytest.fixture(scope="session")
def engine():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from stereo_match_traditional_b200.api import Engine

    e = Engine(0)
    yield e
    e.close()


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) over the volume, and the fraction of exactly equal cells."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = np.maximum(np.abs(b), 1e-30)
    r = np.abs(a - b) / denom
    r[(a == b)] = 0.0
    return float(r.max()), float((a == b).mean())
