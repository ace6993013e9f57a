import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def build_lib():
    """libfhestr_engine.so, built in-tree (nvcc cross-compiles without a GPU)."""
    from fhestring_b200 import build
    return build.build()


SMALL_N = 8


@pytest.fixture(scope="session")
def small_oracle():
    from oracle.tfhe_oracle import Oracle, PARAM_MESSAGE_2_CARRY_2_KS_PBS as P
    p = dict(P)
    p.update(n=SMALL_N)
    o = Oracle(**p)
    return o, o.keygen(11)


@pytest.fixture(scope="session")
def full_oracle():
    from oracle.tfhe_oracle import Oracle, PARAM_MESSAGE_2_CARRY_2_KS_PBS as P
    o = Oracle(**P)
    return o, o.keygen(1)


def monomial_mul(poly: np.ndarray, e: int) -> np.ndarray:
    """X^e * poly, negacyclic, e in [0, 2N) -- numpy restatement used by several tests."""
    N = poly.shape[-1]
    j = np.arange(N)
    q = (j - e) % (2 * N)
    v = poly[..., q % N]
    with np.errstate(over="ignore"):
        return np.where(q >= N, np.uint64(0) - v, v)
