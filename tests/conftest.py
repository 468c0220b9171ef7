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


def pytest_sessionstart(session):
    """The C-ABI library is built in-tree (nvcc cross-compiles without a GPU); make sure it exists so that the
    ABI tests and the GPU tests exercise the real thing.  __graft_entry__.build() does the same."""
    try:
        from hopper_mpc_inertial_b200 import build
        build.build_lib(force=False)
    except Exception as e:  # pragma: no cover - reported by the tests that need the library
        print(f"warning: could not build libhmpc_b200.so: {e}")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def reference():
    """The reference's own modules through oracle/refshim.py (build container only)."""
    from oracle import refshim
    if not refshim.available():
        pytest.skip("/root/reference not present")
    return refshim.load()


def normalised_oracle_qp(qp, prm):
    """Oracle condensed QP with the height rows scaled as the device stores them (include/hmpc.h)."""
    N = prm.N
    n = 6 * N
    A = qp["A"].copy()
    lo = np.maximum(qp["l"], -1e30)
    hi = np.minimum(qp["u"], 1e30)
    for k in range(2, N):
        r = n + 4 * N + k
        s = prm.mpc_dt ** 2 * (k - 1) / prm.m
        A[r] /= s
        if lo[r] > -1e26:
            lo[r] /= s
    return A, lo, hi


def u_tol(Uo):
    """Parity bound of BASELINE.json north_star: 1e-5 abs + 1e-4 rel (FP64)."""
    return 1e-5 + 1e-4 * np.abs(Uo)
