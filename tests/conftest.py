import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """A fresh checkout has no built libraries (they are git-ignored): build them once, the way the driver's build() does.
    On the GPU box the libraries travel with the snapshot, so nothing happens there."""
    lib = os.path.join(ROOT, "hector_simulation_b200", "libhector_mpc_b200.so")
    if os.path.exists(lib):
        return
    try:
        import __graft_entry__

        __graft_entry__.build()
    except Exception as e:  # the tests that need the library then fail with their own, more specific message
        print("conftest: __graft_entry__.build() failed: %r" % (e,))


def load_golden(name):
    from hector_simulation_b200.scenarios import UPDATE_DTYPE

    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["records"] = np.ascontiguousarray(d["records"]).view(UPDATE_DTYPE).reshape(-1)
    d["horizon"] = int(d["horizon"])
    return d


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Skips solver-dependent checks if qpOASES is not linked."""
    from oracle import oracle_py

    oracle_py.lib()
    return oracle_py


def rel_err(a, b, width=None):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if width is not None:
        a, b = a[:, :width], b[:, :width]
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-9)
