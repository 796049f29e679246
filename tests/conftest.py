import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The CPU tests drive the reference's Python loops (thousands of tiny ATen ops): with one OpenMP team per op on a
    # shared host they spend their time in thread hand-offs (measured: the config-4 case 592 s with 8 threads, 75 s
    # with 1).  Two threads keep the few large convolutions reasonable.  FP8FQ_TEST_THREADS overrides.
    # (Not on the GPU box: there the parity tests run the CPU oracle over full-size tensors and keep torch's default.)
    if "FP8FQ_TEST_THREADS" in os.environ or not torch.cuda.is_available():
        torch.set_num_threads(int(os.environ.get("FP8FQ_TEST_THREADS", "2")))


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Make sure libfp8fq.so and the test-only CPU checkers exist (nvcc / gcc cross-compile here)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def oracle_c(built):
    import ctypes

    return ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libfp8_oracle_c.so"))


@pytest.fixture(scope="session")
def host_emul(built):
    import ctypes

    return ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libfp8_host_emul.so"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def bits(t):
    return t.contiguous().view(torch.int32)


def ulp_diff(a, b):
    """Max distance in units of fp32 representation steps between same-signed floats (NaN==NaN -> 0)."""
    ia, ib = bits(a).to(torch.int64), bits(b).to(torch.int64)
    # map the sign-magnitude float ordering onto a monotone integer line
    ia = torch.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = torch.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    d = (ia - ib).abs()
    both_nan = torch.isnan(a) & torch.isnan(b)
    d = torch.where(both_nan, torch.zeros_like(d), d)
    one_nan = torch.isnan(a) ^ torch.isnan(b)
    d = torch.where(one_nan, torch.full_like(d, 1 << 40), d)
    return d


@pytest.fixture()
def simdev(built):
    """Puts the host SIMULATION of the kernels (tests/host_sim: the product's own .cu compiled by g++) behind the
    product's Python layer for the duration of one test, so that module / estimator / whole-model logic can be
    checked on CPU tensors without a GPU (tests/host_sim/harness.py).  TEST ONLY: the product has no CPU path --
    without this fixture a CPU tensor raises Fp8fqError (tests/test_capi_symbols.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "host_sim"))
    try:
        from harness import install_simulation
    finally:
        sys.path.pop(0)
    restore = install_simulation()
    try:
        yield torch.device("cpu")
    finally:
        restore()
