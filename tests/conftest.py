import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "fullsize: BASELINE-size parity cases (minutes of CPU oracle time); deselect with -m 'gpu and not fullsize'")


def need_gpu():
    """GPU tests skip on a box without a device (so a plain `pytest tests` is a usable CPU gate) unless
    CALES_REQUIRE_GPU=1, in which case a missing device is a failure (there is no CPU fallback to test)."""
    import torch
    if not torch.cuda.is_available():
        if os.environ.get("CALES_REQUIRE_GPU") == "1":
            pytest.fail("no CUDA device visible: GPU tests must run on the B200 box (no CPU fallback exists)")
        pytest.skip("no CUDA device visible")


@pytest.fixture(scope="session")
def lib():
    from cales_b200 import lib as L
    return L.load()


@pytest.fixture(params=["strict", "fma"])
def arith(request):
    """Both arithmetic variants of the library run every GPU parity test: "strict" (-fmad=false, bit-identical to the
    non-contracting oracle: stencil kernels compare with array_equal) and "fma" (the product build, contraction on as in
    the reference's own GPU build: north-star tolerances).  Sets the default variant of cales_b200.lib.load()."""
    from cales_b200 import lib as L
    old = L.DEFAULT_ARITH
    L.DEFAULT_ARITH = request.param
    yield request.param
    L.DEFAULT_ARITH = old


@pytest.fixture(autouse=True)
def _default_torch_stream(request):
    """A Simulation with graph replay makes its own stream torch's current stream until close(); a test that fails before
    close() must not leak that (non-blocking) stream into the next test, whose uploads would then race the library's work."""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    import torch
    if torch.cuda.is_available():
        torch.cuda.synchronize()
        torch.cuda.set_stream(torch.cuda.default_stream())
    yield
    if torch.cuda.is_available():
        torch.cuda.synchronize()
        torch.cuda.set_stream(torch.cuda.default_stream())
