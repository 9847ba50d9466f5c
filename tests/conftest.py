import os
import sys

# P in-process ranks x (main + comm + reducer) streams must not alias onto the default 8 hardware queues: a
# device-side flag wait (p2p transport) queued in front of the put it waits for would deadlock.  Must be set
# before CUDA initialises.  (One process per GPU -- the deployment shape -- uses 3-4 streams and does not need it.)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# Same reason, second mechanism: CUDA loads kernel code lazily at first launch and the load synchronises the
# context, so a rank's first-ever launch of ANY kernel would block behind another rank's spinning flag wait.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def built():
    """Make sure libbnsgcn.so and the oracle's C library exist (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build()
    return True
