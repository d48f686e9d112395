import os
import sys

import pytest

# several logical shards of a sharded build share one GPU in the tests; each needs its own hardware queue, or a
# shard spinning at a barrier can sit in front of the kernels of the shard it waits for (default: 8 queues)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import __graft_entry__ as ge
    return ge.load_oracle()


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    p = ge.load_package()
    p.build_native()
    return p
