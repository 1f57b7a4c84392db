"""pytest configuration: markers, import paths, shared fixtures.

`-m "not gpu"` : oracle vs the reference's golden vectors, snapshot codec, host logic, C-ABI symbol/ABI checks,
                 world_size-2 gloo test of the sharding scheme.  No GPU, a few minutes at most.
`-m gpu`       : the parity tests proper — everything goes through the C ABI (liboon_gpu.so) on cuda:0 and is
                 checked against the oracle / golden fixtures.  Nothing here reads /root/reference.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) device; run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "multigpu: needs at least 2 GPUs")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def kat1(golden_dir):
    from out_of_nothing_b200 import snapshot
    return (snapshot.load(os.path.join(golden_dir, "kat1_start.state")),
            snapshot.load(os.path.join(golden_dir, "kat1_end.state")))
