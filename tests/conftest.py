import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gtn():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import grassmanntn_b200 as g
    return g


@pytest.fixture
def host_tables(monkeypatch):
    """run the product's host code without a GPU: buffers on the host, every sign+permute launch replaced by the
    numpy emulation of the kernel's addressing / sign rule"""
    import torch
    from test_tables_cpu import emulate
    from grassmanntn_b200 import _engine as E, _ops

    class HostPlan:
        def __init__(self, jobs):
            self.jobs = jobs

        def run(self, src, dst, scale=1.0):
            s, d = src.numpy(), dst.numpy()
            for f, tabs in self.jobs:
                emulate(f, tabs, s, d, scale)
    saved = dict(E._plan_cache)
    E._plan_cache.clear()
    monkeypatch.setattr(E, "PermutePlan", HostPlan)
    monkeypatch.setattr(_ops, "PermutePlan", HostPlan)
    monkeypatch.setattr(E, "require_cuda", lambda: torch.device("cpu"))
    monkeypatch.setattr(_ops, "require_cuda", lambda: torch.device("cpu"))
    yield
    E._plan_cache.clear()
    E._plan_cache.update(saved)
