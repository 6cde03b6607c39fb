import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "em-spec_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib_built():
    """Build the C-ABI library once per session (nvcc cross-compiles without a GPU)."""
    import build_emspec as ems_build  # em-spec_b200/build_emspec.py
    return ems_build.build()
