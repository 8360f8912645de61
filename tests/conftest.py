import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "dynamic-visual-slam_b200", "python"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    import c_oracle
    c_oracle.lib()
    return c_oracle


@pytest.fixture(scope="session")
def refx(built):
    """The REFERENCE's own ORBextractor.cpp, compiled unmodified (oracle/_ref, built in the container where /root/reference exists;
    the prebuilt library travels to the GPU box).  None where it is unavailable — tests/test_gpu_parity.py::test_reference_binary_present
    makes that loud on the GPU box."""
    import ref_oracle
    if not ref_oracle.available():
        return None
    return ref_oracle.RefExtractor()
