import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build the CUDA library once (nvcc
    cross-compiles without a GPU).  The oracle libraries are built on demand by oracle/sworacle.py."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "schwarzwald_b200", "libswgpu.so")
    if not os.path.exists(lib) and not os.environ.get("SWGPU_LIB") and shutil.which("nvcc"):
        subprocess.run(["bash", os.path.join(ROOT, "build_native.sh")], check=True)


@pytest.fixture(scope="session")
def port_oracle():
    from oracle import sworacle
    return sworacle.Oracle("port")


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle import sworacle
    if not sworacle.have_ref():
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    return sworacle.Oracle("ref")
