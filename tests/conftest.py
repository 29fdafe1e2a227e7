import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    # The built artefacts are git-ignored: on a fresh checkout build them once (nvcc cross-compiles
    # sm_100a without a GPU) so that collection does not depend on who ran build() first.
    lib = os.path.join(ROOT, "snpgenie_b200", "libsnpgenie_cuda.so")
    glue = os.path.join(ROOT, "perl", "SNPGenieCUDA", "lib", "auto", "SNPGenieCUDA", "SNPGenieCUDA.so")
    if not (os.path.exists(lib) and os.path.exists(glue) and os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so"))):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
