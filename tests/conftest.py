import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _use_emulated_abi():
    """PBK_TEST_EMULATED_ABI=1 (set by tests/test_abi_emulated_cpu.py for a child pytest only): the `-m gpu` tests are pointed
    at the C ABI compiled for the host (tests/cpu_emul/cuda_rt_shim.h) -- a check of pbk_api.cu's host logic in the GPU-less
    container.  TEST ONLY: the product never looks at this variable."""
    import emul_helper
    from platanus_b_b200 import build
    build.LIB = emul_helper.abi_lib_path()
    build.build_cli = lambda force=False: emul_helper.abi_cli_path()      # pbk_assemble linked against the same library


def pytest_configure(config):
    if os.environ.get("PBK_TEST_EMULATED_ABI"):
        _use_emulated_abi()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: larger CPU cases")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O
