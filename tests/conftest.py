import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """The product library must exist; tests never fall back to anything else."""
    from imd_b200 import api
    if not os.path.exists(api.LIB_PATH):
        import subprocess
        subprocess.check_call(["make", "-C", ROOT, "-s"])
    return api.load_library()
