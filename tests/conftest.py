"""pytest configuration: the `gpu` marker, import paths and shared fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "greb-climate-model_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_sessionstart(session):
    """Keep the in-tree CUDA library in step with its sources (make is a no-op when it is): a stale
    libgreb_b200.so would test yesterday's kernels.  Never fatal here — the ABI tests fail loudly if
    the library is missing."""
    import shutil
    import subprocess
    if os.environ.get("GREB_B200_LIB") or shutil.which("nvcc") is None or shutil.which("make") is None:
        return
    r = subprocess.run(["make", "-C", os.path.join(PKG, "csrc")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True)
    if r.returncode != 0:
        print("conftest: rebuilding libgreb_b200.so failed:\n" + r.stdout[-2000:])


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running")


@pytest.fixture(scope="session")
def forcing():
    from greb_b200 import synth
    cache = os.environ.get("GREB_FORCING_CACHE", "/tmp/greb_b200_cache")
    return synth.cached_forcing(cache_dir=cache)


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as om
    om.build()
    return om


@pytest.fixture()
def orc(oracle_mod, forcing):
    return oracle_mod.Oracle(forcing)
