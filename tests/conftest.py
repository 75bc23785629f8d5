import os, subprocess, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libdmnd_oracle.so")
PRODUCT_LIB = os.path.join(ROOT, "diamond_b200", "libdmnd_b200.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "diamond")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The TEST-ONLY library: host pipeline linked over the CPU restatement of the kernels (oracle/dmnd_oracle.c)."""
    subprocess.run(["make", "-s", "oracle"], cwd=ROOT, check=True)
    from diamond_b200 import api
    return api.load(ORACLE_LIB)


@pytest.fixture(scope="session")
def product_lib():
    from diamond_b200 import api
    if not os.path.exists(PRODUCT_LIB):
        subprocess.run(["make", "-s", "lib"], cwd=ROOT, check=True)
    return api.load(PRODUCT_LIB)


_cache = {}


def workload_blocks(name):
    """(w, q_raw, q_limits, r_raw, r_limits) of a named synthetic workload, cached per session."""
    if name not in _cache:
        from diamond_b200 import api, synth
        w = synth.named(name)
        q_raw, q_lim = api.block_image(w["q_letters"], w["q_off"])
        r_raw, r_lim = api.block_image(w["db_letters"], w["db_off"])
        _cache[name] = (w, q_raw, q_lim, r_raw, r_lim)
    return _cache[name]
