import os
import shutil
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # never test a stale library: rebuild (mtime check) before anything imports it
    from aligngraph_b200 import build
    build.build()


@pytest.fixture(scope="session")
def harness():
    from oracle import harness as h
    h.build_tools()
    return h


@pytest.fixture()
def workdir():
    d = tempfile.mkdtemp(prefix="ag_test_")
    yield d
    shutil.rmtree(d, ignore_errors=True)


def golden_dir(name):
    return os.path.join(ROOT, "tests", "golden", name)


def compare_with_golden(harness, work, name):
    """Byte-compare the three per-unit outputs in work/tmp with the committed outputs of the real reference."""
    g = golden_dir(name)
    n = harness.n_units(work)
    assert n > 0
    for u in range(n):
        for pat in harness.UNIT_FILES:
            fn = pat.format(u)
            with open(os.path.join(g, fn), "rb") as f:
                want = f.read()
            with open(os.path.join(work, "tmp", fn), "rb") as f:
                got = f.read()
            assert got == want, f"{name}: {fn} differs from the reference's output"
