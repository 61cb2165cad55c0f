"""The C-ABI library loads and exports every symbol include/aligngraph_b200.h declares; no compute calls without a GPU."""
import ctypes
import os
import re

import pytest

import aligngraph_b200 as ag


def _declared():
    text = open(ag.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ag_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from aligngraph_b200 import build
    build.build()
    return ag.load_library()


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_python_binding_covers_the_header(lib):
    assert sorted(lib._ag_signatures) == _declared()


def test_struct_sizes():
    assert ctypes.sizeof(ag.Params) == 16
    assert ctypes.sizeof(ag.UnitView) == 104


def test_no_cpu_fallback(lib):
    """Without a CUDA device the context cannot be created — the product path never routes through a CPU implementation."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ag.AlignGraphError, match="no CUDA device"):
        ag.Context()


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(ag.AlignGraphError, match="no CPU fallback"):
        ag.load_library(str(tmp_path / "nope.so"))
