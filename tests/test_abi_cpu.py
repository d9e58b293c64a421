"""The C-ABI shared library loads without a GPU and exports every symbol include/tpgan_b200.h declares; the ctypes
binding table covers exactly that set.  No compute calls here (no GPU in the CPU suite)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tpgan_b200.h")).read()
    return sorted(set(re.findall(r"TPGAN_API\s+[\w\s\*]+?\b(tpgan_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from tpgan_b200 import _lib
    names = _declared()
    assert len(names) >= 30
    assert os.path.exists(_lib.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SYMBOLS.keys()) == names, (set(names) ^ set(_lib.SYMBOLS.keys()))


def test_abi_version_and_error_string():
    from tpgan_b200 import _lib
    lib = _lib.load()
    assert lib.tpgan_abi_version() == 1
    assert isinstance(lib.tpgan_last_error(), bytes)
    assert lib.tpgan_launch_count() >= 0


def test_product_path_has_no_cpu_fallback():
    """CPU tensors are rejected loudly instead of silently running somewhere else."""
    import torch
    from tpgan_b200 import D_and_G_model as M
    D = M.Discriminator(False)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        D(torch.zeros(1, 3, 128, 128))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        M.LocalFuser()(torch.zeros(1, 3, 40, 40), torch.zeros(1, 3, 40, 40), torch.zeros(1, 3, 32, 40),
                       torch.zeros(1, 3, 32, 48))


def test_missing_library_fails_loudly(monkeypatch):
    from tpgan_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libtpgan_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_product_does_not_import_oracle():
    """Nothing under tpgan_b200/ may import the oracle (test infrastructure only)."""
    pkg = os.path.join(ROOT, "tpgan_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn
