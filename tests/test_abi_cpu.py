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
    assert lib.tpgan_abi_version() == 3     # v2: dtype + out16 (bf16 operands); v3: in_lo / w_lo_packed (one-launch 3xTF32)
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


def test_invalid_arguments_are_rejected_before_any_launch():
    """Error behaviour of the boundary: argument validation happens on the host before any CUDA call, returns
    TPGAN_ERR_INVALID (-1) and leaves a message in tpgan_last_error() - checkable without a GPU.  (No compute is launched.)"""
    from tpgan_b200 import _lib
    lib = _lib.load()
    V = _lib.View
    null = _lib.NULL_VIEW
    c0 = lib.tpgan_launch_count()
    fake = 0x1000      # a non-null "pointer" that is never dereferenced: validation fails first
    cases = {
        "multitask_loss": lambda: lib.tpgan_multitask_loss(fake, fake, fake, None, 1, 0, 2, 5, 5, 1, 128.0, 128.0, 30.0, 0.1, 5.0, 1.0,
                                                           None, None, None, fake, None),
        "multitask_loss_classes": lambda: lib.tpgan_multitask_loss(fake, fake, fake, None, 1, 394, 788, 1970, 4, 39, 128.0, 128.0,
                                                                   30.0, 0.1, 5.0, 1.0, None, None, None, fake, None),
        "ssd_decode": lambda: lib.tpgan_ssd_decode(fake, fake, 1, 394, 788, 1970, 5, 0, 0.5, 20.0, fake, fake, fake, None, None, None),
        "sgd_step": lambda: lib.tpgan_sgd_step(fake, fake, fake, 3, fake, 0.9, 5e-4, 1, 1.0, None),            # n % 4 != 0
        "dwconv3x3": lambda: lib.tpgan_dwconv3x3(V(fake, 64 * 6, 8 * 6, 6, 1, 8, 8, 6), V(fake, 64 * 6, 8 * 6, 6, 1, 8, 8, 6),
                                                 fake, 1, None),                                                # C % 4 != 0
        "dwconv3x3_stride": lambda: lib.tpgan_dwconv3x3(V(fake, 512, 64, 8, 1, 8, 8, 8), V(fake, 512, 64, 8, 1, 8, 8, 8), fake, 3, None),
        "bn_forward": lambda: lib.tpgan_bn_forward(V(fake, 512, 64, 8, 1, 8, 8, 8), null, V(fake, 512, 64, 8, 1, 8, 8, 8), None, fake,
                                                   None, None, 0.1, 1e-5, 1, 1, 0.0, 0, fake, fake, None),      # gamma missing
        "bn_forward_res_relu6": lambda: lib.tpgan_bn_forward(V(fake, 512, 64, 8, 1, 8, 8, 8), V(fake, 512, 64, 8, 1, 8, 8, 8),
                                                             V(fake, 512, 64, 8, 1, 8, 8, 8), fake, fake, fake, fake, 0.1, 1e-5, 1, 1, 0.0, 0,
                                                             fake, fake, None),                                 # residual + ReLU6
        "rows_gather": lambda: lib.tpgan_rows_gather(V(fake, 512, 64, 8, 1, 8, 8, 8), fake, 100, 0, 0, None),   # row too short
        "pyramid": lambda: lib.tpgan_pyramid(V(fake, 1, 1, 1, 1, 6, 6, 3), V(fake, 1, 1, 1, 1, 3, 3, 3), V(fake, 1, 1, 1, 1, 1, 1, 3), None),
        "landmarks_reduce": lambda: lib.tpgan_landmarks_reduce(fake, 0, 68, fake, 5, 1.0, 1.0, fake, None),
        "u8_to_nhwc": lambda: lib.tpgan_u8_to_nhwc(None, V(fake, 1, 1, 1, 1, 1, 1, 3), 0, None),
        "adam_step": lambda: lib.tpgan_adam_step(fake, fake, fake, fake, 0, 1e-4, 0.9, 0.999, 1e-8, 0.0, 1, 1.0, None),
    }
    for name, call in cases.items():
        rc = call()
        assert rc == -1, (name, rc)
        assert len(lib.tpgan_last_error()) > 0, name
    assert lib.tpgan_launch_count() == c0      # nothing was launched
