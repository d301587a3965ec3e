"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/asis_b200.h declares, validates arguments (no compute without a GPU), and the product
package has no CPU fallback."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "asis_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asis_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from adaptersis_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libasis_b200.so does not export {n}"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert lib.asis_abi_version() == 1


def test_argument_validation_without_gpu():
    from adaptersis_b200 import _lib
    lib = _lib.load()
    rc = lib.asis_msda_forward(None, 0, None, None, None, None, None, 0, 1, 1, 1, 4, 1, 1, 1, None)
    assert rc == -1 and b"null pointer" in lib.asis_last_error()
    rc = lib.asis_msda_forward(None, 0, None, None, None, None, None, 0, 1, 1, 1, 6, 1, 1, 1, None)
    assert rc == -1 and b"multiple of 4" in lib.asis_last_error()
    rc = lib.asis_gemm(0, None, 0, 8, None, 0, 8, None, 0, 8, 8, 8, 8, 0, None, None, None, None, 0, 0, None)
    assert rc == -1
    assert lib.asis_msda_backward_workspace_bytes(2, 100, 4, 16, 10, 3, 4) >= 2 * 2 * 4 * 100 * 4 + 2 * 4 * 10 * 3 * 4 * 4 * 8
    assert lib.asis_attention_backward_workspace_bytes(0, 1, 10, 2, 16) == 2 * 2 * 100 * 4


def test_no_cpu_fallback():
    import adaptersis_b200 as asis
    m = asis.MSDeformAttn(d_model=32, n_levels=1, n_heads=4, n_points=2)
    q = torch.randn(1, 4, 32)
    ref = torch.rand(1, 4, 1, 2)
    ss = torch.tensor([[2, 2]])
    with pytest.raises(RuntimeError, match="CUDA"):
        m(q, ref, q, ss, torch.tensor([0]))
    blk = asis.Block(dim=32, num_heads=2, init_values=1e-5)
    with pytest.raises(RuntimeError, match="CUDA"):
        blk(torch.randn(1, 3, 32))


def test_missing_library_fails_loudly(monkeypatch):
    from adaptersis_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libasis_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or library fallback"):
        _lib.load()
