"""The C-ABI library loads and exports every symbol include/b200st.h declares (no compute, no GPU)."""
import ctypes
import os

import pytest

from conftest import PKG, ROOT


def test_header_parses_and_library_exports_all_symbols():
    from b200st import lib
    protos = lib.parse_header()
    assert len(protos) >= 30
    handle = lib.load()
    for name in protos:
        assert hasattr(handle, name), f'{name} declared in include/b200st.h but missing from libb200st.so'
    assert handle.b200st_version() >= 100
    assert isinstance(lib.last_error(), str)


def test_missing_library_fails_loudly(tmp_path):
    from b200st import lib
    with pytest.raises(lib.MissingLibrary):
        saved = lib._lib
        lib._lib = None
        try:
            lib.load(str(tmp_path / 'nope.so'))
        finally:
            lib._lib = saved


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the product package may import or execute it."""
    import re
    pat = re.compile(r'^\s*(from|import)\s+oracle\b|st_oracle|oracle\.', re.M)
    bad = []
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith('.py') and pat.search(open(os.path.join(base, f)).read()):
                bad.append(os.path.join(base, f))
    assert not bad, bad


def test_cpu_tensor_rejected_without_gpu():
    import torch
    from b200st.kernels import CudaKernels
    with pytest.raises(RuntimeError):
        CudaKernels().gemm(torch.randn(2, 2), torch.randn(2, 2))
