"""CPU: the C-ABI shared library loads and exports every symbol include/ddrl_b200.h declares, the
ctypes table covers exactly those symbols, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ddrl_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ddrl_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def native():
    import __graft_entry__
    __graft_entry__.build()
    from ddrl_b200 import _native
    return _native


def test_header_declares_something():
    syms = declared_symbols()
    assert "ddrl_rb_create" in syms and "ddrl_rb_sample" in syms and len(syms) >= 10


def test_library_exports_every_declared_symbol(native):
    L = C.CDLL(native.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in the header but not exported"


def test_ctypes_table_matches_header(native):
    assert sorted(native.SIGNATURES) == declared_symbols()


def test_abi_version(native):
    assert native.lib().ddrl_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure(native):
    h = C.c_void_p()
    rc = native.lib().ddrl_rb_create(0, 8, 2, 100, C.byref(h))
    assert rc == native.ECUDA and not h.value
    assert b"cuda" in native.lib().ddrl_last_error().lower()
    with pytest.raises(native.NativeError):
        native.check(rc)
    from ddrl_b200 import ReplayBuffer
    with pytest.raises(RuntimeError):
        ReplayBuffer(8, 2, 100)


def test_argument_validation_needs_no_gpu(native):
    L = native.lib()
    assert L.ddrl_rb_create(0, 8, 2, 100, None) == native.EINVAL
    h = C.c_void_p()
    assert L.ddrl_rb_create(0, 0, 2, 100, C.byref(h)) == native.EINVAL
    assert L.ddrl_rb_store_batch(None, None, None, None, None, None, 1, 0, None) == native.EINVAL
    assert L.ddrl_rb_counts(None, None, None, None, None, None) == native.EINVAL
    assert L.ddrl_rb_destroy(None) == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "distributed-drl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
