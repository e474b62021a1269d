"""CPU-only: the C-ABI library loads without a GPU, exports every symbol include/*.h declares,
and fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            text = open(os.path.join(inc, f)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b(cuspmm_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


@pytest.fixture(scope="module")
def L():
    from __graft_entry__ import load_package
    return load_package().lib()


def test_every_declared_symbol_is_exported(L):
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_version_and_error_string(L):
    assert L.cuspmm_version() == 210
    assert isinstance(L.cuspmm_last_error(), bytes)


def test_no_cpu_fallback_without_device(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    n = ctypes.c_int(-1)
    rc = L.cuspmm_device_count(ctypes.byref(n))
    assert rc == 2 and b"failed" in L.cuspmm_last_error()        # CUSPMM_ERR_CUDA
    from __graft_entry__ import load_package
    b = load_package().binding
    with pytest.raises(b.CuspmmError):
        b.dev_f32([1.0])


def test_argument_validation_needs_no_device(L):
    # variant out of range is rejected before any CUDA call ("Not implemented", engine_csr.hpp:88)
    rc = L.cuspmm_spmm_csr(None, None, None, 4, 4, 0, None, 4, 4, None, 4, 99, None)
    assert rc == 1 and b"variant" in L.cuspmm_last_error()
    rc = L.cuspmm_spmm_csr(None, None, None, 4, 4, 0, None, 8, 4, None, 8, 1, None)
    assert rc == 1                                               # ldb < N


def test_csr_workspace_contract_needs_no_device(L):
    """cuspmm_spmm_csr_workspace: 0 bytes for the row / staged kernels, a positive, nnz- and N-dependent size for the
    row-cutting kernel (variant 6); variant 6 without workspace is refused before any CUDA call; the plain entry point
    never runs it."""
    L.cuspmm_spmm_csr_workspace.restype = ctypes.c_size_t
    L.cuspmm_spmm_csr_workspace.argtypes = [ctypes.c_uint32] * 4 + [ctypes.c_int]
    for v in (1, 2, 3, 4, 5):
        assert L.cuspmm_spmm_csr_workspace(2798, 21074, 81671, 512, v) == 0
    w6 = L.cuspmm_spmm_csr_workspace(2798, 21074, 81671, 512, 6)
    assert w6 > 0 and w6 % 16 == 0
    assert L.cuspmm_spmm_csr_workspace(2798, 21074, 81671, 128, 6) < w6          # fewer 128-column tiles
    assert L.cuspmm_spmm_csr_workspace(2798, 21074, 81671, 512, 0) == w6         # few long rows: the selector may want it
    assert L.cuspmm_spmm_csr_workspace(25605, 25605, 65571195, 512, 0) == 0      # staged territory: never
    rc = L.cuspmm_spmm_csr(None, None, None, 4, 4, 0, None, 4, 4, None, 4, 6, None)
    assert rc != 0 and b"workspace" in L.cuspmm_last_error()


def test_plain_c_client_compiles_links_and_runs(tmp_path):
    """The header is C (not C++) and a gcc-built program links against the shared library."""
    import subprocess
    libdir = os.path.join(ROOT, "cuda-optimization-for-spmm_b200")
    exe = str(tmp_path / "abi_c_client")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_c_client.c"), "-o", exe, "-L", libdir, "-lcuspmm_b200",
                    "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0 and "abi_c_client ok" in p.stdout, p.stdout + p.stderr
