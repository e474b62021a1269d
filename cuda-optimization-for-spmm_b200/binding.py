"""ctypes view of the C ABI (include/cuspmm_b200.h) over torch CUDA tensors.

No CPU fallback: if the shared library is missing, or a call fails, this raises.  torch is
used only to own device memory and streams; every multiply goes through libcuspmm_b200.so.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcuspmm_b200.so")
if os.environ.get("CUSPMM_LIB"):          # tuning hook: an alternative build of the same library
    LIB_PATH = os.environ["CUSPMM_LIB"]
_lib = None

U32 = C.c_uint32
SZ = C.c_size_t
P = C.c_void_p


class CuspmmError(RuntimeError):
    pass


def lib():
    """Load libcuspmm_b200.so (built in-tree by __graft_entry__.build()); fail loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CuspmmError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.cuspmm_last_error.restype = C.c_char_p
        L.cuspmm_launch_count.restype = C.c_ulonglong
        L.cuspmm_spmm_coo_workspace.restype = SZ
        L.cuspmm_spmm_coo_workspace.argtypes = [U32, U32, U32, C.c_int]
        L.cuspmm_spmm_csr.argtypes = [P, P, P, U32, U32, U32, P, U32, SZ, P, SZ, C.c_int, P]
        L.cuspmm_csr_selected_variant.argtypes = [U32, U32, U32, U32, C.c_int]
        L.cuspmm_set_csr_tensor_mode.argtypes = [C.c_int]
        L.cuspmm_spmm_csr_workspace.restype = SZ
        L.cuspmm_spmm_csr_workspace.argtypes = [U32, U32, U32, U32, C.c_int]
        L.cuspmm_spmm_csr_ws.argtypes = [P, P, P, U32, U32, U32, P, U32, SZ, P, SZ, C.c_int, P, SZ, P]
        L.cuspmm_spmm_coo.argtypes = [P, P, P, U32, U32, U32, P, U32, SZ, P, SZ, C.c_int, P, SZ, P]
        L.cuspmm_spmm_sell.argtypes = [P, P, P, U32, U32, U32, U32, P, U32, SZ, P, SZ, C.c_int, P]
        L.cuspmm_spmm_bsr_f32.argtypes = [P, P, P, U32, U32, U32, U32, P, U32, SZ, P, SZ, P]
        L.cuspmm_colell_to_csr.argtypes = [P, P, U32, U32, U32, U32, P, P, P, P]
        L.cuspmm_csr_to_sell_count.argtypes = [P, U32, U32, P, C.POINTER(U32), P]
        L.cuspmm_csr_to_sell_fill.argtypes = [P, P, P, U32, U32, P, P, P, P]
        L.cuspmm_csr_to_bsr_count.argtypes = [P, P, U32, U32, U32, U32, U32, P, C.POINTER(U32), P]
        L.cuspmm_csr_to_bsr_fill.argtypes = [P, P, P, U32, U32, U32, U32, U32, U32, P, P, P]
        L.cuspmm_partition_rows_by_nnz.argtypes = [P, U32, U32, U32, C.POINTER(U32), P]
        L.cuspmm_coo_to_csr_rowptrs.argtypes = [P, U32, U32, P, P]
        L.cuspmm_spmm_csr_host.argtypes = [P, P, P, U32, U32, U32, P, U32, P, C.c_int, C.POINTER(C.c_float)]
        L.cuspmm_spmm_csr_host_devB.argtypes = [P, P, P, U32, U32, U32, P, SZ, P, U32, P, C.c_int, C.POINTER(C.c_float)]
        L.cuspmm_spmm_coo_host.argtypes = [P, P, P, U32, U32, U32, P, U32, P, C.c_int, C.POINTER(C.c_float)]
        L.cuspmm_spmm_sell_host.argtypes = [P, P, P, U32, U32, U32, U32, P, U32, P, C.c_int, C.POINTER(C.c_float)]
        L.cuspmm_spmm_bsr_host.argtypes = [P, P, P, U32, U32, U32, U32, P, U32, P, C.c_int, C.POINTER(C.c_float)]
        L.cuspmm_host_pipeline_release.argtypes = [C.c_int]
        L.cuspmm_csr_check_sorted.argtypes = [P, P, U32, U32, C.POINTER(U32), P]
        L.cuspmm_host_alloc.argtypes = [C.POINTER(P), SZ]
        L.cuspmm_host_free.argtypes = [P]
        L.cuspmm_cusparse_spmm.argtypes = [C.c_int, P, P, P, U32, U32, U32, P, U32, P, C.c_int, C.c_int, C.c_int,
                                           C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.cuspmm_cusparse_spmm_bsr.argtypes = [P, P, P, U32, U32, U32, U32, P, U32, P, C.c_int, C.c_int,
                                               C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.cuspmm_cusparse_spmm_blockedell.argtypes = [P, P, P, U32, U32, U32, U32, P, U32, P, C.c_int, C.c_int,
                                                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(U32)]
        L.cuspmm_mgpu_create_csr.argtypes = [C.POINTER(P), C.c_int, C.POINTER(C.c_int), P, P, P, U32, U32, U32, U32]
        L.cuspmm_mgpu_create_coo.argtypes = [C.POINTER(P), C.c_int, C.POINTER(C.c_int), P, P, P, U32, U32, U32, U32]
        L.cuspmm_mgpu_create_sell.argtypes = [C.POINTER(P), C.c_int, C.POINTER(C.c_int), P, P, P, U32, U32, U32, U32, U32]
        L.cuspmm_mgpu_create_bsr.argtypes = [C.POINTER(P), C.c_int, C.POINTER(C.c_int), P, P, P, U32, U32, U32, U32, U32]
        L.cuspmm_mgpu_get_counts.argtypes = [P, C.POINTER(U32)]
        L.cuspmm_mgpu_set_B.argtypes = [P, P, U32]
        L.cuspmm_mgpu_run.argtypes = [P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.cuspmm_mgpu_get_splits.argtypes = [P, C.POINTER(U32)]
        L.cuspmm_mgpu_get_C.argtypes = [P, P]
        L.cuspmm_mgpu_destroy.argtypes = [P]
        if hasattr(L, "cuspmm_bsr_tc_plan_create"):
            L.cuspmm_bsr_tc_plan_create.argtypes = [C.POINTER(P), P, P, P, U32, U32, U32, U32, U32, C.c_int, P]
            L.cuspmm_bsr_tc_prepare_B.argtypes = [P, P, U32, SZ, P]
            L.cuspmm_bsr_tc_run.argtypes = [P, P, SZ, P]
            L.cuspmm_bsr_tc_plan_destroy.argtypes = [P]
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise CuspmmError(f"{what} failed with status {rc}: {lib().cuspmm_last_error().decode()}")


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise CuspmmError("no CUDA device: the SpMM engine has no CPU path")
    return torch


def dev_u32(a, device=None):
    """numpy uint32 -> CUDA tensor (stored as int32 bits)."""
    torch = _torch()
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return torch.from_numpy(a.view(np.int32)).to(device or "cuda")


def dev_f32(a, device=None):
    torch = _torch()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device or "cuda")


def host_u32(t):
    return t.cpu().numpy().view(np.uint32)


def _stream():
    return _torch().cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


# ------------------------------------------------------------------------------- SpMM
_CSR_WS = {}


def spmm_csr(rowPtrs, colIdxs, vals, M, K, B, variant=0, out=None, nnz=None, allow_split=False):
    """variant 0 through cuspmm_spmm_csr keeps the all-variants-bit-identical order; allow_split=True goes through
    cuspmm_spmm_csr_ws, whose selector may pick the row-cutting kernel (variant 6) for few / skewed rows."""
    torch = _torch()
    N = B.shape[1]
    Cm = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=B.device)
    nnz = int(colIdxs.numel()) if nnz is None else nnz
    wsb = lib().cuspmm_spmm_csr_workspace(M, K, nnz, N, variant) if (variant == 6 or allow_split) else 0
    if wsb == 0:
        check(lib().cuspmm_spmm_csr(_ptr(rowPtrs), _ptr(colIdxs), _ptr(vals), M, K, nnz, _ptr(B), N, B.stride(0),
                                    _ptr(Cm), Cm.stride(0), variant, _stream()), f"cuspmm_spmm_csr(variant={variant})")
        return Cm
    # variant 6 (or the selector may want it): caller-provided workspace, kept per (device, size) between calls
    key = (B.device.index, wsb)
    ws = _CSR_WS.get(key)
    if ws is None:
        _CSR_WS.clear()
        ws = _CSR_WS[key] = torch.empty(wsb, dtype=torch.uint8, device=B.device)
    check(lib().cuspmm_spmm_csr_ws(_ptr(rowPtrs), _ptr(colIdxs), _ptr(vals), M, K, nnz, _ptr(B), N, B.stride(0),
                                   _ptr(Cm), Cm.stride(0), variant, _ptr(ws), wsb, _stream()),
          f"cuspmm_spmm_csr_ws(variant={variant})")
    return Cm


CSR_KERNEL_NAMES = {1: "csr_rowsplit_vec", 2: "csr_subwarp_vec", 3: "csr_staged", 4: "csr_rowsplit_scalar", 5: "csr_dual",
                    6: "csr_nnzsplit", 7: "csr_quad", 8: "csr_tensor"}


def csr_selected_variant(M, K, nnz, N, sell=False):
    """The variant the selector (variant 0) runs for this shape on the current device."""
    return int(lib().cuspmm_csr_selected_variant(M, K, nnz, N, 1 if sell else 0))


def set_csr_tensor_mode(mode):
    """1: the selector may choose the tensor-core kernel (variant 8); 0: fp32 FMA kernels only.  Returns the previous mode."""
    return int(lib().cuspmm_set_csr_tensor_mode(1 if mode else 0))


def spmm_coo(rowIdxs, colIdxs, vals, M, K, B, variant=0, out=None):
    torch = _torch()
    N = B.shape[1]
    nnz = int(colIdxs.numel())
    Cm = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=B.device)
    wsb = lib().cuspmm_spmm_coo_workspace(M, nnz, N, variant)
    ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=B.device)
    check(lib().cuspmm_spmm_coo(_ptr(rowIdxs), _ptr(colIdxs), _ptr(vals), M, K, nnz, _ptr(B), N, B.stride(0),
                                _ptr(Cm), Cm.stride(0), variant, _ptr(ws), wsb, _stream()),
          f"cuspmm_spmm_coo(variant={variant})")
    return Cm


def spmm_sell(slicePtrs, colIdxs, vals, M, K, B, variant=0, out=None):
    torch = _torch()
    N = B.shape[1]
    Cm = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=B.device)
    check(lib().cuspmm_spmm_sell(_ptr(slicePtrs), _ptr(colIdxs), _ptr(vals), M, K, 32, int(colIdxs.numel()), _ptr(B), N,
                                 B.stride(0), _ptr(Cm), Cm.stride(0), variant, _stream()), f"cuspmm_spmm_sell(variant={variant})")
    return Cm


def spmm_bsr_f32(blockRowPtrs, blockColIdxs, blocks, numBlockRows, br, bc, K, B, out=None):
    torch = _torch()
    N = B.shape[1]
    Cm = out if out is not None else torch.empty((numBlockRows * br, N), dtype=torch.float32, device=B.device)
    check(lib().cuspmm_spmm_bsr_f32(_ptr(blockRowPtrs), _ptr(blockColIdxs), _ptr(blocks), numBlockRows, br, bc, K,
                                    _ptr(B), N, B.stride(0), _ptr(Cm), Cm.stride(0), _stream()), "cuspmm_spmm_bsr_f32")
    return Cm


class BsrTcPlan:
    """Tensor-core BSR plan (tcgen05): owns the bf16/fp16 re-tiled blocks and B."""

    def __init__(self, blockRowPtrs, blockColIdxs, blocks, numBlockRows, blockSize, K, maxN, dtype="bf16"):
        self.h = P()
        self.M = numBlockRows * blockSize
        self.keep = (blockRowPtrs, blockColIdxs)
        check(lib().cuspmm_bsr_tc_plan_create(C.byref(self.h), _ptr(blockRowPtrs), _ptr(blockColIdxs), _ptr(blocks),
                                              numBlockRows, int(blockColIdxs.numel()), blockSize, K, maxN,
                                              0 if dtype == "bf16" else 1, _stream()), "cuspmm_bsr_tc_plan_create")
        self.N = 0

    def prepare_B(self, B):
        self.N = B.shape[1]
        check(lib().cuspmm_bsr_tc_prepare_B(self.h, _ptr(B), self.N, B.stride(0), _stream()), "cuspmm_bsr_tc_prepare_B")

    def run(self, out=None):
        torch = _torch()
        Cm = out if out is not None else torch.empty((self.M, self.N), dtype=torch.float32, device="cuda")
        check(lib().cuspmm_bsr_tc_run(self.h, _ptr(Cm), Cm.stride(0), _stream()), "cuspmm_bsr_tc_run")
        return Cm

    def close(self):
        if self.h:
            lib().cuspmm_bsr_tc_plan_destroy(self.h)
            self.h = P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------- conversions
def csr_to_sell(rowPtrs, colIdxs, vals, M):
    torch = _torch()
    slices = (M + 31) // 32
    sp = torch.empty(slices + 1, dtype=torch.int32, device=rowPtrs.device)
    slots = U32(0)
    check(lib().cuspmm_csr_to_sell_count(_ptr(rowPtrs), M, 32, _ptr(sp), C.byref(slots), _stream()), "csr_to_sell_count")
    n = max(int(slots.value), 1)
    sc = torch.empty(n, dtype=torch.int32, device=rowPtrs.device)
    sv = torch.empty(n, dtype=torch.float32, device=rowPtrs.device)
    check(lib().cuspmm_csr_to_sell_fill(_ptr(rowPtrs), _ptr(colIdxs), _ptr(vals), M, 32, _ptr(sp), _ptr(sc), _ptr(sv),
                                        _stream()), "csr_to_sell_fill")
    return sp, sc[:slots.value], sv[:slots.value]


def csr_to_bsr(rowPtrs, colIdxs, vals, M, K, br, bc):
    torch = _torch()
    nnz = int(colIdxs.numel())
    nbr = (M + br - 1) // br
    rp = torch.empty(nbr + 1, dtype=torch.int32, device=rowPtrs.device)
    nb = U32(0)
    check(lib().cuspmm_csr_to_bsr_count(_ptr(rowPtrs), _ptr(colIdxs), M, K, nnz, br, bc, _ptr(rp), C.byref(nb), _stream()),
          "csr_to_bsr_count")
    n = int(nb.value)
    ci = torch.empty(max(n, 1), dtype=torch.int32, device=rowPtrs.device)
    bl = torch.empty(max(n, 1) * br * bc, dtype=torch.float32, device=rowPtrs.device)
    check(lib().cuspmm_csr_to_bsr_fill(_ptr(rowPtrs), _ptr(colIdxs), _ptr(vals), M, K, nnz, br, bc, n, _ptr(ci), _ptr(bl),
                                       _stream()), "csr_to_bsr_fill")
    return rp, ci[:n], bl[:n * br * bc]


def colell_to_csr(ellRowIdxs, ellVals, M, K, W, nnz):
    torch = _torch()
    rp = torch.empty(M + 1, dtype=torch.int32, device=ellRowIdxs.device)
    ci = torch.empty(max(nnz, 1), dtype=torch.int32, device=ellRowIdxs.device)
    va = torch.empty(max(nnz, 1), dtype=torch.float32, device=ellRowIdxs.device)
    check(lib().cuspmm_colell_to_csr(_ptr(ellRowIdxs), _ptr(ellVals), M, K, W, nnz, _ptr(rp), _ptr(ci), _ptr(va), _stream()),
          "colell_to_csr")
    return rp, ci[:nnz], va[:nnz]


def coo_to_csr_rowptrs(rowIdxs, M):
    torch = _torch()
    rp = torch.empty(M + 1, dtype=torch.int32, device=rowIdxs.device)
    check(lib().cuspmm_coo_to_csr_rowptrs(_ptr(rowIdxs), M, int(rowIdxs.numel()), _ptr(rp), _stream()), "coo_to_csr_rowptrs")
    return rp


def partition_rows_by_nnz(rowPtrs, M, nnz, parts):
    out = (U32 * (parts + 1))()
    check(lib().cuspmm_partition_rows_by_nnz(_ptr(rowPtrs), M, nnz, parts, out, _stream()), "partition_rows_by_nnz")
    return np.array(list(out), dtype=np.uint32)


# ------------------------------------------------------------------------------- host-buffer path
def pinned(a):
    """numpy array -> pinned torch CPU tensor holding the same data."""
    torch = _torch()
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a).pin_memory()


def spmm_csr_host(rowPtrs_h, colIdxs_h, vals_h, M, K, B_h, C_h, variant=0):
    """All arguments are (pinned) torch CPU tensors; returns device milliseconds."""
    ms = C.c_float(0)
    N = B_h.shape[1]
    check(lib().cuspmm_spmm_csr_host(rowPtrs_h.data_ptr(), colIdxs_h.data_ptr(), vals_h.data_ptr(), M, K,
                                     int(colIdxs_h.numel()), B_h.data_ptr(), N, C_h.data_ptr(), variant, C.byref(ms)),
          "cuspmm_spmm_csr_host")
    return ms.value


def spmm_csr_host_devB(rowPtrs_h, colIdxs_h, vals_h, M, K, B_dev, C_h, variant=0, nnz=None):
    """Host CSR (pinned torch CPU tensors or views of them), B already on the device (torch CUDA tensor; the kernels wait for
    torch's current stream); returns device milliseconds."""
    ms = C.c_float(0)
    N = B_dev.shape[1]
    nnz = int(colIdxs_h.numel()) if nnz is None else nnz
    check(lib().cuspmm_spmm_csr_host_devB(rowPtrs_h.data_ptr(), colIdxs_h.data_ptr(), vals_h.data_ptr(), M, K, nnz,
                                          B_dev.data_ptr(), B_dev.stride(0), _stream(), N, C_h.data_ptr(), variant,
                                          C.byref(ms)), "cuspmm_spmm_csr_host_devB")
    return ms.value


def spmm_coo_host(rowIdxs_h, colIdxs_h, vals_h, M, K, B_h, C_h, variant=0):
    ms = C.c_float(0)
    check(lib().cuspmm_spmm_coo_host(rowIdxs_h.data_ptr(), colIdxs_h.data_ptr(), vals_h.data_ptr(), M, K,
                                     int(colIdxs_h.numel()), B_h.data_ptr(), B_h.shape[1], C_h.data_ptr(), variant,
                                     C.byref(ms)), "cuspmm_spmm_coo_host")
    return ms.value


def spmm_sell_host(slicePtrs_h, colIdxs_h, vals_h, M, K, B_h, C_h, variant=0):
    ms = C.c_float(0)
    check(lib().cuspmm_spmm_sell_host(slicePtrs_h.data_ptr(), colIdxs_h.data_ptr(), vals_h.data_ptr(), M, K, 32,
                                      int(colIdxs_h.numel()), B_h.data_ptr(), B_h.shape[1], C_h.data_ptr(), variant,
                                      C.byref(ms)), "cuspmm_spmm_sell_host")
    return ms.value


def spmm_bsr_host(blockRowPtrs_h, blockColIdxs_h, blocks_h, numBlockRows, br, bc, K, B_h, C_h, variant=1):
    """variant 1: fp32 kernels; 2 / 3: bf16 / fp16 tensor-core plan (cast + re-tiling inside the call)."""
    ms = C.c_float(0)
    check(lib().cuspmm_spmm_bsr_host(blockRowPtrs_h.data_ptr(), blockColIdxs_h.data_ptr(), blocks_h.data_ptr(), numBlockRows,
                                     br, bc, K, B_h.data_ptr(), B_h.shape[1], C_h.data_ptr(), variant, C.byref(ms)),
          "cuspmm_spmm_bsr_host")
    return ms.value


def csr_check_sorted(rowPtrs, colIdxs, M, K):
    """-> number of rows whose column indices are not strictly ascending / out of range (0 = the staged kernels' precondition holds)."""
    bad = U32(0)
    check(lib().cuspmm_csr_check_sorted(_ptr(rowPtrs), _ptr(colIdxs), M, K, C.byref(bad), _stream()), "cuspmm_csr_check_sorted")
    return int(bad.value)


def cusparse_spmm(fmt, rowOrPtr, colIdxs, vals, M, K, B, out, alg=0, warmup=3, iters=10):
    avg, mn = C.c_float(0), C.c_float(0)
    _torch().cuda.synchronize()
    check(lib().cuspmm_cusparse_spmm(fmt, _ptr(rowOrPtr), _ptr(colIdxs), _ptr(vals), M, K, int(colIdxs.numel()), _ptr(B),
                                     B.shape[1], _ptr(out), alg, warmup, iters, C.byref(avg), C.byref(mn)),
          "cuspmm_cusparse_spmm")
    return avg.value, mn.value


def cusparse_spmm_bsr(blockRowPtrs, blockColIdxs, blocks, nbr, nbc, bs, B, out, warmup=2, iters=5):
    avg, mn = C.c_float(0), C.c_float(0)
    _torch().cuda.synchronize()
    check(lib().cuspmm_cusparse_spmm_bsr(_ptr(blockRowPtrs), _ptr(blockColIdxs), _ptr(blocks), nbr, nbc,
                                         int(blockColIdxs.numel()), bs, _ptr(B), B.shape[1], _ptr(out), warmup, iters,
                                         C.byref(avg), C.byref(mn)), "cuspmm_cusparse_spmm_bsr")
    return avg.value, mn.value


def cusparse_spmm_blockedell(blockRowPtrs, blockColIdxs, blocks, nbr, nbc, bs, B, out, warmup=2, iters=5):
    """-> (avg ms, min ms, padded width in blocks) of cuSPARSE's Blocked-ELL SpMM on the BSR operand padded to its longest block row."""
    avg, mn, w = C.c_float(0), C.c_float(0), U32(0)
    _torch().cuda.synchronize()
    check(lib().cuspmm_cusparse_spmm_blockedell(_ptr(blockRowPtrs), _ptr(blockColIdxs), _ptr(blocks), nbr, nbc,
                                                int(blockColIdxs.numel()), bs, _ptr(B), B.shape[1], _ptr(out), warmup, iters,
                                                C.byref(avg), C.byref(mn), C.byref(w)), "cuspmm_cusparse_spmm_blockedell")
    return avg.value, mn.value, int(w.value)


class MgpuPlan:
    """One process driving `ngpus` devices.  fmt = "csr" (rowPtrs, colIdxs, vals), "coo" (rowIdxs, colIdxs, vals),
    "sell" (slicePtrs, colIdxs, vals) or "bsr" (blockRowPtrs, blockColIdxs, blocks; M = numBlockRows, br, bc given): host numpy arrays."""

    def __init__(self, ngpus, a0_h, a1_h, a2_h, M, K, maxN, devices=None, fmt="csr", br=1, bc=1):
        self.h = P()
        self.n, self.M, self.fmt = ngpus, M, fmt
        devs = (C.c_int * ngpus)(*(devices or list(range(ngpus))))
        L = lib()
        if fmt == "csr":
            check(L.cuspmm_mgpu_create_csr(C.byref(self.h), ngpus, devs, a0_h.ctypes.data, a1_h.ctypes.data, a2_h.ctypes.data,
                                           M, K, int(a1_h.shape[0]), maxN), "mgpu_create_csr")
        elif fmt == "coo":
            check(L.cuspmm_mgpu_create_coo(C.byref(self.h), ngpus, devs, a0_h.ctypes.data, a1_h.ctypes.data, a2_h.ctypes.data,
                                           M, K, int(a1_h.shape[0]), maxN), "mgpu_create_coo")
        elif fmt == "sell":
            check(L.cuspmm_mgpu_create_sell(C.byref(self.h), ngpus, devs, a0_h.ctypes.data, a1_h.ctypes.data, a2_h.ctypes.data,
                                            M, K, 32, int(a1_h.shape[0]), maxN), "mgpu_create_sell")
        elif fmt == "bsr":
            check(L.cuspmm_mgpu_create_bsr(C.byref(self.h), ngpus, devs, a0_h.ctypes.data, a1_h.ctypes.data, a2_h.ctypes.data,
                                           M, br, bc, K, maxN), "mgpu_create_bsr")
            self.M = M * br
        else:
            raise ValueError(fmt)

    def set_B(self, B_h):
        self.N = B_h.shape[1]
        check(lib().cuspmm_mgpu_set_B(self.h, B_h.ctypes.data, self.N), "mgpu_set_B")

    def run(self, variant=0, gather=False, iters=1):
        ms = C.c_float(0)
        check(lib().cuspmm_mgpu_run(self.h, variant, int(gather), iters, C.byref(ms)), "mgpu_run")
        return ms.value

    def splits(self):
        out = (U32 * (self.n + 1))()
        check(lib().cuspmm_mgpu_get_splits(self.h, out), "mgpu_get_splits")
        return np.array(list(out), dtype=np.uint32)

    def counts(self):
        out = (U32 * self.n)()
        check(lib().cuspmm_mgpu_get_counts(self.h, out), "mgpu_get_counts")
        return np.array(list(out), dtype=np.uint32)

    def get_C(self):
        out = np.empty((self.M, self.N), dtype=np.float32)
        check(lib().cuspmm_mgpu_get_C(self.h, out.ctypes.data), "mgpu_get_C")
        return out

    def close(self):
        if self.h:
            lib().cuspmm_mgpu_destroy(self.h)
            self.h = P()
